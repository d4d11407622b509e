"""Every product kernel once, on small shapes, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py

Small shapes keep the instrumented run short; the calls go through the same C ABI as the tests.  Prints one line per
stage so a sanitizer report can be matched to the kernel that was running."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import conformer_nemo_b200 as cn  # noqa: E402
import gpu_util as gu  # noqa: E402
from conformer_nemo_b200 import _lib  # noqa: E402
from oracle import conformer_oracle as oc  # noqa: E402
from oracle import rnnt_oracle as ro  # noqa: E402


def stage(name):
    torch.cuda.synchronize()
    print("stage:", name, flush=True)


def main():
    only = set(sys.argv[1:])
    want = lambda n: not only or n in only
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    if want("gemm"):
        M, N, K = 200, 256, 128
        A = torch.randn(M, K, generator=g).to(dev).bfloat16()
        W = torch.randn(N, K, generator=g).to(dev).bfloat16()
        bias = torch.randn(N, generator=g).to(dev)
        for epi, name in ((_lib.EPI_LINEAR, "linear"), (_lib.EPI_SWISH, "swish"), (_lib.EPI_RELU, "relu")):
            stage("gemm " + name)
            gu.op_gemm(True, epi, A, W, bias, out=torch.empty(M, N, dtype=torch.bfloat16, device=dev))
        stage("gemm resid")
        gu.op_gemm(True, _lib.EPI_RESID, A, W, bias, out=torch.zeros(M, N, dtype=torch.float32, device=dev), alpha=0.5)
        stage("gemm glu")
        lens = torch.tensor([90, 100], dtype=torch.int32, device=dev)
        gu.op_gemm(True, _lib.EPI_GLU, A, W, bias, out=torch.empty(M, N // 2, dtype=torch.bfloat16, device=dev), lens=lens,
                   frames_per_seq=100)
        stage("gemm qkv")
        Dp = 128
        Wq = torch.randn(3 * Dp, K, generator=g).to(dev).bfloat16()
        gu.op_gemm(True, _lib.EPI_QKV, A, Wq, torch.randn(3 * Dp, generator=g).to(dev), torch.randn(Dp, generator=g).to(dev),
                   out=torch.empty(M, 4 * Dp, dtype=torch.bfloat16, device=dev), qkv_dp=Dp)
        stage("gemm cta pair (N 2048)")
        W2 = torch.randn(2048, K, generator=g).to(dev).bfloat16()
        gu.op_gemm(True, _lib.EPI_SWISH, A, W2, torch.randn(2048, generator=g).to(dev),
                   out=torch.empty(M, 2048, dtype=torch.bfloat16, device=dev))
    if want("elementwise"):
        stage("layernorm")
        x = torch.randn(300, 512, generator=g).to(dev)
        gu.op_layernorm(x, torch.ones(512, device=dev), torch.zeros(512, device=dev),
                        torch.empty(300, 512, dtype=torch.bfloat16, device=dev))
        stage("depthwise")
        xb = torch.randn(2, 150, 256, generator=g).to(dev).bfloat16()
        gu.op_depthwise(xb, torch.randn(256, 31, generator=g).to(dev), torch.randn(256, generator=g).to(dev), torch.empty_like(xb))
        stage("fused conv tail")
        gu.op_dw_pw2(xb, torch.randn(256, 31, generator=g).to(dev) * 0.1, torch.randn(256, generator=g).to(dev),
                     (torch.randn(256, 256, generator=g) * 0.05).to(dev).bfloat16(), torch.randn(256, generator=g).to(dev),
                     torch.zeros(2, 150, 256, device=dev))
    if want("attention"):
        for persist in ("0", "1"):
            os.environ["CFB_ATTN_PERSIST"] = persist
            stage("attention persist=" + persist)
            B, T, H, dk = 2, 200, 2, 64
            qkv = (torch.randn(B * T, 4 * H * 64, generator=g) * 0.3).to(dev).bfloat16()
            pos = (torch.randn(2 * T - 1, H * 64, generator=g) * 0.3).to(dev).bfloat16()
            gu.op_attention(True, qkv, pos, torch.empty(B * T, H * 64, dtype=torch.bfloat16, device=dev),
                            torch.tensor([200, 77], dtype=torch.int32, device=dev), B, T, H, dk)
        os.environ.pop("CFB_ATTN_PERSIST", None)
    if want("encoder"):
        cfg = oc.EncoderConfig(feat_in=80, n_layers=1, d_model=256, n_heads=4)
        enc = cn.ConformerEncoder(feat_in=80, n_layers=1, d_model=256, n_heads=4)
        enc.load_state_dict(oc.random_state_dict(cfg, 0), strict=False)
        enc = enc.cuda().eval()
        lens = [300, 120, 37]
        x, length = oc.synthetic_batch(3, 80, 300, lens, seed=1)
        stage("encoder dense")
        enc.packed = False
        enc(audio_signal=x.cuda(), length=length.cuda())
        stage("encoder packed (one group)")
        os.environ["CFB_PACKED_GROUPS"] = "1"
        enc.packed = True
        enc(audio_signal=x.cuda(), length=length.cuda(), length_host=lens)
        torch.cuda.synchronize()
        stage("encoder packed (two groups on two streams)")
        os.environ["CFB_PACKED_GROUPS"] = "2"
        enc(audio_signal=x.cuda(), length=length.cuda(), length_host=lens)
        torch.cuda.synchronize()
        os.environ.pop("CFB_PACKED_GROUPS", None)
    if want("ctc"):
        stage("ctc head + collapse")
        dec = cn.ConvASRDecoder(feat_in=256, num_classes=28).cuda()
        y = torch.randn(2, 256, 40, generator=g).to(dev)
        dec.greedy_tokens(y, torch.tensor([40, 17], dtype=torch.int32, device=dev))
    if want("frontend"):
        stage("log-mel front-end")
        pre = cn.AudioToMelSpectrogramPreprocessor().cuda()
        pre(input_signal=torch.randn(2, 16000, generator=g).to(dev), length=torch.tensor([16000, 9000], device=dev))
    if want("rnnt"):
        stage("rnnt greedy decode")
        dims = (64, 64, 64, 28)
        dec_sd, joint_sd = ro.random_rnnt_state_dicts(*dims, seed=0, blank_bias=0.4)
        dec = cn.RNNTDecoder(prednet=dict(pred_hidden=64, pred_rnn_layers=1, dropout=0.1), vocab_size=28)
        joint = cn.RNNTJoint(jointnet=dict(encoder_hidden=64, pred_hidden=64, joint_hidden=64, activation="relu", dropout=0.1),
                             num_classes=28)
        dec.load_state_dict(dec_sd)
        joint.load_state_dict(joint_sd)
        greedy = cn.GreedyBatchedRNNTInfer(dec.cuda(), joint.cuda(), blank_index=28, max_symbols_per_step=5)
        greedy(encoder_output=torch.randn(2, 64, 20, generator=g).to(dev), encoded_lengths=torch.tensor([20, 11], device=dev))
    stage("done")


if __name__ == "__main__":
    main()
