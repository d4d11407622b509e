"""Whole inference path on one B200, host waveforms in -> token ids on the host out (what SURVEY.md 8(f) adds around the
encoder, composed): pinned H2D of 32 x 20 s of 16 kHz audio -> log-mel kernels -> Conformer-L encoder (CUDA-graph replay)
-> (a) CTC head + greedy collapse, or (b) transducer greedy decode.  Random-init weights of the recipes' sizes, synthetic
audio; blank biases calibrated so that the decoders emit a speech-like 0.2 symbols per encoder frame.

    python tools/bench_pipeline.py [--batch 32] [--seconds 20] [--steps 10]

Prints one JSON line with the end-to-end audio-seconds per second of both pipelines and the share of each stage
(CUDA events on the compute stream).
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import conformer_nemo_b200 as cn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--seconds", type=int, default=20)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--decode-group", type=int, default=4,
                    help="transducer: encoder batches decoded together in one launch (the decode kernel is latency-bound "
                         "per lock-step iteration, so it gains from more utterances per launch)")
    args = ap.parse_args()
    from oracle import conformer_oracle as oc  # seeded weight generators only
    from oracle import ctc_head_oracle as ho
    from oracle import rnnt_oracle as ro

    B, L = args.batch, args.seconds * 16000
    dev = torch.device("cuda:0")
    audio = (torch.randn(B, L, generator=torch.Generator().manual_seed(0)) * 0.1).pin_memory()
    lengths = torch.full((B,), L, dtype=torch.int64).pin_memory()
    pre = cn.AudioToMelSpectrogramPreprocessor(window_size=0.025, window_stride=0.01, features=80, n_fft=512, pad_to=0).to(dev)
    cfg = oc.EncoderConfig(feat_in=80, n_layers=17, d_model=512, n_heads=8)
    enc = cn.ConformerEncoder(feat_in=80, n_layers=17, d_model=512, n_heads=8)
    enc.load_state_dict(oc.random_state_dict(cfg, 0), strict=False)
    enc = enc.to(dev).eval()
    enc.enable_cuda_graphs(True)
    head = cn.ConvASRDecoder(feat_in=512, num_classes=1024)
    head_sd = ho.random_head_state_dict(512, 1024, 0)
    head.load_state_dict(head_sd)
    head = head.to(dev)

    def front(a, n):
        feats, flen = pre(input_signal=a, length=n, check_lengths=False)
        return enc(audio_signal=feats, length=flen)

    a_dev, n_dev = audio.to(dev), lengths.to(dev)
    encoded, elen = front(a_dev, n_dev)
    frames = int(elen.sum())

    # calibrate the blank biases to ~0.2 symbols per frame (random weights would emit at every frame)
    def ctc_rate(bias):
        sd = dict(head_sd)
        sd["decoder_layers.0.bias"] = head_sd["decoder_layers.0.bias"].clone()
        sd["decoder_layers.0.bias"][1024] += bias
        head.load_state_dict(sd)
        _, pred = head.forward_with_predictions(encoded)
        return sum(len(s) for s in cn.ctc_greedy_decode(pred, elen, 1024)) / frames

    lo, hi = 0.0, 8.0
    for _ in range(10):
        mid = 0.5 * (lo + hi)
        lo, hi = (mid, hi) if ctc_rate(mid) > 0.2 else (lo, mid)
    ctc_symbols = ctc_rate(hi)

    def rnnt(bias):
        dec_sd, joint_sd = ro.random_rnnt_state_dicts(512, 640, 640, 1024, seed=0, blank_bias=bias)
        dec = cn.RNNTDecoder(prednet=dict(pred_hidden=640, pred_rnn_layers=1, dropout=0.1), vocab_size=1024)
        joint = cn.RNNTJoint(jointnet=dict(encoder_hidden=512, pred_hidden=640, joint_hidden=640, activation="relu", dropout=0.1),
                             num_classes=1024)
        dec.load_state_dict(dec_sd)
        joint.load_state_dict(joint_sd)
        return cn.GreedyBatchedRNNTInfer(dec.to(dev), joint.to(dev), 1024, 30)

    lo, hi = 0.5, 3.0
    for _ in range(10):
        mid = 0.5 * (lo + hi)
        greedy = rnnt(mid)
        rate = greedy.decode_arrays(encoded, elen, max_tokens=30 * encoded.shape[2])["n_tokens"].float().sum().item() / frames
        lo, hi = (mid, hi) if rate > 0.2 else (lo, mid)
    rnnt_symbols = rate

    def run_ctc():
        a, n = audio.to(dev, non_blocking=True), lengths.to(dev, non_blocking=True)
        y, yl = front(a, n)
        _, pred = head.forward_with_predictions(y)
        return cn.ctc_greedy_decode(pred, yl, 1024)  # collapse kernel on the device (metrics/wer.py:152-164), then one read of the ids

    def run_rnnt():
        a, n = audio.to(dev, non_blocking=True), lengths.to(dev, non_blocking=True)
        y, yl = front(a, n)
        return greedy(encoder_output=y, encoded_lengths=yl)[0]

    def run_rnnt_grouped():
        ys, yls = [], []
        for _ in range(args.decode_group):
            a, n = audio.to(dev, non_blocking=True), lengths.to(dev, non_blocking=True)
            y, yl = front(a, n)
            ys.append(y.transpose(1, 2).clone())  # the graph's static output is overwritten by the next replay
            yls.append(yl.clone())
        return greedy(encoder_output=torch.cat(ys, 0).transpose(1, 2), encoded_lengths=torch.cat(yls, 0))[0]

    out = {}
    for name, fn in (("ctc", run_ctc), ("transducer", run_rnnt), (f"transducer_decode_{args.decode_group}_batches_together", run_rnnt_grouped)):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = fn()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / args.steps * 1e3
        nb = args.decode_group if fn is run_rnnt_grouped else 1
        out[name] = dict(value=nb * B * args.seconds / (ms * 1e-3), unit="audio-sec/sec", ms_per_batch=ms / nb)

    # stage shares (device time, CUDA events)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    torch.cuda.synchronize()
    ev[0].record()
    a, n = audio.to(dev, non_blocking=True), lengths.to(dev, non_blocking=True)
    ev[1].record()
    feats, flen = pre(input_signal=a, length=n, check_lengths=False)
    ev[2].record()
    y, yl = enc(audio_signal=feats, length=flen)
    ev[3].record()
    head.forward_with_predictions(y)
    ev[4].record()
    greedy.decode_arrays(y, yl)
    ev[5].record()
    torch.cuda.synchronize()
    stages = dict(zip(["h2d_audio", "log_mel", "encoder", "ctc_head", "transducer_greedy"],
                      [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]))
    print(json.dumps(dict(metric="waveform -> token ids, audio-sec/sec (host in, host out)", pipelines=out, stage_ms=stages,
                          config=dict(workload=f"{B} x {args.seconds} s of 16 kHz audio; log-mel 80; Conformer-L encoder 17 x 512; "
                                               "CTC head 1025 classes | transducer decoder/joint 640/640/1025, max_symbols 30",
                                      ctc_symbols_per_frame=ctc_symbols, transducer_symbols_per_frame=rnnt_symbols,
                                      h2d_bytes_per_batch=B * L * 4 + B * 8),
                          data="synthetic (randn audio, random-init weights, calibrated blank biases)")))


if __name__ == "__main__":
    main()
