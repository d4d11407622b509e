"""Micro-benchmark of the fused residual + LayerNorm GEMM alone (cfg2 layer shapes)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gpu_util import ptr, stream
from conformer_nemo_b200 import _lib

lib = _lib.load_library()
M, N = 16000, 512
for K in (512, 2048):
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    x = torch.randn(M, N, device="cuda")
    g = torch.ones(N, device="cuda"); b = torch.zeros(N, device="cuda")
    a = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    junk = torch.empty(64 * 1024 * 1024, device="cuda")
    for mode in ("ln2", "ln1+ln2"):
        g1 = g if mode == "ln1+ln2" else None
        def run():
            rc = lib.cfb_op_gemm_ln(ptr(A), K, ptr(W), K, ptr(bias), 0.5, ptr(x), N, ptr(g1), ptr(b if g1 is not None else None),
                                    ptr(g), ptr(b), M, N, K, ptr(x), N, ptr(a), N, None, 1, stream())
            assert rc == 0, _lib.last_error(None)
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        n = 10
        for _ in range(n):
            junk.zero_()  # flush L2
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        e0.record()
        for _ in range(n): run()
        e1.record(); torch.cuda.synchronize()
        hot = e0.elapsed_time(e1) / n
        print(f"K={K} {mode}: cold {tot / n * 1e3:.1f} us, back-to-back {hot * 1e3:.1f} us  ({2.0 * M * N * K / hot / 1e9:.0f} TFLOP/s)")
        if os.environ.get("CFB_LN_TRACE"):
            buf = (ctypes.c_longlong * 16)()
            lib.cfb_debug_ln_trace(buf)
            t = list(buf)
            print("  epilogue warp 0 of CTA 0 (cycles): start->acc_full %d | pass1 %d | combine %d | pass2 %d | combine %d | ln2-stats %d | out pass %d | drain %d | total %d"
                  % (t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5], t[7] - t[6], t[8] - t[7], t[9] - t[0]))
