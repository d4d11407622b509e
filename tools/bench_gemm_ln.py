"""Residual GEMM + fused LayerNorm (cfb_op_gemm_ln, gemm_lnc.cu) against the two launches it replaces
(cfb_op_gemm RESID + cfb_op_layernorm / the dual LayerNorm), on the layer shapes of cfg 2 (M = 16000 and the
half-batch M = 8000) and cfg 4.  python tools/bench_gemm_ln.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gpu_util import ptr, stream
from conformer_nemo_b200 import _lib

lib = _lib.load_library()


def timed(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, M, N, K in [("linear2 + norm", 16000, 512, 2048), ("linear_out + norm", 16000, 512, 512),
                      ("linear2 + norm (half batch)", 8000, 512, 2048), ("linear_out + norm (half batch)", 8000, 512, 512),
                      ("linear2 + norm (8-GPU share)", 3176, 512, 2048), ("linear_out + norm (8-GPU share)", 3176, 512, 512),
                      ("cfg4 linear2 + norm", 25600, 256, 1024), ("cfg4 linear_out + norm", 25600, 256, 256)]:
    A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    g1, b1, g2, b2 = (torch.rand(N, device="cuda") + 0.5 for _ in range(4))
    x = torch.randn(M, N, device="cuda")
    a = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)

    def gemm():
        rc = lib.cfb_op_gemm(1, 3, ptr(A), K, ptr(W), K, ptr(bias), None, M, N, K, ptr(x), N, _lib.CFB_F32, 0.5, None, 1, 0, None, stream())
        assert rc == 0, _lib.last_error(None)

    def ln():
        rc = lib.cfb_op_layernorm(ptr(x), ptr(g2), ptr(b2), ptr(a), _lib.CFB_BF16, M, N, None, 1, stream())
        assert rc == 0, _lib.last_error(None)

    def two():
        gemm(); ln()

    def fused(dual):
        def f():
            rc = lib.cfb_op_gemm_ln(ptr(A), K, ptr(W), K, ptr(bias), 0.5, ptr(x), N, ptr(g1) if dual else None, ptr(b1) if dual else None,
                                    ptr(g2), ptr(b2), M, N, K, ptr(a), N, stream())
            assert rc == 0, _lib.last_error(None)
        return f

    t_g, t_two, t_f, t_fd = timed(gemm), timed(two), timed(fused(False)), timed(fused(True))
    print(f"{name:34s} M={M:6d} N={N:4d} K={K:5d}: gemm {t_g:6.1f} us | gemm + layernorm {t_two:6.1f} us | fused {t_f:6.1f} us | fused, two norms {t_fd:6.1f} us")
    if os.environ.get("CFB_LNC_TRACE"):
        import ctypes
        for dual in (False, True):
            fused(dual)()
            buf = (ctypes.c_longlong * 64)()
            lib.cfb_debug_lnc_trace(buf)
            t = list(buf)
            base = min(v for v in t if v > 0)
            for it in range(4):
                r = t[it * 8: it * 8 + 8]
                if r[0] > 0:
                    print(f"   {'dual' if dual else 'one '} tile {it}: start {r[0] - base} | acc ready {r[1] - base} | pass1 +{r[2] - r[1]} | exchange +{r[3] - r[2]} | pass1b+exchange +{r[4] - r[3]} | pass2 +{r[5] - r[4]} | end {r[7] - base}")
            print("        mma (start, acc_empty ok, committed):", [(t[32 + i * 4] - base, t[33 + i * 4] - base, t[34 + i * 4] - base) for i in range(4) if t[32 + i * 4] > 0])
