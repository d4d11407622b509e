"""Clock trace of the LayerNorm-prologue GEMM (gemm_lnt.cu) on the cfg2 linear1 shape: CFB_LNT_TRACE=1."""
import ctypes, os, sys
os.environ["CFB_LNT_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gpu_util import ptr, stream
from conformer_nemo_b200 import _lib
lib = _lib.load_library()
M, d, N = 16000, 512, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x = torch.randn(M, d, device="cuda"); gam = torch.ones(d, device="cuda"); bet = torch.zeros(d, device="cuda")
W = (torch.randn(N, d, device="cuda") / d ** 0.5).bfloat16(); bias = torch.randn(N, device="cuda")
out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    assert lib.cfb_op_gemm_lnt(1, ptr(x), d, None, None, None, ptr(gam), ptr(bet), ptr(W), d, ptr(bias), None, M, N, d, ptr(out), N, None, 1, 0, stream()) == 0
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
assert lib.cfb_debug_lnt_trace(buf) == 0
t = list(buf); t0 = t[0]
r = lambda v: v - t0 if v else -1
print("quarters normalised:", [r(t[i]) for i in range(1, 5)], " A complete:", r(t[5]))
for nt in range(N // 128):
    print(f"tile {nt:2d}: epi wait {r(t[16+4*nt]):7d} acc ready {r(t[17+4*nt]):7d} done {r(t[18+4*nt]):7d} | mma start {r(t[128+4*nt]):7d} committed {r(t[129+4*nt]):7d}")
print("producer issue times (k-block index: cycles):", [(i, r(t[256 + i])) for i in range(0, 64, 4)])
for i in range(8):
    print(f"row {i}: loop top {r(t[400+8*i]):6d} ring ready {r(t[401+8*i]):6d} stats {r(t[402+8*i]):6d} refill issued {r(t[403+8*i]):6d} stored {r(t[404+8*i]):6d}")
