"""Micro-benchmark of the fused rel-pos attention kernel alone (cfg2 layer shape by default)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gpu_util import op_attention, ptr, stream
from conformer_nemo_b200 import _lib

B, T, H, dk = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (32, 500, 8, 64))]
Dp = H * 64
torch.manual_seed(0)
qkv = (torch.randn(B * T, 4 * Dp, device="cuda") * 0.5).bfloat16()
pos = (torch.randn(2 * T - 1, Dp, device="cuda") * 0.5).bfloat16()
ctx = torch.empty(B * T, Dp, device="cuda", dtype=torch.bfloat16)
lens = torch.full((B,), T, dtype=torch.int32, device="cuda")
lib = _lib.load_library()
def run():
    rc = lib.cfb_op_rel_attention(1, ptr(qkv), ptr(pos), pos.stride(0), ptr(ctx), ptr(lens), B, T, H, dk, 64, stream())
    assert rc == 0
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 20
e0.record()
for _ in range(n): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
fl = 6.0 * B * T * T * H * dk
print(f"CFB_ATTN_DEBUG={os.environ.get('CFB_ATTN_DEBUG','0')} B={B} T={T} H={H}: {ms*1e3:.1f} us  {fl/ms/1e9:.1f} TFLOP/s")

if int(os.environ.get("CFB_ATTN_DEBUG", "0")) & 8:
    import ctypes, numpy as np
    buf = (ctypes.c_longlong * 1024)()
    rc = lib.cfb_debug_attn_trace(buf)
    a = np.array(buf[:], dtype=np.int64)
    sm = a[:512].reshape(64, 8); isr = a[512:].reshape(64, 8)
    base = min(x for x in a if x > 0)
    print("softmax warp (set 0): it | wait_sg_full_start, sg_full_done, sv_done, exp_done, p_arrived | o_full_done, fold_done (relative cycles)")
    for it in range(6):
        print(it, [int(x - base) if x > 0 else -1 for x in sm[it, :7]])
    print("issuer (set 0): it | start, loads_ready, sg_free_ok, sg_issued, p_ready_ok, pv_issued")
    for it in range(6):
        print(it, [int(x - base) if x > 0 else -1 for x in isr[it, :6]])
