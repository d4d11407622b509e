"""Micro-benchmark of the LayerNorm kernels on the cfg2 shape (16000 x 512): L2-hot (same buffers back to back) and
L2-cold (a 256 MB memset between launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from gpu_util import ptr, stream
from conformer_nemo_b200 import _lib

lib = _lib.load_library()
M, d = 16000, 512
x = torch.randn(M, d, device="cuda"); g = torch.ones(d, device="cuda"); b = torch.zeros(d, device="cuda")
a = torch.empty(M, d, device="cuda", dtype=torch.bfloat16)
junk = torch.empty(64 * 1024 * 1024, device="cuda")
def run():
    assert lib.cfb_op_layernorm(ptr(x), ptr(g), ptr(b), ptr(a), _lib.CFB_BF16, M, d, None, 1, stream()) == 0
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): run()
e1.record(); torch.cuda.synchronize()
hot = e0.elapsed_time(e1) / 50
tot = 0.0
for _ in range(10):
    junk.zero_()
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    tot += e0.elapsed_time(e1)
print(f"layernorm 16000x512 fp32->bf16: hot {hot * 1e3:.1f} us ({48e6 / hot / 1e6:.0f} GB/s), cold {tot / 10 * 1e3:.1f} us ({48e6 / (tot / 10) / 1e6:.0f} GB/s)")
# copy kernel reference for the same bytes
y = torch.empty(12_000_000, device="cuda"); z = torch.empty_like(y)
for _ in range(3): z.copy_(y)
torch.cuda.synchronize()
e0.record()
for _ in range(50): z.copy_(y)
e1.record(); torch.cuda.synchronize()
c = e0.elapsed_time(e1) / 50
print(f"torch copy of 48 MB read + 48 MB write: {c * 1e3:.1f} us ({96e6 / c / 1e6:.0f} GB/s)")
