import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch
from gpu_util import ptr, stream
from conformer_nemo_b200 import _lib
lib = _lib.load_library()
d = 512
g = torch.ones(d, device="cuda"); b = torch.zeros(d, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for M in (1000, 2000, 4000, 8000, 16000, 32000, 64000):
    x = torch.randn(M, d, device="cuda"); a = torch.empty(M, d, device="cuda", dtype=torch.bfloat16)
    def run():
        assert lib.cfb_op_layernorm(ptr(x), ptr(g), ptr(b), ptr(a), _lib.CFB_BF16, M, d, None, 1, stream()) == 0
    for _ in range(5): run()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(100): run()
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 100
    # graph replay of 20 launches: no CPU launch gaps
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): run()
        with torch.cuda.graph(gr, stream=s):
            for _ in range(20): 
                assert lib.cfb_op_layernorm(ptr(x), ptr(g), ptr(b), ptr(a), _lib.CFB_BF16, M, d, None, 1, ctypes.c_void_p(s.cuda_stream)) == 0 if False else lib.cfb_op_layernorm(ptr(x), ptr(g), ptr(b), ptr(a), _lib.CFB_BF16, M, d, None, 1, __import__('ctypes').c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    torch.cuda.synchronize()
    gr.replay(); torch.cuda.synchronize()
    e0.record()
    for _ in range(10): gr.replay()
    e1.record(); torch.cuda.synchronize()
    tg = e0.elapsed_time(e1) / 200
    y = torch.empty(M * d * 3 // 8, device="cuda"); z = torch.empty_like(y)   # same bytes: 6 B per element read+write -> 3 B each way
    for _ in range(3): z.copy_(y)
    torch.cuda.synchronize(); e0.record()
    for _ in range(100): z.copy_(y)
    e1.record(); torch.cuda.synchronize()
    tc = e0.elapsed_time(e1) / 100
    print(f"M={M:6d}: layernorm eager {t*1e3:6.1f} us  graph {tg*1e3:6.1f} us | torch copy of the same bytes {tc*1e3:6.1f} us")
