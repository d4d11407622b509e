"""Micro-benchmark of the CTC head on the cfg2 encoder output (16000 frames x 512, BPE-1024 and char-28 heads)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import conformer_nemo_b200 as cn

x = torch.randn(32, 500, 512, device="cuda").transpose(1, 2)
for v in (1024, 128, 28):
    dec = cn.ConvASRDecoder(feat_in=512, num_classes=v).cuda()
    for _ in range(3): dec.forward_with_predictions(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dec.forward_with_predictions(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    byts = 16000 * 512 * 4 + 16000 * (v + 1) * 4
    print(f"CTC head V+1={v + 1}: {ms * 1e3:.1f} us per batch of 640 audio-s  (GEMM {2 * 16000 * 512 * (v + 1) / 1e9:.1f} GFLOP, in+out {byts / 1e6:.0f} MB)")
