"""cfg 2 (32 x 20 s) under different micro-batch schedules, CUDA-graph replay, device-resident inputs:
whole batch in one call (two half-batches on two streams inside cfb_forward), or 2 / 4 calls of 16 / 8 utterances run one
after the other (smaller working set, more launches) or concurrently (forward_many).  python tools/exp_schedules.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import conformer_nemo_b200 as cn
from oracle import conformer_oracle as oc

cfg = oc.EncoderConfig(feat_in=80, n_layers=17, d_model=512, n_heads=8)
sd = oc.random_state_dict(cfg, 0)
enc = cn.ConformerEncoder(feat_in=80, n_layers=17, d_model=512, n_heads=8)
enc.load_state_dict(sd, strict=False)
enc = enc.cuda().eval()
B, T = 32, 2000
x = torch.randn(B, 80, T, device="cuda")
ln = torch.full((B,), T, dtype=torch.int64, device="cuda")


def timed(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


enc.enable_cuda_graphs(True, max_shapes=8, private_workspaces=True)
for parts in (1, 2, 4):
    nb = B // parts
    chunks = [(x[i * nb:(i + 1) * nb].contiguous(), ln[i * nb:(i + 1) * nb].contiguous(), None) for i in range(parts)]
    def seq():
        for xd, ld, _ in chunks:
            enc(audio_signal=xd, length=ld)
    def conc():
        enc.forward_many(chunks, parts)
    t_seq = timed(seq)
    t_con = timed(conc) if parts > 1 else float("nan")
    print(f"{parts} call(s) of {nb:2d} utterances: one after the other {t_seq:6.3f} ms | concurrently {t_con:6.3f} ms   (CFB_MICROBATCH={os.environ.get('CFB_MICROBATCH', 'default')})")
