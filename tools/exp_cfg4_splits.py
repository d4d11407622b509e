"""Experiment: cfg4 (256 x 4 s, d_model 256) as K independent sub-batches on K streams (separate encoder instances, so that
every sub-batch has its own graph) against one batch with the engine's internal two-way micro-batching."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import conformer_nemo_b200 as cn
from oracle import conformer_oracle as oc

kw = dict(n_layers=18, d_model=256, n_heads=4)
B, T = 256, 400
cfg = oc.EncoderConfig(feat_in=80, **kw)
sd = oc.random_state_dict(cfg, 0)
x = torch.randn(B, 80, T, device='cuda'); ln = torch.full((B,), T, dtype=torch.int64, device='cuda')
for K in (1, 2, 4):
    encs, streams = [], []
    for k in range(K):
        e = cn.ConformerEncoder(feat_in=80, **kw); e.load_state_dict(sd, strict=False); e = e.cuda().eval(); e.enable_cuda_graphs(True)
        encs.append(e); streams.append(torch.cuda.Stream())
    n = B // K
    parts = [(x[k*n:(k+1)*n].contiguous(), ln[k*n:(k+1)*n].contiguous()) for k in range(K)]
    def step():
        cur = torch.cuda.current_stream()
        for k in range(K):
            streams[k].wait_stream(cur)
            with torch.cuda.stream(streams[k]):
                encs[k](audio_signal=parts[k][0], length=parts[k][1])
        for k in range(K): cur.wait_stream(streams[k])
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"K={K} sub-batches of {n}: {ms:.3f} ms  {B*T*0.01/ms*1e3:.0f} audio-s/s  (CFB_MICROBATCH={os.environ.get('CFB_MICROBATCH','1')})", flush=True)
