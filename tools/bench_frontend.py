"""Throughput of the log-mel front-end kernels (cfb_op_logmel through the AudioToMelSpectrogramPreprocessor drop-in) on
the cfg2-equivalent input (32 utterances x 20 s of 16 kHz audio), with the CPU oracle timed beside it on a bounded
sample.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import conformer_nemo_b200 as cn
from oracle import frontend_oracle as fo

B, SEC = [int(v) for v in (sys.argv[1:3] if len(sys.argv) >= 3 else (32, 20))]
L = SEC * 16000
g = torch.Generator().manual_seed(0)
x = torch.randn(B, L, generator=g) * 0.1
lengths = torch.full((B,), L, dtype=torch.int64)
pre = cn.AudioToMelSpectrogramPreprocessor(window_size=0.025, window_stride=0.01, features=80, n_fft=512, pad_to=0)
want, wl = fo.filterbank_features(x[:2], lengths[:2], pre.featurizer.window, pre.featurizer.fb[0], pad_to=0)
t0 = time.perf_counter(); fo.filterbank_features(x[:2], lengths[:2], pre.featurizer.window, pre.featurizer.fb[0], pad_to=0)
cpu_s = time.perf_counter() - t0
pre = pre.cuda()
xd, ld = x.cuda(), lengths.cuda()
for _ in range(3): y, yl = pre(input_signal=xd, length=ld, check_lengths=False)
torch.cuda.synchronize()
err = float((y[:2].cpu() - want).abs().max())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 50
e0.record()
for _ in range(n): pre(input_signal=xd, length=ld, check_lengths=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
T = y.shape[2]
alg_bytes = B * L * 4 + B * 80 * T * 4 * 3   # audio in; features written, then read + written once by the normalisation
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {"hbm_gbs": 6454.6}
print(json.dumps({"metric": "log-mel front-end audio-sec/sec", "value": B * SEC / (ms / 1e3), "unit": "audio-sec/sec", "ms_per_call": ms,
                  "config": {"workload": f"{B} x {SEC} s of 16 kHz audio -> (B, 80, {T}) normalised log-mel", "n_fft": 512, "win": 400, "hop": 160},
                  "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms / 1e3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                               "frac": alg_bytes / (ms / 1e3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": alg_bytes},
                  "max_abs_vs_oracle": err,
                  "cpu_baseline": {"value": 2 * SEC / cpu_s, "unit": "audio-sec/sec", "kind": "port", "cores": torch.get_num_threads(),
                                   "sample": f"2 x {SEC} s through oracle/frontend_oracle.py (torch.stft etc.), one call"}}))
