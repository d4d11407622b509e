"""Multi-GPU = pure data parallelism over independent utterances: no collective on the data path.

Host-side only.  ``plan_shards`` assigns utterances to ranks with a longest-processing-time greedy rule on an
analytic cost model and cuts each rank's share into length-bucketed sub-batches so padding stays bounded;
``forward_sharded`` runs one rank's sub-batches through any ``encode(audio_signal, length)`` callable and returns
results on the host in the caller's utterance order.  The reference's only collective on this path
(all_reduce(MAX) of the input length, conformer_encoder.py:283-294) is unnecessary here because positional tables
are per call.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Sequence, Tuple

import torch


def _out_frames(t: int) -> int:
    for _ in range(2):
        t = (t - 1) // 2 + 1
    return t


def utterance_cost(length: int, d_model: int = 512, n_layers: int = 17, conv_channels: int = None,
                   feat_in: int = 80) -> float:
    """Algorithmic FLOPs of one utterance: linear in T' for GEMMs / subsampling, quadratic for attention."""
    c = d_model if conv_channels is None else conv_channels
    t1 = (length - 1) // 2 + 1
    t2 = (t1 - 1) // 2 + 1
    f1 = (feat_in - 1) // 2 + 1
    f2 = (f1 - 1) // 2 + 1
    sub = 2 * 9 * c * t1 * f1 + 2 * 9 * c * c * t2 * f2 + 2 * t2 * (f2 * c) * d_model
    return float(sub + n_layers * (46 * t2 * d_model * d_model + 6 * t2 * t2 * d_model + 62 * t2 * d_model))


@dataclass
class ShardPlan:
    n_ranks: int
    batches: List[List[List[int]]] = field(default_factory=list)  # [rank][sub-batch] -> utterance indices
    cost: List[float] = field(default_factory=list)                # per-rank modelled cost

    def rank_indices(self, rank: int) -> List[int]:
        return [i for sub in self.batches[rank] for i in sub]


# A forward is ~250 dependent kernel launches whatever the batch holds: measured on B200, a sub-batch costs about 1 ms
# of fixed time on top of its arithmetic (profiles/r01r_bench_cfg3.json vs cfg2).  Expressed in the cost model's unit
# (FLOPs at the ~0.7 PFLOP/s the big GEMMs sustain) so that padding and launch overhead can be traded off.
SUB_BATCH_OVERHEAD_FLOPS = 7.0e11


def plan_cost(lengths: Sequence[int], plan: "ShardPlan", cost_fn: Callable[[int], float] = None) -> float:
    """Modelled time of the slowest rank: every sub-batch pays the fixed overhead plus the PADDED cost of its rows."""
    cost_fn = cost_fn or utterance_cost
    worst = 0.0
    for subs in plan.batches:
        t = 0.0
        for sub in subs:
            t += SUB_BATCH_OVERHEAD_FLOPS + len(sub) * cost_fn(max(int(lengths[i]) for i in sub))
        worst = max(worst, t)
    return worst


def plan_shards(lengths: Sequence[int], n_ranks: int, max_batch: int = 64, bucket_frames=128,
                cost_fn: Callable[[int], float] = utterance_cost) -> ShardPlan:
    """Deterministic (every rank computes the same plan from the same lengths; no communication).

    1. sort utterances by length (descending), assign each to the currently cheapest rank (LPT greedy);
    2. inside a rank, walk its utterances in descending length and start a new sub-batch whenever the batch is full
       or the length falls more than ``bucket_frames`` input frames below the sub-batch's longest utterance
       (``bucket_frames="auto"``: the width in {128 .. 2048, unbounded} that minimises ``plan_cost``).
    """
    if n_ranks < 1:
        raise ValueError("n_ranks must be >= 1")
    if bucket_frames == "auto":
        # padding vs. per-sub-batch overhead: take the bucket width with the cheapest modelled slowest rank
        best = None
        for width in (128, 256, 512, 1024, 2048, 1 << 30):
            cand = plan_shards(lengths, n_ranks, max_batch, width, cost_fn)
            c = plan_cost(lengths, cand, cost_fn)
            if best is None or c < best[0]:
                best = (c, cand)
        return best[1]
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0.0] * n_ranks
    per_rank: List[List[int]] = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda k: (loads[k], k))
        per_rank[r].append(i)
        loads[r] += cost_fn(int(lengths[i]))
    plan = ShardPlan(n_ranks=n_ranks, cost=loads)
    for r in range(n_ranks):
        subs: List[List[int]] = []
        cur: List[int] = []
        head = 0
        for i in per_rank[r]:
            li = int(lengths[i])
            if cur and (len(cur) >= max_batch or head - li > bucket_frames):
                subs.append(cur)
                cur = []
            if not cur:
                head = li
            cur.append(i)
        if cur:
            subs.append(cur)
        plan.batches.append(subs)
    return plan


def forward_sharded(encode: Callable[[torch.Tensor, torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                    features: Sequence[torch.Tensor], plan: ShardPlan, rank: int,
                    device=None, n_streams: int = 3) -> Dict[int, Tuple[torch.Tensor, int]]:
    """Runs rank ``rank``'s sub-batches.  ``features[i]`` is utterance i as a (feat_in, len_i) host tensor.
    Returns {utterance index: (encoded (d_out, T'_i) on the host, T'_i)} -- per-rank outputs go back to the host,
    nothing is exchanged between ranks.  ``encode`` is any ``(audio_signal, length) -> (encoded, encoded_len)`` callable;
    a ``ConformerEncoder`` also receives the lengths as host values (so ragged sub-batches run in its packed layout) and
    runs the rank's sub-batches through ``forward_many`` (concurrently when its CUDA graphs own their workspaces)."""
    results: Dict[int, Tuple[torch.Tensor, int]] = {}
    staged = []
    for sub in plan.batches[rank]:
        lens = [int(features[i].shape[1]) for i in sub]
        t_max = max(lens)
        batch = torch.zeros(len(sub), features[sub[0]].shape[0], t_max, dtype=features[sub[0]].dtype)
        for row, i in enumerate(sub):
            batch[row, :, : lens[row]] = features[i]
        length = torch.tensor(lens, dtype=torch.int64)
        if device is not None:
            batch = batch.pin_memory().to(device, non_blocking=True) if torch.device(device).type == "cuda" else batch.to(device)
            length = length.to(device)
        staged.append((sub, batch, length, lens))
    if hasattr(encode, "forward_many"):
        outs = encode.forward_many([(batch, length, lens) for _, batch, length, lens in staged], n_streams)
    else:
        outs = [encode(batch, length) for _, batch, length, _ in staged]
    for (sub, _, _, _), (encoded, enc_len) in zip(staged, outs):
        encoded = encoded.detach().to("cpu")
        enc_len = enc_len.detach().to("cpu")
        for row, i in enumerate(sub):
            n = int(enc_len[row])
            results[i] = (encoded[row, :, :n].clone(), n)
    return results
