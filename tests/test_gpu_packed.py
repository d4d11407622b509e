"""GPU: the packed (variable-length) forward, cfb_forward_packed, against the dense forward and the oracle.

The contract (include/cfb.h): same inputs, same outputs, bit-identical to cfb_forward on every frame t < encoded_len[b],
zeros behind it -- including the reference's padded-batch edge effect at the end of every utterance
(subsampling.py:172-175 runs the strided convolutions over the padded batch).  The sharded path of a mixed-length batch
(sharding.plan_shards -> forward_many, the reference side being BucketingDataset + DDP, audio_to_text.py:1488-1534) is
checked per utterance against the oracle on exactly the planned sub-batches.
"""
import random

import pytest
import torch

import conformer_nemo_b200 as cn
from conformer_nemo_b200.sharding import forward_sharded, plan_shards
from oracle import conformer_oracle as oc

pytestmark = pytest.mark.gpu


def build(cfg, sd):
    enc = cn.ConformerEncoder(feat_in=cfg.feat_in, n_layers=cfg.n_layers, d_model=cfg.d_model, n_heads=cfg.n_heads,
                              conv_kernel_size=cfg.conv_kernel_size, untie_biases=cfg.untie_biases)
    enc.load_state_dict(sd, strict=False)
    return enc.cuda().eval()


def dense_and_packed(enc, x, lens):
    xd, ld = x.cuda(), lens.cuda()
    enc.packed = False
    yd, yld = enc(audio_signal=xd, length=ld)
    yd, yld = yd.clone(), yld.clone()
    enc.packed = True
    yp, ylp = enc(audio_signal=xd, length=ld, length_host=lens.tolist())
    torch.cuda.synchronize()
    return yd, yld, yp, ylp


# lengths chosen to hit: T1 odd / even, T' odd / even, the longest row (zero padding behind it), a row one frame shorter
# than the longest, very short rows (a slot that is almost all gap), a zero-length row
LENGTH_SETS = [
    [400, 399, 398, 397, 396, 395, 394, 393],
    [1000, 37, 640, 333, 5, 999, 2, 1],
    [257, 513, 129, 65, 33, 17, 9, 0],
    [801, 803, 802, 804],
    [1, 1, 1],
    [2000],
]


@pytest.mark.parametrize("lens", LENGTH_SETS)
@pytest.mark.parametrize("dm", [(256, 4), (512, 8)])
def test_packed_forward_is_bit_identical_to_dense_forward(lens, dm):
    d, heads = dm
    cfg = oc.EncoderConfig(feat_in=80, n_layers=2, d_model=d, n_heads=heads)
    enc = build(cfg, oc.random_state_dict(cfg, 3))
    t = max(max(lens), 4)
    x, length = oc.synthetic_batch(len(lens), 80, t, lens, seed=7)
    # NON-zero features behind every utterance: the convolutions must read what the caller passed, as the reference does
    g = torch.Generator().manual_seed(11)
    for b, n in enumerate(lens):
        x[b, :, n:] = torch.randn(80, t - n, generator=g)
    yd, yld, yp, ylp = dense_and_packed(enc, x, length)
    assert torch.equal(yld, ylp)
    assert yd.shape == yp.shape
    assert not torch.isnan(yp).any()
    for b in range(len(lens)):
        n = int(yld[b])
        assert torch.equal(yd[b, :, :n], yp[b, :, :n]), (b, lens[b], float((yd[b, :, :n] - yp[b, :, :n]).abs().max()))
        assert float(yp[b, :, n:].abs().max()) == 0.0 if n < yp.shape[2] else True


def test_packed_forward_against_oracle_full_depth():
    cfg = oc.EncoderConfig(feat_in=80, n_layers=17, d_model=512, n_heads=8)
    sd = oc.random_state_dict(cfg, 0)
    enc = build(cfg, sd)
    lens = [1203, 310, 777, 50]
    x, length = oc.synthetic_batch(len(lens), 80, max(lens), lens, seed=5)
    want, want_len = oc.encoder_forward(sd, cfg, x, length)
    enc.packed = True
    y, ylen = enc(audio_signal=x.cuda(), length=length.cuda(), length_host=lens)
    assert torch.equal(ylen.cpu(), want_len)
    m = (torch.arange(want.shape[2])[None] < want_len[:, None].long())[:, None, :].expand_as(want)
    gg, ww = y.cpu().double()[m], want.double()[m]
    rel, mx = float((gg - ww).norm() / ww.norm()), float((gg - ww).abs().max())
    assert rel <= 1e-2 and mx <= 5e-2, (rel, mx)


def test_packed_graph_replay_and_persistent_attention_match_eager(monkeypatch):
    cfg = oc.EncoderConfig(feat_in=80, n_layers=3, d_model=256, n_heads=4)
    enc = build(cfg, oc.random_state_dict(cfg, 1))
    lens = [900, 333, 120, 64, 700]
    x, length = oc.synthetic_batch(len(lens), 80, max(lens), lens, seed=2)
    xd, ld = x.cuda(), length.cuda()
    enc.packed = True
    y0, l0 = enc(audio_signal=xd, length=ld, length_host=lens)
    y0 = y0.clone()
    enc.enable_cuda_graphs(True)
    for _ in range(3):
        y1, l1 = enc(audio_signal=xd, length=ld, length_host=lens)
    assert torch.equal(y0, y1) and torch.equal(l0, l1)
    # another set of lengths of the same (B, T) shape is another graph, not a stale replay
    lens2 = [900, 100, 800, 64, 200]
    x2, length2 = oc.synthetic_batch(len(lens2), 80, max(lens2), lens2, seed=2)
    y2, _ = enc(audio_signal=x2.cuda(), length=length2.cuda(), length_host=lens2)
    enc.enable_cuda_graphs(False)
    y2e, _ = enc(audio_signal=x2.cuda(), length=length2.cuda(), length_host=lens2)
    assert torch.equal(y2, y2e)
    monkeypatch.setenv("CFB_ATTN_PERSIST", "1")
    y3, _ = enc(audio_signal=xd, length=ld, length_host=lens)
    assert torch.equal(y0, y3)


@pytest.mark.parametrize("groups", [1, 2, 3, 4])
def test_packed_groups_on_their_own_streams_are_bit_identical(groups, monkeypatch):
    """A small packed batch runs as interleaved groups of utterances on separate streams (engine.cu, packed_groups)."""
    cfg = oc.EncoderConfig(feat_in=80, n_layers=2, d_model=256, n_heads=4)
    enc = build(cfg, oc.random_state_dict(cfg, 4))
    lens = [700, 650, 420, 333, 200, 64, 31]
    x, length = oc.synthetic_batch(len(lens), 80, max(lens), lens, seed=9)
    enc.packed = False
    yd, yld = enc(audio_signal=x.cuda(), length=length.cuda())
    yd = yd.clone()
    monkeypatch.setenv("CFB_PACKED_GROUPS", str(groups))
    enc.packed = True
    n0 = None
    for _ in range(2):
        yp, ylp = enc(audio_signal=x.cuda(), length=length.cuda(), length_host=lens)
        n0 = enc.last_launch_count()
    assert torch.equal(yld, ylp)
    for b in range(len(lens)):
        n = int(yld[b])
        assert torch.equal(yd[b, :, :n], yp[b, :, :n]), (groups, b)
        assert float(yp[b, :, n:].abs().sum()) == 0.0
    assert n0 > 0


def test_auto_mode_packs_only_ragged_batches_with_host_lengths():
    cfg = oc.EncoderConfig(feat_in=80, n_layers=1, d_model=256, n_heads=4)
    enc = build(cfg, oc.random_state_dict(cfg, 1))
    assert enc._want_packed((400, 100, 50), 3, 400, 100)
    assert not enc._want_packed((400, 400, 390), 3, 400, 100)
    assert not enc._want_packed(None, 3, 400, 100)
    x, length = oc.synthetic_batch(3, 80, 400, [400, 100, 50], seed=2)
    ya, _ = enc(audio_signal=x.cuda(), length=length)            # CPU lengths: packed, no sync needed
    n_packed = enc.last_launch_count()
    yb, _ = enc(audio_signal=x.cuda(), length=length.cuda())     # device lengths only: dense
    assert enc.last_launch_count() != n_packed
    for b, n in enumerate([100, 25, 13]):
        assert torch.equal(ya[b, :, :n], yb[b, :, :n])


@pytest.mark.parametrize("world", [1, 2, 8])
def test_sharded_cfg3_plan_through_cuda_encoder_matches_oracle_on_planned_sub_batches(world):
    """the cfg-3 plan (64 utterances, seed 1234) for every rank through forward_sharded on the CUDA
    encoder (forward_many underneath, packed sub-batches), each utterance against the oracle run on exactly the planned
    sub-batch (an utterance's last frame depends on the padded extent of the batch it is in)."""
    cfg = oc.EncoderConfig(feat_in=80, n_layers=2, d_model=256, n_heads=4)
    sd = oc.random_state_dict(cfg, 0)
    enc = build(cfg, sd)
    rnd = random.Random(1234)
    lengths = [rnd.randint(200, 3000) for _ in range(64)]
    g = torch.Generator().manual_seed(1234)
    feats = [torch.randn(80, n, generator=g) for n in lengths]
    plan = plan_shards(lengths, world, bucket_frames="auto")
    seen = set()
    ranks = range(world) if world <= 2 else (0, world - 1)
    for rank in ranks:
        got = forward_sharded(enc, feats, plan, rank, device="cuda")
        for sub in plan.batches[rank]:
            lens = [lengths[i] for i in sub]
            x = torch.zeros(len(sub), 80, max(lens))
            for row, i in enumerate(sub):
                x[row, :, : lens[row]] = feats[i]
            want, want_len = oc.encoder_forward(sd, cfg, x, torch.tensor(lens))
            for row, i in enumerate(sub):
                n = int(want_len[row])
                y, ny = got[i]
                assert ny == n and y.shape == (cfg.d_model, n)
                w = want[row, :, :n].double()
                rel = float((y.double() - w).norm() / w.norm())
                mx = float((y.double() - w).abs().max())
                assert rel <= 1e-2 and mx <= 5e-2, (rank, i, rel, mx)
                seen.add(i)
    if world <= 2:
        assert seen == set(range(64))
