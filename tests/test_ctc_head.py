"""CTC head: oracle vs the reference's golden vectors and host logic on CPU; the CUDA head
(through cfb_op_ctc_head) against the oracle on the GPU."""
import glob
import os

import numpy as np
import pytest
import torch

import conformer_nemo_b200 as cn
from oracle import ctc_head_oracle as ho

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "ctc_head_*.npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    x = torch.from_numpy(z["encoder_output"])
    v = int(z["num_classes"])
    if "weight" in z.files:
        sd = {"decoder_layers.0.weight": torch.from_numpy(z["weight"]), "decoder_layers.0.bias": torch.from_numpy(z["bias"])}
    else:
        sd = ho.random_head_state_dict(x.shape[1], v, int(z["weight_seed"]))
    assert abs(float(sd["decoder_layers.0.weight"].double().sum()) - float(z["weight_checksum"])) < 1e-6
    return x, v, sd, torch.from_numpy(z["log_probs"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    x, v, sd, want = load(name)
    got = ho.ctc_head_forward(sd, x)
    assert got.shape == want.shape == (x.shape[0], x.shape[2], v + 1)
    assert float((got - want).abs().max()) <= 2e-6


def test_oracle_matches_live_reference_when_mounted():
    from oracle import reference_loader as rl

    if not rl.reference_available():
        pytest.skip("reference tree not mounted")
    cls = rl.load_reference_ctc_decoder_class()
    sd = ho.random_head_state_dict(64, 40, 9)
    dec = cls(feat_in=64, num_classes=40)
    dec.load_state_dict(sd, strict=True)
    x = torch.randn(2, 64, 19, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        want = dec(encoder_output=x)
    assert float((ho.ctc_head_forward(sd, x) - want).abs().max()) == 0.0


def _collapse_golden():
    z = np.load(os.path.join(GOLDEN, "ctc_collapse.npz"))
    for n in range(int(z["n_cases"])):
        k = f"c{n}"
        lens = z[k + "_lens"]
        want = [z[k + "_tokens"][i, :c].tolist() for i, c in enumerate(z[k + "_count"])]
        yield torch.from_numpy(z[k + "_pred"]), (torch.from_numpy(lens) if lens.size else None), int(z[k + "_blank"]), \
            bool(z[k + "_fold"]), want


def test_greedy_collapse_matches_reference_golden():
    """The collapse pinned to the UNMODIFIED reference (WER.ctc_decoder_predictions_tensor, metrics/wer.py:122-188; vectors
    from tests/golden/make_golden_ctc_collapse.py): the oracle restatement and the product's host path."""
    n = 0
    for pred, lens, blank, fold, want in _collapse_golden():
        if not fold:
            # fold_consecutive=False (wer.py:165-169) keeps every non-blank frame: no recipe selects it and the product
            # API has no such switch; the vectors are there to show the generator exercises the reference's other branch
            assert want == [[p for p in (row[: int(lens[i])] if lens is not None else row) if p != blank]
                            for i, row in enumerate(pred.tolist())]
            continue
        assert ho.greedy_collapse(pred, None if lens is None else lens.tolist(), blank) == want
        assert cn.ctc_greedy_decode(pred, lens, blank) == want
        n += 1
    assert n >= 10


def test_greedy_collapse_matches_live_reference_when_mounted():
    from oracle import reference_loader as rl

    if not rl.reference_available():
        pytest.skip("reference tree not mounted")
    g = torch.Generator().manual_seed(5)
    for v in (3, 29, 129):
        pred = torch.randint(0, v + 1, (7, 61), generator=g)
        pred[torch.rand(7, 61, generator=g) < 0.5] = v
        lens = torch.randint(0, 62, (7,), generator=g)
        want = rl.reference_ctc_collapse(pred, lens, v)
        assert ho.greedy_collapse(pred, lens.tolist(), v) == want
        assert cn.ctc_greedy_decode(pred, lens, v) == want
        assert cn.ctc_greedy_decode(pred, None, v) == rl.reference_ctc_collapse(pred, None, v)


@pytest.mark.gpu
def test_gpu_collapse_kernel_matches_reference_golden():
    for pred, lens, blank, fold, want in _collapse_golden():
        if fold:
            assert cn.ctc_greedy_decode(pred.cuda(), None if lens is None else lens.cuda(), blank) == want


def test_greedy_collapse_cases():
    blank = 4
    pred = torch.tensor([[1, 1, 4, 1, 2, 2, 4, 4, 3, 3], [4, 4, 4, 4, 4, 4, 4, 4, 4, 4], [0, 0, 0, 1, 4, 0, 2, 2, 2, 2]])
    assert ho.greedy_collapse(pred, [10, 10, 10], blank) == [[1, 1, 2, 3], [], [0, 1, 0, 2]]
    assert ho.greedy_collapse(pred, [3, 0, 4], blank) == [[1], [], [0, 1]]
    assert cn.ctc_greedy_decode(pred, torch.tensor([3, 0, 4]), blank) == ho.greedy_collapse(pred, [3, 0, 4], blank)
    assert cn.ctc_greedy_decode(pred, None, blank) == ho.greedy_collapse(pred, None, blank)


def test_decoder_surface_matches_reference():
    dec = cn.ConvASRDecoder(feat_in=176, num_classes=28)
    assert {k: tuple(v.shape) for k, v in dec.state_dict().items()} == {
        "decoder_layers.0.weight": (29, 176, 1), "decoder_layers.0.bias": (29,)}
    with pytest.raises(ValueError):
        cn.ConvASRDecoder(feat_in=176, num_classes=-1)
    with pytest.raises(ValueError):
        cn.ConvASRDecoder(feat_in=176, num_classes=3, vocabulary=["a", "b"])
    dec = cn.ConvASRDecoder(feat_in=176, num_classes=-1, vocabulary=list("abc"))
    assert dec.num_classes_with_blank == 4 and dec.vocabulary == ["a", "b", "c"]
    with pytest.raises(RuntimeError):  # no CPU path
        dec(encoder_output=torch.zeros(1, 176, 5))


def assert_argmax_agrees(pred, want):
    """Greedy agreement >= 99 % (BASELINE.json north_star); a disagreement is only tolerated where the reference's own
    margin between the two classes is inside the bf16 tolerance (random-init heads have many near-ties)."""
    ref = want.argmax(-1)
    margin = want.gather(-1, ref.unsqueeze(-1)).squeeze(-1) - want.gather(-1, pred.unsqueeze(-1)).squeeze(-1)
    assert float(margin.max()) <= 0.1, float(margin.max())
    assert float((pred == ref).float().mean()) >= (0.99 if pred.numel() >= 1000 else 0.95)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_head_matches_golden(name):
    x, v, sd, want = load(name)
    dec = cn.ConvASRDecoder(feat_in=x.shape[1], num_classes=v)
    dec.load_state_dict(sd)
    dec = dec.cuda()
    lp, pred = dec.forward_with_predictions(x.cuda())
    torch.cuda.synchronize()
    assert lp.shape == want.shape and pred.shape == want.shape[:2]
    assert torch.isfinite(lp).all()
    assert float((lp.cpu() - want).abs().max()) <= 5e-2              # bf16 operands, fp32 accumulation + log_softmax
    assert float((lp.exp().sum(-1) - 1).abs().max()) <= 1e-4
    assert torch.equal(pred, lp.argmax(-1))                            # the kernel's argmax == torch's on its own output
    assert_argmax_agrees(pred.cpu(), want)


@pytest.mark.gpu
@pytest.mark.parametrize("b,t,d,v,dtype", [(32, 500, 512, 1024, torch.float32), (4, 333, 256, 128, torch.bfloat16),
                                          (1, 1, 176, 28, torch.float32), (3, 77, 512, 5000, torch.float32)])
def test_cuda_head_matches_oracle_on_encoder_like_input(b, t, d, v, dtype):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(b, t, d, generator=g)                             # contiguous (B, T, D): what the encoder hands over
    sd = ho.random_head_state_dict(d, v, 11)
    dec = cn.ConvASRDecoder(feat_in=d, num_classes=v)
    dec.load_state_dict(sd)
    dec = dec.cuda()
    xin = x.to(dtype).cuda().transpose(1, 2)
    lp, pred = dec.forward_with_predictions(xin)
    want = ho.ctc_head_forward(sd, x.to(dtype).float().transpose(1, 2))
    torch.cuda.synchronize()
    err = (lp.cpu() - want).abs()
    assert float(err.max()) <= 5e-2, float(err.max())
    assert_argmax_agrees(pred.cpu(), want)
    lens = [t] * b
    got = cn.ctc_greedy_decode(pred, lens, v)
    ref = ho.greedy_collapse(lp.argmax(-1).cpu(), lens, v)
    assert got == ref


@pytest.mark.gpu
@pytest.mark.parametrize("b,t,v,seed", [(1, 1, 4, 0), (3, 37, 5, 1), (32, 500, 1024, 2), (7, 700, 28, 3), (130, 257, 3, 4)])
def test_gpu_collapse_kernel_equals_the_host_loop(b, t, v, seed):
    """cfb_op_ctc_collapse against the reference's collapse loop (metrics/wer.py:152-164 as restated in the oracle): random
    class streams with long blank / repeat runs, ragged lengths incl. 0, more frames than one 256-frame block."""
    g = torch.Generator().manual_seed(seed)
    pred = torch.randint(0, v + 1, (b, t), generator=g)
    runs = torch.rand(b, t, generator=g) < 0.6               # make repeats and blank runs common
    for i in range(1, t):
        pred[:, i] = torch.where(runs[:, i], pred[:, i - 1], pred[:, i])
    lens = torch.randint(0, t + 1, (b,), generator=g)
    lens[0] = t
    tokens, n = cn.ctc_collapse_device(pred.cuda(), lens.cuda(), v)
    want = ho.greedy_collapse(pred, lens.tolist(), v)
    n = n.cpu().tolist()
    assert n == [len(w) for w in want]
    host = tokens.cpu()
    for i, w in enumerate(want):
        assert host[i, :n[i]].tolist() == w
    assert cn.ctc_greedy_decode(pred.cuda(), lens, v) == want
    assert cn.ctc_greedy_decode(pred.cuda(), None, v) == ho.greedy_collapse(pred, None, v)


@pytest.mark.gpu
def test_gpu_features_to_token_ids_on_the_device():
    d, v, b, t = 256, 128, 4, 90
    sd = ho.random_head_state_dict(d, v, 21)
    dec = cn.ConvASRDecoder(feat_in=d, num_classes=v)
    dec.load_state_dict(sd)
    dec = dec.cuda()
    x = torch.randn(b, d, t, generator=torch.Generator().manual_seed(8)).cuda()
    lens = torch.tensor([90, 61, 7, 0]).cuda()
    tokens, n = dec.greedy_tokens(x, lens)
    _, pred = dec.forward_with_predictions(x)
    want = ho.greedy_collapse(pred.cpu(), lens.tolist(), v)
    assert tokens.dtype == torch.int32 and tokens.is_cuda and n.cpu().tolist() == [len(w) for w in want]
    for i, w in enumerate(want):
        assert tokens[i, :len(w)].cpu().tolist() == w
