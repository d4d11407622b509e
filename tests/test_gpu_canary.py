"""GPU: out-of-bounds writes of the forward, checked with guard bands of our own.

compute-sanitizer is closed on this GPU pool (profiles/r5s_compute_sanitizer_closed.log), so the memcheck the survey asks
for is done the manual way: every buffer the C ABI writes -- workspace, encoded, encoded_len -- sits
between two guard bands filled with a byte pattern, the forward runs (dense, packed, packed in 1..4 groups, odd shapes),
and the bands must come back untouched.  The workspace is handed over with EXACTLY the byte count cfb_workspace_bytes /
cfb_packed_workspace_bytes reports, so a kernel that writes past its carve-out lands in the band behind it.
"""
import ctypes

import pytest
import torch

import conformer_nemo_b200 as cn
from conformer_nemo_b200 import _lib
from oracle import conformer_oracle as oc

pytestmark = pytest.mark.gpu
GUARD = 1 << 20
PATTERN = 0xA5


class Guarded:
    """`nbytes` of device memory, 256-byte aligned, between two GUARD-byte bands of PATTERN."""

    def __init__(self, nbytes):
        self.nbytes = nbytes
        self.raw = torch.full((GUARD + nbytes + 256 + GUARD,), PATTERN, dtype=torch.uint8, device="cuda")
        base = self.raw.data_ptr() + GUARD
        self.off = GUARD + ((-base) % 256)
        self.ptr = self.raw.data_ptr() + self.off

    def view(self, dtype, shape):
        n = int(torch.tensor([], dtype=dtype).element_size())
        count = 1
        for s in shape:
            count *= s
        return self.raw[self.off: self.off + count * n].view(dtype).view(*shape)

    def bands_intact(self):
        lo = self.raw[: self.off]
        hi = self.raw[self.off + self.nbytes:]
        return bool((lo == PATTERN).all()) and bool((hi == PATTERN).all())


def run_guarded(enc, x, lens, packed):
    lib = _lib.load_library()
    b, f, t = x.shape
    t_out = enc.output_frames(t)
    d = enc._feat_out
    xd = x.cuda().contiguous()
    ld = torch.tensor(lens, dtype=torch.int64, device="cuda")
    nbytes = ctypes.c_size_t()
    hl = (ctypes.c_int64 * b)(*lens)
    if packed:
        _lib.check(lib.cfb_packed_workspace_bytes(enc._handle, hl, b, t, ctypes.byref(nbytes)), enc._handle, "ws")
    else:
        _lib.check(lib.cfb_workspace_bytes(enc._handle, b, t, ctypes.byref(nbytes)), enc._handle, "ws")
    ws = Guarded(nbytes.value)
    out = Guarded(b * t_out * d * 4)
    olen = Guarded(b * 4)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    vp = ctypes.c_void_p
    if packed:
        rc = lib.cfb_forward_packed(enc._handle, vp(xd.data_ptr()), _lib.CFB_F32, vp(ld.data_ptr()), hl, b, t, vp(out.ptr),
                                    _lib.CFB_F32, vp(olen.ptr), vp(ws.ptr), nbytes.value, stream)
    else:
        rc = lib.cfb_forward(enc._handle, vp(xd.data_ptr()), _lib.CFB_F32, vp(ld.data_ptr()), b, t, vp(out.ptr), _lib.CFB_F32,
                             vp(olen.ptr), vp(ws.ptr), nbytes.value, stream)
    _lib.check(rc, enc._handle, "forward")
    torch.cuda.synchronize()
    assert ws.bands_intact(), "a kernel wrote outside the workspace"
    assert out.bands_intact(), "a kernel wrote outside `encoded`"
    assert olen.bands_intact(), "a kernel wrote outside `encoded_len`"
    return out.view(torch.float32, (b, t_out, d)).clone(), olen.view(torch.int32, (b,)).clone()


CASES = [  # (d_model, heads, lengths): tile-size tails in every dimension, a zero-length row, a single long row
    (256, 4, [400, 399, 37, 1]),
    (176, 4, [333, 130, 129, 128, 127]),
    (512, 8, [1031, 5]),
    (256, 4, [0, 260]),
    (512, 8, [2049]),
    (256, 4, [97] * 9),
]


@pytest.mark.parametrize("d,heads,lens", CASES)
def test_dense_and_packed_forward_stay_inside_their_buffers(d, heads, lens, monkeypatch):
    cfg = oc.EncoderConfig(feat_in=80, n_layers=2, d_model=d, n_heads=heads)
    enc = cn.ConformerEncoder(feat_in=80, n_layers=2, d_model=d, n_heads=heads)
    enc.load_state_dict(oc.random_state_dict(cfg, 1), strict=False)
    enc = enc.cuda().eval()
    enc.prepare()
    t = max(max(lens), 8)
    x, _ = oc.synthetic_batch(len(lens), 80, t, lens, seed=3)
    yd, ld = run_guarded(enc, x, lens, packed=False)
    for groups in (1, 2, 4):
        monkeypatch.setenv("CFB_PACKED_GROUPS", str(groups))
        yp, lp = run_guarded(enc, x, lens, packed=True)
        assert torch.equal(ld, lp)
        for b in range(len(lens)):
            n = int(ld[b])
            assert torch.equal(yd[b, :n], yp[b, :n]), (groups, b)
