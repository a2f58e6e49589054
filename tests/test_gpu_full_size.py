"""Parity at BASELINE.json's full sizes (4096^2 RGBA, 8192^2 Gray, 1024^2 -> 8192^2) through
properties that do not need the CPU oracle to chew through 67 Mpixel -- exact commutativity and
identities of the Mix ops, toroidal shift-equivariance and unit length of HeightToNormal, pixel
replication of the Nearest upsample, partition of unity of the other filters, u8 round trips --
plus the oracle itself on row samples of the same results."""
import numpy as np
import pytest

import kanter_core_b200 as kc
import oracle
from kanter_core_b200 import MixType, ResizeFilter, Size

pytestmark = pytest.mark.gpu


def rnd(seed, h, w):
    return np.random.default_rng(seed).random((h, w), dtype=np.float32)


def test_mix_4096_rgba_identities_and_row_samples(tex_pro):
    S = 4096
    A = [rnd(1 + c, S, S) for c in range(4)]
    B = [rnd(5 + c, S, S) for c in range(4)]
    ia, ib = kc.SlotImage.from_planes(tex_pro, A), kc.SlotImage.from_planes(tex_pro, B)
    ab = kc.mix(tex_pro, MixType.Multiply, ia, ib).planes()
    ba = kc.mix(tex_pro, MixType.Multiply, ib, ia).planes()
    add_ab = kc.mix(tex_pro, MixType.Add, ia, ib).planes()
    add_ba = kc.mix(tex_pro, MixType.Add, ib, ia).planes()
    zero = kc.mix(tex_pro, MixType.Subtract, ia, ia).planes()
    one = kc.mix(tex_pro, MixType.Divide, ib, ib).planes()
    for c in range(3):
        assert np.array_equal(ab[c], ba[c]) and np.array_equal(add_ab[c], add_ba[c])     # IEEE * and + commute
        assert not zero[c].any()                                                          # x - x == +0
        nz = B[c] != 0
        assert np.array_equal(one[c][nz], np.ones(int(nz.sum()), np.float32))             # x / x == 1
        assert np.isnan(one[c][~nz]).all()                                                # 0 / 0 == NaN (-> 255 on export)
    for p in (ab, add_ab, zero, one):
        assert np.array_equal(p[3], np.ones((S, S), np.float32))                          # alpha := 1.0, mix.rs:203-212
    # configs[1] itself against the oracle on three bands of rows
    pw = kc.mix(tex_pro, MixType.Pow, kc.mix(tex_pro, MixType.Multiply, ia, ib), ib).planes()
    for rows in (slice(0, 16), slice(2040, 2056), slice(S - 16, S)):
        for c in range(3):
            want = oracle.mix_plane(4, oracle.mix_plane(2, A[c][rows], B[c][rows]), B[c][rows])
            assert np.array_equal(pw[c][rows].view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("fixture_name", ["tex_pro", "tex_pro_fast"])
def test_height_to_normal_8192_shift_equivariance_and_unit_length(fixture_name, request):
    tp = request.getfixturevalue(fixture_name)
    S = 8192
    H = rnd(3, S, S)
    n = kc.height_to_normal(tp, kc.SlotImage.from_planes(tp, [H])).planes()
    # the stencil wraps around (process_shared.rs:52-60): rolling the input rolls the output, bit for bit
    k, m = 1234, 77
    n2 = kc.height_to_normal(tp, kc.SlotImage.from_planes(tp, [np.roll(H, (k, m), axis=(0, 1))])).planes()
    for c in range(3):
        assert np.array_equal(np.roll(n[c], (k, m), axis=(0, 1)), n2[c])
    # a unit normal packed as n*0.5+0.5
    v = [(n[c][::64].astype(np.float64) * 2.0 - 1.0) for c in range(3)]
    assert np.abs(v[0] ** 2 + v[1] ** 2 + v[2] ** 2 - 1.0).max() < 1e-6
    assert (v[2] > 0).all()
    # and the oracle on two bands of rows (the first one needs the wrapped last row)
    want = oracle.height_to_normal_strip(H[0:8], S, H[S - 1])
    band = oracle.height_to_normal_strip(H[5000:5008], S, H[4999])
    for c in range(3):
        if fixture_name == "tex_pro":
            assert np.array_equal(n[c][0:8].view(np.uint32), want[c].view(np.uint32))
            assert np.array_equal(n[c][5000:5008].view(np.uint32), band[c].view(np.uint32))
        else:
            assert np.abs(n[c][0:8].astype(np.float64) - want[c]).max() <= 1e-6 + 1e-5
    assert np.array_equal(n[3], np.ones((S, S), np.float32))


def test_resize_1024_to_8192_properties(tex_pro):
    S, D = 1024, 8192
    L = rnd(4, S, S)
    img = kc.SlotImage.from_planes(tex_pro, [L])
    near = kc.resize(tex_pro, img, Size(D, D), ResizeFilter.Nearest).planes()[0]
    # Nearest, integer ratio: every source pixel replicated 8 x 8 (window [floor(c), ceil(c)) -> one tap of weight 1)
    assert np.array_equal(near, np.repeat(np.repeat(L, 8, axis=0), 8, axis=1))
    const = kc.SlotImage.from_planes(tex_pro, [np.full((S, S), 0.3125, np.float32)])
    for filt in (ResizeFilter.Triangle, ResizeFilter.CatmullRom, ResizeFilter.Gaussian, ResizeFilter.Lanczos3):
        out = kc.resize(tex_pro, const, Size(D, D), filt).planes()[0]
        assert np.abs(out.astype(np.float64) - 0.3125).max() < 1e-6, filt          # the normalised taps sum to one
        up = kc.resize(tex_pro, img, Size(D, D), filt).planes()[0]
        assert up.min() >= 0.0 and up.max() <= 1.0                                  # the second pass clamps
    # Lanczos3 rows against the oracle: the full 8192^2 result on the CPU takes a few seconds, do it once
    want = oracle.resize_plane(L, D, D, int(ResizeFilter.Lanczos3))
    got = kc.resize(tex_pro, img, Size(D, D), ResizeFilter.Lanczos3).planes()[0]
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_u8_round_trip_4096(tex_pro):
    S = 4096
    a = np.random.default_rng(9).integers(0, 256, (S, S, 4), dtype=np.uint8)
    img = kc.SlotImage.from_u8(tex_pro, a)
    assert np.array_equal(img.to_u8(False), a)          # trunc(min(clamp(b/255)*255, 255)) == b for every byte
