"""Pins the CPU oracle: every golden check of the reference's test-suite
(tests/integration_tests.rs) must reproduce BYTE-EXACTLY, plus
the known-answer values the reference asserts directly.  CPU only."""
import numpy as np
import pytest

import oracle
from tests import graphs


@pytest.mark.parametrize("name", sorted(graphs.GOLDEN_CASES))
def test_oracle_reproduces_reference_golden(name):
    case = graphs.GOLDEN_CASES[name]()
    og = graphs.run_oracle(case)
    got = og.buffer_rgba(int(case.node), 0)
    want = case.expected()
    assert got.shape == want.shape == (case.size[1], case.size[0], 4)
    assert np.array_equal(got, want), "%d bytes differ" % int((got != want).sum())


@pytest.mark.parametrize("name,policy,p1,p2,size", graphs.RESIZE_POLICY_CASES, ids=[c[0] for c in graphs.RESIZE_POLICY_CASES])
def test_oracle_resize_policy_sizes(name, policy, p1, p2, size):
    # tests/integration_tests.rs:894-949
    case = graphs.resize_policy_case(policy, p1, p2)
    og = graphs.run_oracle(case)
    planes = og.slot(int(case.node), 0)
    assert planes[0].shape == (size[1], size[0])


def test_oracle_read_dirty_read_known_answer():
    # Value(0.5) -> Combine.red reads back [127, 0, 0, 255]; tests/integration_tests.rs:1388-1437
    from kanter_core_b200 import Node, NodeGraph, NodeType, SlotId
    g = NodeGraph.new()
    v = g.add_node(Node.new(NodeType.Value(0.5)))
    c = g.add_node(Node.new(NodeType.CombineRgba))
    g.connect(v, c, SlotId(0), SlotId(0))
    og = graphs.run_oracle(graphs.Case(g, c))
    assert og.buffer_rgba(int(c), 0).reshape(-1).tolist() == [127, 0, 0, 255]


def test_oracle_special_value_census():
    # divide_node_gray has 51 868 NaN and 6 231 +inf pixels, all exported as 255
    case = graphs.mix_node_gray(graphs.MixType.Divide, "divide_node_gray.png")
    og = graphs.run_oracle(case)
    p = og.slot(int(case.node), 0)[0]
    assert int(np.isnan(p).sum()) == 51868
    assert int(np.isposinf(p).sum()) == 6231
    rgba = og.buffer_rgba(int(case.node), 0)
    assert (rgba[np.isnan(p)][:, :3] == 255).all()


def test_oracle_u8_roundtrip_is_identity():
    # u8/255 -> f32 -> clamp*255 trunc round-trips every u8 value (input_output, embedded_node_data)
    a = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(4, axis=2)
    planes = oracle.deconstruct_u8(a)
    assert np.array_equal(oracle.to_u8(planes), a)
    gray = oracle.deconstruct_u8(a[:, :, :1])
    assert np.array_equal(gray[1], np.zeros((16, 16), np.float32)) and np.array_equal(gray[3], np.ones((16, 16), np.float32))


def test_oracle_resize_weights_triangle_upsample():
    # Triangle 110 -> 128 (the filter irregular_sizes.png pins): windows of <= 3 taps that sum to 1
    left, count, w = oracle.resize_weights(110, 128, 1)
    assert count.max() <= 3 and count.min() >= 1
    assert np.allclose(w.sum(axis=1), 1.0, atol=1e-6)
    # 1x1 -> N is a single tap of weight exactly 1.0 for every filter (value_node.png pins Triangle)
    for f in range(5):
        l, c, ww = oracle.resize_weights(1, 256, f)
        assert (c == 1).all() and (ww[:, 0] == 1.0).all() and (l == 0).all()


@pytest.mark.parametrize("filt, pil_filter", [(1, "BILINEAR"), (2, "BICUBIC"), (4, "LANCZOS")])
@pytest.mark.parametrize("dst", [(120, 156), (20, 26), (57, 33), (40, 52)])
def test_oracle_resize_agrees_with_an_independent_implementation(filt, pil_filter, dst):
    """No golden of the reference pins CatmullRom / Lanczos3 pixels or Triangle downsampling,
    and the image crate itself cannot be built here.  Pillow's resampler is an independent implementation
    of the same separable design (window = support x max(ratio, 1) around the pixel centre, weights
    normalised by their sum, BICUBIC with a = -0.5 = CatmullRom): on f32 data kept inside [0, 1] the
    oracle agrees with it to float rounding, up-, down- and mixed-sampling.  This pins the arithmetic of
    the restated algorithm (windows, kernels, normalisation), not its bits."""
    import numpy as np
    from PIL import Image
    import oracle
    r = np.random.default_rng(5)
    k = np.outer(np.hanning(9), np.hanning(9)).astype(np.float64)
    noise = r.random((48, 60))
    smooth = np.zeros((40, 52))
    for dy in range(9):
        for dx in range(9):
            smooth += k[dy, dx] * noise[dy:dy + 40, dx:dx + 52]
    src = (0.2 + 0.6 * (smooth - smooth.min()) / (smooth.max() - smooth.min())).astype(np.float32)
    dh, dw = dst
    got = oracle.resize_plane(src, dw, dh, filt)
    want = np.asarray(Image.fromarray(src, mode="F").resize((dw, dh), getattr(Image, pil_filter)), dtype=np.float32)
    assert got.shape == want.shape
    assert float(np.abs(got - want).max()) < 2e-6


@pytest.mark.parametrize("filt", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("dst", [(90, 31), (13, 77), (40, 52)])
def test_oracle_resize_against_the_formula_in_float64(filt, dst):
    """The documented algorithm of image 0.24's sample.rs written out a second time, in
    numpy float64 with the centres and windows in float32 as the crate computes them: the C++ oracle
    (f32 accumulation) must agree to rounding for all five filters, Gaussian and Nearest included."""
    import numpy as np
    import oracle
    f32 = np.float32

    def kernel(x):
        a = abs(x)
        if filt == 0:
            return 1.0
        if filt == 1:
            return max(0.0, 1.0 - a)
        if filt == 2:   # bc_cubic, B = 0, C = 0.5
            if a < 1:
                return ((12 - 6 * 0.5) * a ** 3 + (-18 + 6 * 0.5) * a ** 2 + 6) / 6
            if a < 2:
                return ((-6 * 0.5) * a ** 3 + (30 * 0.5) * a ** 2 + (-48 * 0.5) * a + 24 * 0.5) / 6
            return 0.0
        if filt == 3:   # gaussian(x, r = 0.5), not truncated at the support
            return float(np.exp(-x * x / (2 * 0.25)) / (np.sqrt(2 * np.pi) * 0.5))
        if a >= 3:
            return 0.0
        s = lambda t: 1.0 if t == 0 else float(np.sin(np.pi * t) / (np.pi * t))
        return s(x) * s(x / 3)

    support = [0.0, 1.0, 2.0, 3.0, 3.0][filt]

    def axis(src_len, new_len):
        ratio = f32(src_len) / f32(new_len)
        sratio = max(ratio, f32(1.0))
        ssup = f32(support) * sratio
        out = []
        for o in range(new_len):
            c = (f32(o) + f32(0.5)) * ratio
            left = int(min(max(np.floor(c - ssup), 0), src_len - 1))
            right = int(min(max(np.ceil(c + ssup), left + 1), src_len))
            c = c - f32(0.5)
            w = np.array([kernel((f32(i) - c) / sratio) for i in range(left, right)], dtype=np.float64)
            out.append((left, w / w.sum()))
        return out

    r = np.random.default_rng(8)
    src = (0.1 + 0.8 * r.random((40, 52))).astype(np.float32)
    dh, dw = dst
    tmp = np.stack([sum(wk * src[l + k].astype(np.float64) for k, wk in enumerate(w)) for l, w in axis(40, dh)])
    want = np.stack([sum(wk * tmp[:, l + k] for k, wk in enumerate(w)) for l, w in axis(52, dw)], axis=1)
    want = np.clip(want, 0.0, 1.0)
    got = oracle.resize_plane(src, dw, dh, filt)
    assert got.shape == want.shape
    assert float(np.abs(got - want).max()) < 5e-6


def test_oracle_height_to_normal_against_the_formula_in_float64():
    """height_to_normal::process (src/node/height_to_normal.rs:16-77) written out in numpy float64:
    tangent (1/W, 0, h - left), bitangent (0, 1/H, up - h), both normalised, their cross product
    normalised, * 0.5 + 0.5, with the toroidal wrap of wrapping_sample_subtract.  The golden pins the
    oracle at 8-bit precision only; this pins its f32 planes to rounding."""
    import numpy as np
    import oracle
    r = np.random.default_rng(9)
    h, w = 37, 53
    hgt = r.random((h, w)).astype(np.float32)
    H = hgt.astype(np.float64)
    left = np.roll(H, 1, axis=1)
    up = np.roll(H, 1, axis=0)
    t = np.stack([np.full_like(H, 1.0 / w), np.zeros_like(H), H - left], axis=-1)
    b = np.stack([np.zeros_like(H), np.full_like(H, 1.0 / h), up - H], axis=-1)
    t /= np.linalg.norm(t, axis=-1, keepdims=True)
    b /= np.linalg.norm(b, axis=-1, keepdims=True)
    n = np.cross(t, b)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    want = n * 0.5 + 0.5
    got = oracle.height_to_normal(hgt)
    assert len(got) == 3                                       # R, G, B; the node adds the alpha of 1.0
    for c in range(3):
        assert float(np.abs(got[c].astype(np.float64) - want[..., c]).max()) < 2e-6, c


def test_oracle_mix_and_gray_average_are_plain_f32_arithmetic():
    """mix.rs:136-302 and SlotImage::as_type (slot_image.rs:242-253) are single IEEE f32 operations per
    sample: the oracle equals numpy's float32 arithmetic bit for bit (pow goes through libm's powf and is
    compared to rounding), including 0/0, x/0 and negative bases."""
    import numpy as np
    import oracle
    r = np.random.default_rng(10)
    a = (r.random((33, 47)) * 3 - 1).astype(np.float32)
    b = (r.random((33, 47)) * 3 - 1).astype(np.float32)
    a[0, :4] = [0.0, 1.0, -2.0, 0.0]
    b[0, :4] = [0.0, 0.0, 0.5, 3.0]
    with np.errstate(all="ignore"):
        for op, fn in ((0, np.add), (1, np.subtract), (2, np.multiply), (3, np.divide)):
            got, want = oracle.mix_plane(op, a, b), fn(a, b)
            assert np.array_equal(np.isnan(got), np.isnan(want))
            assert np.array_equal(got[~np.isnan(got)].view(np.uint32), want[~np.isnan(want)].view(np.uint32)), op
        got = oracle.mix_plane(4, a, b).astype(np.float64)
        want = np.power(a.astype(np.float64), b.astype(np.float64))
        fin = np.isfinite(want) & np.isfinite(got)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        assert (np.abs(got[fin] - want[fin]) <= 1.5e-7 * np.abs(want[fin]) + 1e-45).all()
        assert got[0, 0] == 1.0                                   # 0^0 = 1 (pinned by pow_node_*.png too)
    g = oracle.rgb_to_gray(a, b, a)
    assert np.array_equal(g.view(np.uint32), (((a + b) + a) / np.float32(3.0)).view(np.uint32))


def test_oracle_srgb_export_against_the_formula_in_float64():
    """SlotImage::to_u8_srgb (src/slot_image.rs:172-207) + srgb_to_linear (src/slot_data.rs:100-109): no test of the
    reference calls it, so the oracle's bytes are checked against the formula in float64 - equal except where
    the f32 result sits within rounding of a truncation boundary, and then one level apart."""
    import numpy as np
    import oracle
    r = np.random.default_rng(11)
    P = [(r.random((64, 80)) * 1.4 - 0.2).astype(np.float32) for _ in range(4)]
    P[0].flat[:256] = np.arange(256, dtype=np.float32) / np.float32(255.0)
    got = oracle.to_u8(P, True).astype(np.int32)
    want = np.empty_like(got)
    for c in range(4):
        v = np.clip(P[c].astype(np.float64), 0.0, 1.0)
        if c < 3:
            v = np.where(v <= 0.0, v, np.where(v <= np.float64(np.float32(0.04045)), v / np.float64(np.float32(12.92)),
                                               ((v + np.float64(np.float32(0.055))) / np.float64(np.float32(1.055))) ** 2.4))
        want[..., c] = np.minimum(v * 255.0, 255.0).astype(np.int32)
    assert np.abs(got - want).max() <= 1
    assert (got != want).mean() < 1e-3
    assert np.array_equal(got[..., 3], want[..., 3])             # alpha stays linear
