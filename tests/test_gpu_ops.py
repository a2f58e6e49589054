"""GPU kernels against the CPU oracle on seeded random inputs: every Mix op, the
Rgba->Gray average, export, HeightToNormal and all five resize filters, at
ragged sizes, in EXACT (bit-exact) and FAST (1e-5 rel / 1e-6 abs) modes."""
import numpy as np
import pytest

import kanter_core_b200 as kc
import oracle
from kanter_core_b200 import MixType, Node, NodeType, ResizeFilter, ResizePolicy, Size, SlotId
from kanter_core_b200._lib import TexProError

pytestmark = pytest.mark.gpu

REL, ABS = 1e-5, 1e-6  # BASELINE.json north_star tolerance for f32 arithmetic


def rnd(seed, h, w, lo=0.0, hi=1.0):
    r = np.random.default_rng(seed)
    return (r.random((h, w), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)


def bits_equal(a, b):
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    na, nb = np.isnan(a), np.isnan(b)
    return a.shape == b.shape and np.array_equal(na, nb) and np.array_equal(a[~na].view(np.uint32), b[~nb].view(np.uint32))


def close(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    fin = np.isfinite(a) & np.isfinite(b)
    if not np.array_equal(a[~fin & ~na], b[~fin & ~nb]):
        return False
    return bool((np.abs(a[fin] - b[fin]) <= ABS + REL * np.abs(b[fin])).all())


def mix_graph(tp, op, left, right):
    """Mix through the live graph with embedded inputs; returns the result image."""
    lg = tp.new_live_graph()
    m = lg.add_node(Node.new(NodeType.Mix(op)))
    for slot, planes in enumerate((left, right)):
        if planes is None:
            continue
        img = kc.SlotImage.from_planes(tp, planes)
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, img), slot)
        e = lg.add_node(Node.new(NodeType.Embed(slot)))
        lg.connect(e, m, SlotId(0), SlotId(slot))
    kc.LiveGraph.await_clean_read(lg, m)
    return lg.slot_data(m, SlotId(0)).image


SHAPES = [(1, 1), (3, 5), (17, 33), (64, 64), (255, 257)]


@pytest.mark.parametrize("op", list(MixType))
@pytest.mark.parametrize("shape", SHAPES)
def test_mix_gray_exact(tex_pro, op, shape):
    h, w = shape
    l, r = rnd(1, h, w, -1.0, 2.0), rnd(2, h, w, -0.5, 1.5)
    l.flat[0] = 0.0; r.flat[0] = 0.0  # 0/0, 0^0
    got = mix_graph(tex_pro, op, [l], [r]).planes()[0]
    assert bits_equal(got, oracle.mix_plane(int(op), l, r))


@pytest.mark.parametrize("op", list(MixType))
def test_mix_rgba_exact_and_alpha_is_one(tex_pro, op):
    h, w = 37, 41
    L = [rnd(10 + c, h, w) for c in range(4)]
    R = [rnd(20 + c, h, w) for c in range(4)]
    img = mix_graph(tex_pro, op, L, R)
    got = img.planes()
    for c in range(3):
        assert bits_equal(got[c], oracle.mix_plane(int(op), L[c], R[c]))
    assert np.array_equal(got[3], np.ones((h, w), np.float32))  # mix.rs:203-212
    assert img.plane_is_constant(3)[0] is False  # a requested node hands back real planes


@pytest.mark.parametrize("op", list(MixType))
def test_mix_fast_within_tolerance(tex_pro_fast, op):
    h, w = 128, 96
    l, r = rnd(3, h, w, 0.0, 1.0), rnd(4, h, w, 0.0, 1.0)
    l[0, :8] = [0.0, 1.0, 0.5, 2.0, 1e-20, 1e20, -1.0, 0.25]
    r[0, :8] = [0.0, 5.0, 0.0, 0.5, 0.5, 0.5, 2.0, -3.0]
    got = mix_graph(tex_pro_fast, op, [l], [r]).planes()[0]
    assert close(got, oracle.mix_plane(int(op), l, r))


def test_mix_type_coercion(tex_pro):
    # right is converted to left's type (mix.rs:62): Rgba->Gray is ((r+g)+b)/3, Gray->Rgba aliases
    h, w = 19, 23
    G = [rnd(5, h, w)]
    C = [rnd(30 + c, h, w) for c in range(4)]
    got = mix_graph(tex_pro, MixType.Add, G, C).planes()
    assert len(got) == 1
    assert bits_equal(got[0], oracle.mix_plane(0, G[0], oracle.rgb_to_gray(C[0], C[1], C[2])))
    got = mix_graph(tex_pro, MixType.Multiply, C, G).planes()
    assert len(got) == 4
    for c in range(3):
        assert bits_equal(got[c], oracle.mix_plane(2, C[c], G[0]))


def test_mix_missing_sides(tex_pro):
    h, w = 8, 9
    C = [rnd(40 + c, h, w) for c in range(4)]
    got = mix_graph(tex_pro, MixType.Subtract, None, C).planes()  # zeros - right
    for c in range(3):
        assert bits_equal(got[c], oracle.mix_plane(1, np.zeros((h, w), np.float32), C[c]))
    img = mix_graph(tex_pro, MixType.Add, None, None)  # neither: 1x1 Gray 0 (mix.rs:77-83)
    assert not img.is_rgba() and img.size() == Size(1, 1) and img.planes()[0][0, 0] == 0.0


@pytest.mark.parametrize("shape", [(1, 1), (2, 3), (5, 4), (64, 64), (33, 100), (100, 36)])
def test_height_to_normal_exact(tex_pro, shape):
    h, w = shape
    hgt = rnd(7, h, w)
    img = kc.SlotImage.from_planes(tex_pro, [hgt])
    res = kc.SlotImage(tex_pro._ctx, _h2n_direct(tex_pro, img))
    got = res.planes()
    want = oracle.height_to_normal(hgt)
    for c_ in range(3):
        assert bits_equal(got[c_], want[c_]), "channel %d" % c_
    assert np.array_equal(got[3], np.ones((h, w), np.float32))


@pytest.mark.parametrize("case", ["flat", "tiny_steps", "huge", "mixed_scales", "signed", "big_random", "smooth"])
def test_height_to_normal_exact_adversarial(tex_pro, case):
    """EXACT HeightToNormal = the oracle bit for bit on inputs that exercise both the written-out
    div/sqrt sequences (differences 0 or in [2^-20, 4]) and the guarded IEEE fallback (differences
    tiny, denormal, huge, infinite)."""
    r = np.random.default_rng(21)
    h, w = 96, 256                      # w % 4 == 0: the vector kernel
    if case == "flat":
        hgt = np.full((h, w), 0.375, np.float32)
        hgt[10:20, 30:60] = 0.5         # a plateau: zero differences inside, steps on its rim
    elif case == "tiny_steps":
        base = r.random((h, w), dtype=np.float32)
        step = np.float32(2.0) ** r.integers(-40, -15, (h, w)).astype(np.float32)   # straddles 2^-20
        hgt = (base + step).astype(np.float32)
        hgt[::3] = base[::3]
        hgt[5, :] = np.float32(1e-42)   # denormals
        hgt[6, :] = 0.0
    elif case == "huge":
        hgt = (r.random((h, w), dtype=np.float32) * np.float32(1e30)).astype(np.float32)
        hgt[7, 9] = np.inf
        hgt[50, 100] = np.float32(3e38)
    elif case == "mixed_scales":
        hgt = (r.random((h, w), dtype=np.float32) * (np.float32(2.0) ** r.integers(-30, 6, (h, w)).astype(np.float32))).astype(np.float32)
    elif case == "signed":
        hgt = (r.random((h, w), dtype=np.float32) * 8 - 4).astype(np.float32)         # differences up to 8: both paths
    elif case == "big_random":
        h, w = 512, 1024
        hgt = r.random((h, w), dtype=np.float32)
    else:
        yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
        hgt = (0.5 + 0.25 * np.sin(xx / 17.0) * np.cos(yy / 11.0)).astype(np.float32)
    img = kc.SlotImage.from_planes(tex_pro, [hgt])
    with np.errstate(all="ignore"):
        want = oracle.height_to_normal(hgt)
    got = kc.SlotImage(tex_pro._ctx, _h2n_direct(tex_pro, img)).planes()
    for c_ in range(3):
        assert bits_equal(got[c_], want[c_]), "channel %d: %d samples differ" % (
            c_, int((got[c_].view(np.uint32) != want[c_].view(np.uint32)).sum()))


def _h2n_direct(tp, img):
    import ctypes as C
    from kanter_core_b200._lib import call, kc_image
    out = kc_image()
    call("kc_height_to_normal", tp._ctx._h, C.byref(img._im), C.byref(out))
    return out


def test_height_to_normal_fast_within_tolerance(tex_pro_fast):
    hgt = rnd(8, 96, 128)
    img = kc.SlotImage.from_planes(tex_pro_fast, [hgt])
    got = kc.SlotImage(tex_pro_fast._ctx, _h2n_direct(tex_pro_fast, img)).planes()
    want = oracle.height_to_normal(hgt)
    for c in range(3):
        assert close(got[c], want[c])


def _resize_direct(tp, img, w, h, filt):
    import ctypes as C
    from kanter_core_b200._lib import call, kc_image
    out = kc_image()
    call("kc_resize", tp._ctx._h, C.byref(img._im), w, h, int(filt), C.byref(out))
    return kc.SlotImage(tp._ctx, out)


RESIZES = [((110, 110), (128, 128)), ((64, 48), (200, 150)), ((256, 256), (100, 77)), ((7, 9), (8, 8)),
           ((1, 1), (16, 16)), ((5, 1), (3, 9)), ((31, 17), (31, 40))]


@pytest.mark.parametrize("filt", list(ResizeFilter))
@pytest.mark.parametrize("src,dst", RESIZES)
def test_resize_exact(tex_pro, filt, src, dst):
    (sw, sh), (dw, dh) = src, dst
    p = rnd(9, sh, sw, -0.25, 1.25)  # out-of-range values exercise the [0,1] clamp of the second pass
    img = kc.SlotImage.from_planes(tex_pro, [p])
    got = _resize_direct(tex_pro, img, dw, dh, filt).planes()[0]
    want = oracle.resize_plane(p, dw, dh, int(filt))
    assert bits_equal(got, want)


@pytest.mark.parametrize("filt", list(ResizeFilter))
def test_resize_fast_within_tolerance(tex_pro_fast, filt):
    p = rnd(11, 40, 50)
    img = kc.SlotImage.from_planes(tex_pro_fast, [p])
    got = _resize_direct(tex_pro_fast, img, 123, 99, filt).planes()[0]
    assert close(got, oracle.resize_plane(p, 123, 99, int(filt)))


@pytest.mark.parametrize("filt", list(ResizeFilter))
@pytest.mark.parametrize("src,dst", [((96, 80), (768, 640)), ((150, 40), (1100, 90)), ((300, 200), (640, 333))])
def test_resize_exact_multi_strip(tex_pro, filt, src, dst):
    """Sizes that span several column strips and full + ragged row groups of the fused kernel."""
    (sw, sh), (dw, dh) = src, dst
    p = rnd(13, sh, sw, -0.25, 1.25)
    img = kc.SlotImage.from_planes(tex_pro, [p])
    got = _resize_direct(tex_pro, img, dw, dh, filt).planes()[0]
    assert bits_equal(got, oracle.resize_plane(p, dw, dh, int(filt)))


def _set_resize_knobs(**kw):
    from kanter_core_b200._lib import call
    for k in ("resize_tma", "resize_g", "resize_rc", "resize_minb", "resize_store"):
        call("kc_debug_set_tuning", k.encode(), int(kw.get(k, 0)))


@pytest.mark.parametrize("knobs", [dict(resize_tma=-1), dict(resize_store=-1, resize_g=8, resize_rc=4, resize_minb=6), dict(resize_store=-1, resize_g=8, resize_rc=4, resize_minb=8),
                                   dict(resize_store=-1, resize_g=8, resize_rc=8, resize_minb=6), dict(resize_store=-1, resize_g=8, resize_rc=8, resize_minb=8),
                                   dict(resize_store=-1, resize_g=16, resize_rc=4, resize_minb=4), dict(resize_store=-1, resize_g=16, resize_rc=4, resize_minb=6),
                                   dict(resize_store=-1, resize_g=16, resize_rc=8), dict(resize_store=-1, resize_g=16, resize_rc=16),
                                   dict(resize_g=8, resize_rc=8), dict(resize_g=8, resize_rc=4, resize_minb=8), dict(resize_g=16, resize_rc=4),
                                   dict(resize_g=16, resize_rc=8), dict(resize_g=16, resize_rc=16), dict(resize_g=32, resize_rc=4), dict(resize_g=32, resize_rc=8),
                                   dict(resize_store=-1, resize_g=32, resize_rc=8)],
                         ids=lambda k: "-".join("%s%d" % (a.split("_")[1], b) for a, b in k.items()))
@pytest.mark.parametrize("filt", [ResizeFilter.Nearest, ResizeFilter.Triangle, ResizeFilter.CatmullRom, ResizeFilter.Gaussian, ResizeFilter.Lanczos3])
def test_resize_tensor_map_kernel_every_variant_bit_exact(tex_pro, filt, knobs):
    """The fused upsample kernel with tensor-map loads and stores (kc_resize_tma_kernel: every compiled combination of
    rows per group / rows per accumulator chunk / CTAs per SM) and the cp.async/STG kernel it replaces, bit for bit
    against the oracle: sizes with ragged right edges (dw % 512 != 0), ragged bottom groups, non-integer ratios
    (no shared windows), strips that start at odd rows, NaN / inf / out-of-range samples."""
    import ctypes as C
    from kanter_core_b200._lib import call, kc_image
    _set_resize_knobs(**knobs)
    try:
        for (sw, sh), (dw, dh) in (((96, 80), (768, 640)), ((300, 200), (640, 333)), ((128, 64), (1024 + 512 + 4, 515)), ((64, 40), (520, 47))):
            p = rnd(13 + sw, sh, sw, -0.25, 1.25)
            p[sh // 2, sw // 3] = np.nan
            p[sh // 3, sw // 2] = np.inf
            p[1, 1] = -np.inf
            img = kc.SlotImage.from_planes(tex_pro, [p])
            want = oracle.resize_plane(p, dw, dh, int(filt))
            got = _resize_direct(tex_pro, img, dw, dh, filt).planes()[0]
            assert bits_equal(got, want), ((sw, sh), (dw, dh))
            r0, nr = 7, dh - 18                  # a strip that starts and ends off the group grid
            out = kc_image()
            call("kc_resize_rows", tex_pro._ctx._h, C.byref(img._im), dw, dh, int(filt), r0, nr, C.byref(out))
            assert bits_equal(kc.SlotImage(tex_pro._ctx, out).planes()[0], want[r0:r0 + nr]), ("strip", (sw, sh), (dw, dh))
    finally:
        _set_resize_knobs()


@pytest.mark.parametrize("mode", ["exact", "fast"])
@pytest.mark.parametrize("filt", [ResizeFilter.Triangle, ResizeFilter.Lanczos3, ResizeFilter.Gaussian])
def test_resize_non_finite_values_propagate_like_the_reference(tex_pro, tex_pro_fast, mode, filt):
    """NaN / inf / huge samples (what Divide leaves behind) must come out of the resize as the
    reference's clamp leaves them: NaN stays NaN, +inf -> 1, -inf -> 0; the saturating fast
    store path of the fused kernel may only be taken where no such value is in reach."""
    tp = tex_pro if mode == "exact" else tex_pro_fast
    p = rnd(17, 64, 96)
    p[3, 5] = np.nan
    p[20, 40] = np.inf
    p[21, 41] = -np.inf
    p[40, 70] = 3e38
    p[41, 70] = -3e38
    p[63, 95] = np.nan
    img = kc.SlotImage.from_planes(tp, [p])
    got = _resize_direct(tp, img, 768, 512, filt).planes()[0]
    want = oracle.resize_plane(p, 768, 512, int(filt))
    assert np.array_equal(np.isnan(got), np.isnan(want))
    if mode == "exact":
        assert bits_equal(got, want)
    else:
        ok = ~np.isnan(want)
        assert close(got[ok], want[ok])


@pytest.mark.parametrize("filt", [ResizeFilter.Nearest, ResizeFilter.Triangle, ResizeFilter.Lanczos3])
@pytest.mark.parametrize("src,dst,parts", [((64, 48), (512, 384), 3), ((100, 70), (333, 211), 4), ((256, 256), (100, 77), 2)])
def test_resize_row_strips_equal_whole_image(tex_pro, filt, src, dst, parts):
    """kc_resize_rows: the row strips a row-sharded resize hands to each GPU are the
    rows of the whole result, bit for bit, at strip boundaries that are not multiples of the
    kernel's 16-row groups; the last case takes the two-pass (downsampling) route."""
    import ctypes as C
    from kanter_core_b200._lib import call, kc_image
    (sw, sh), (dw, dh) = src, dst
    p = rnd(31, sh, sw, -0.25, 1.25)
    img = kc.SlotImage.from_planes(tex_pro, [p])
    whole = _resize_direct(tex_pro, img, dw, dh, filt).planes()[0]
    assert bits_equal(whole, oracle.resize_plane(p, dw, dh, int(filt)))
    bounds = [dh * k // parts for k in range(parts + 1)]
    bounds[1] += 5                      # ragged on purpose
    for r0, r1 in zip(bounds[:-1], bounds[1:]):
        out = kc_image()
        call("kc_resize_rows", tex_pro._ctx._h, C.byref(img._im), dw, dh, int(filt), r0, r1 - r0, C.byref(out))
        strip = kc.SlotImage(tex_pro._ctx, out).planes()[0]
        assert strip.shape == (r1 - r0, dw)
        assert bits_equal(strip, whole[r0:r1]), (r0, r1)


@pytest.mark.parametrize("filt", list(ResizeFilter))
@pytest.mark.parametrize("src,dst", [((512, 384), (64, 50)), ((1000, 700), (37, 91)), ((64, 640), (64, 80)),
                                     ((256, 2048), (300, 100)), ((36, 300), (8, 7))])
def test_resize_downsample_exact(tex_pro, filt, src, dst):
    """Long windows: the source-marching vertical pass (ring of eight output rows in registers)
    and the horizontal pass over its intermediate, bit for bit against the oracle; the last case
    has width % 4 != 0 and stays on the per-output kernels."""
    (sw, sh), (dw, dh) = src, dst
    p = rnd(19, sh, sw, -0.25, 1.25)
    if filt == ResizeFilter.Lanczos3:
        p[sh // 2, sw // 3] = np.nan          # a NaN sample must reach exactly the outputs whose windows hold it
        p[sh // 3, sw // 2] = np.inf
    img = kc.SlotImage.from_planes(tex_pro, [p])
    got = _resize_direct(tex_pro, img, dw, dh, filt).planes()[0]
    assert bits_equal(got, oracle.resize_plane(p, dw, dh, int(filt)))


@pytest.mark.parametrize("mode", ["exact", "fast"])
@pytest.mark.parametrize("filt", [ResizeFilter.Triangle, ResizeFilter.Gaussian, ResizeFilter.Lanczos3])
def test_resize_downsample_signed_zeros_and_non_finite_rows(tex_pro, tex_pro_fast, mode, filt):
    """The marching vertical pass multiplies rows that are no tap of an open output by a weight of 0 instead of skipping
    them: that must leave no trace -- bands of -0.0, of negative denormals and of values whose products underflow, and
    whole rows of inf / NaN just outside a window, come out exactly as the oracle's per-window sums (EXACT: bit for bit
    including the sign of zero and the NaN positions; FAST: same NaN positions, values within tolerance)."""
    tp = tex_pro if mode == "exact" else tex_pro_fast
    sw, sh, dw, dh = 512, 1024, 64, 96
    p = rnd(23, sh, sw, -0.25, 1.25)
    p[0:200] = -0.0
    p[200:330, ::2] = np.float32(-1e-44)
    p[200:330, 1::2] = np.float32(1e-30)
    p[400] = np.inf                         # whole rows: every column of the block takes the tested path there
    p[401] = -np.inf
    p[640, 100:300] = np.nan
    p[900:, 17] = np.nan
    img = kc.SlotImage.from_planes(tp, [p])
    got = _resize_direct(tp, img, dw, dh, filt).planes()[0]
    want = oracle.resize_plane(p, dw, dh, int(filt))
    assert np.array_equal(np.isnan(got), np.isnan(want))
    if mode == "exact":
        assert bits_equal(got, want)
    else:
        ok = ~np.isnan(want)
        assert close(got[ok], want[ok])


@pytest.mark.parametrize("filt", [ResizeFilter.Triangle, ResizeFilter.Lanczos3])
def test_resize_rgba_planes_in_one_launch(tex_pro, filt):
    """An RGBA image whose planes all need pixels goes through ONE launch of the tensor-map kernel (grid.z = plane): every
    plane bit-identical to the oracle's per-plane resize, a Gray->Rgba image (three aliases of one plane + a constant alpha)
    still resizes its plane once, and kc_resize_rows of an RGBA image equals the rows of the whole result."""
    import ctypes as C
    from kanter_core_b200._lib import call, kc_image
    sw, sh, dw, dh = 96, 80, 768, 640
    planes = [rnd(70 + c, sh, sw, -0.25, 1.25) for c in range(4)]
    img = kc.SlotImage.from_planes(tex_pro, planes)
    k0 = tex_pro.stats()["kernel_launches"]
    got = kc.resize(tex_pro, img, Size(dw, dh), filt)
    assert tex_pro.stats()["kernel_launches"] - k0 == 1
    gp = got.planes()
    want = [oracle.resize_plane(p, dw, dh, int(filt)) for p in planes]
    for c in range(4):
        assert bits_equal(gp[c], want[c]), c
    out = kc_image()
    call("kc_resize_rows", tex_pro._ctx._h, C.byref(img._im), dw, dh, int(filt), 37, 300, C.byref(out))
    sp = kc.SlotImage(tex_pro._ctx, out).planes()
    for c in range(4):
        assert bits_equal(sp[c], want[c][37:337]), c
    # lazy planes are forced by one fused launch first, then resized by one more
    lazy = kc.mix(tex_pro, MixType.Add, img, img)
    k0 = tex_pro.stats()["kernel_launches"]
    lp = kc.resize(tex_pro, lazy, Size(dw, dh), filt).planes()
    assert tex_pro.stats()["kernel_launches"] - k0 == 2
    for c in range(3):
        assert bits_equal(lp[c], oracle.resize_plane(oracle.mix_plane(0, planes[c], planes[c]), dw, dh, int(filt))), c
    gray = kc.SlotImage.from_planes(tex_pro, [planes[0]]).as_type(True)
    gg = kc.resize(tex_pro, gray, Size(dw, dh), filt)
    assert gg.same_plane(0, gg, 1) and gg.same_plane(0, gg, 2)
    assert bits_equal(gg.planes()[0], want[0])


def test_resize_rgba_with_constant_alpha(tex_pro):
    # a constant alpha plane goes through the same tap arithmetic as any other plane
    planes = [rnd(50 + c, 20, 30) for c in range(3)]
    img = kc.SlotImage.from_planes(tex_pro, planes + [np.ones((20, 30), np.float32)])
    lg = tex_pro.new_live_graph()
    lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, img), 0)
    e = lg.add_node(Node.new(NodeType.Embed(0)))
    n = Node.new(NodeType.Mix(MixType.Add))
    n.resize_policy = ResizePolicy.SpecificSize(Size(77, 45))
    n.resize_filter = ResizeFilter.Lanczos3
    m = lg.add_node(n)
    lg.connect(e, m, SlotId(0), SlotId(0))
    kc.LiveGraph.await_clean_read(lg, m)
    got = lg.slot_data(m, SlotId(0)).image.planes()
    for c in range(3):
        r = oracle.resize_plane(planes[c], 77, 45, 4)
        assert bits_equal(got[c], oracle.mix_plane(0, r, np.zeros_like(r)))


@pytest.mark.parametrize("srgb", [False, True])
@pytest.mark.parametrize("shape", [(1, 1), (3, 3), (31, 33), (64, 64)])
def test_to_u8(tex_pro, srgb, shape):
    h, w = shape
    P = [rnd(60 + c, h, w, -0.5, 1.5) for c in range(4)]
    P[0].flat[0] = np.nan; P[1].flat[0] = np.inf; P[2].flat[0] = -np.inf; P[3].flat[0] = -0.0
    img = kc.SlotImage.from_planes(tex_pro, P)
    assert np.array_equal(img.to_u8(srgb), oracle.to_u8(P, srgb))
    g = kc.SlotImage.from_planes(tex_pro, P[:1])
    assert np.array_equal(g.to_u8(srgb), oracle.to_u8(P[:1], srgb))


@pytest.mark.parametrize("srgb", [False, True])
def test_to_u8_fast_within_one_lsb(tex_pro_fast, srgb):
    # FAST export: the plain bytes are identical, the sRGB curve (SFU lg2/ex2, reciprocal multiplies) within 1 LSB;
    # every 8-bit level and both sides of the curve's knee (0.04045) are in the input
    h, w = 96, 128
    P = [rnd(70 + c, h, w, -0.25, 1.25) for c in range(4)]
    P[0].flat[:256] = np.arange(256, dtype=np.float32) / np.float32(255.0)
    P[1].flat[:6] = [0.04045, np.nextafter(np.float32(0.04045), np.float32(1)), 0.0404, 0.0, 1.0, np.nan]
    img = kc.SlotImage.from_planes(tex_pro_fast, P)
    got, want = img.to_u8(srgb).astype(np.int32), oracle.to_u8(P, srgb).astype(np.int32)
    if srgb:
        assert np.abs(got - want).max() <= 1
        assert (got != want).mean() < 0.01
    else:
        assert np.array_equal(got, want)
    g = kc.SlotImage.from_planes(tex_pro_fast, P[:1])
    assert np.abs(g.to_u8(srgb).astype(np.int32) - oracle.to_u8(P[:1], srgb).astype(np.int32)).max() <= (1 if srgb else 0)


@pytest.mark.parametrize("ch", [1, 2, 3, 4])
@pytest.mark.parametrize("shape", [(1, 1), (5, 7), (64, 64), (33, 31)])
def test_from_u8(tex_pro, ch, shape):
    h, w = shape
    a = np.random.default_rng(70).integers(0, 256, (h, w, ch), dtype=np.uint8)
    got = kc.SlotImage.from_u8(tex_pro, a).planes()
    want = oracle.deconstruct_u8(a)
    for c in range(4):
        assert np.array_equal(got[c], want[c])


def test_u8_roundtrip_every_value(tex_pro):
    a = np.arange(256, dtype=np.uint8).reshape(16, 16, 1).repeat(4, axis=2)
    assert np.array_equal(kc.SlotImage.from_u8(tex_pro, a).to_u8(), a)


def test_long_chain_fuses_and_matches(tex_pro):
    """A 40-node chain of Mix nodes is cut into a few fused kernels (not 40) and
    still matches the node-by-node oracle bit for bit."""
    h, w = 50, 70
    A, B = rnd(80, h, w, 0.5, 1.5), rnd(81, h, w, 0.5, 1.5)
    lg = tex_pro.new_live_graph()
    for i, p in enumerate((A, B)):
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tex_pro, [p])), i)
    a = lg.add_node(Node.new(NodeType.Embed(0)))
    b = lg.add_node(Node.new(NodeType.Embed(1)))
    ops = [MixType.Add, MixType.Multiply, MixType.Subtract, MixType.Divide]
    cur, want = a, A
    for i in range(40):
        m = lg.add_node(Node.new(NodeType.Mix(ops[i % 4])))
        lg.connect(cur, m, SlotId(0), SlotId(0))
        lg.connect(b, m, SlotId(0), SlotId(1))
        want = oracle.mix_plane(int(ops[i % 4]), want, B)
        cur = m
    kc.LiveGraph.await_clean_read(lg, cur)
    st = lg.last_run_stats()
    assert 1 <= st["kernels"] <= 4, st
    assert bits_equal(lg.slot_data(cur, SlotId(0)).image.planes()[0], want)


def test_shared_subexpression_and_diamond(tex_pro):
    h, w = 33, 47
    A, B = rnd(90, h, w), rnd(91, h, w)
    lg = tex_pro.new_live_graph()
    for i, p in enumerate((A, B)):
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tex_pro, [p])), i)
    a = lg.add_node(Node.new(NodeType.Embed(0)))
    b = lg.add_node(Node.new(NodeType.Embed(1)))
    s = lg.add_node(Node.new(NodeType.Mix(MixType.Add)))       # s = A + B
    d = lg.add_node(Node.new(NodeType.Mix(MixType.Subtract)))  # d = A - B
    p = lg.add_node(Node.new(NodeType.Mix(MixType.Multiply)))  # p = s * d
    q = lg.add_node(Node.new(NodeType.Mix(MixType.Divide)))    # q = p / s   (s used twice)
    for (o, i, sl) in [(a, s, 0), (b, s, 1), (a, d, 0), (b, d, 1), (s, p, 0), (d, p, 1), (p, q, 0), (s, q, 1)]:
        lg.connect(o, i, SlotId(0), SlotId(sl))
    kc.LiveGraph.await_clean_read(lg, q)
    assert lg.last_run_stats()["kernels"] == 1
    S, D = oracle.mix_plane(0, A, B), oracle.mix_plane(1, A, B)
    want = oracle.mix_plane(3, oracle.mix_plane(2, S, D), S)
    assert bits_equal(lg.slot_data(q, SlotId(0)).image.planes()[0], want)


@pytest.mark.parametrize("h,w,parts", [(64, 64, 2), (50, 36, 3), (33, 100, 4), (16, 12, 8)])
def test_height_to_normal_strips_equal_whole_image(tex_pro, h, w, parts):
    """The multi-GPU tiling emulated on one GPU: each strip, fed the halo row
    from the strip above by a device-to-device row copy, gives exactly its rows of the whole."""
    from kanter_core_b200 import dist as kdist
    hgt = rnd(21, h, w)
    whole = kc.height_to_normal(tex_pro, kc.SlotImage.from_planes(tex_pro, [hgt])).planes()
    want = oracle.height_to_normal(hgt)
    for c in range(3):
        assert bits_equal(whole[c], want[c])
    strips = [kdist.strip_rows(h, r, parts) for r in range(parts)]
    imgs = [kc.SlotImage.from_planes(tex_pro, [hgt[y0:y1]]) for (y0, y1) in strips]
    for r, (y0, y1) in enumerate(strips):
        above = imgs[(r - 1) % parts]
        halo = kc.empty_gray(tex_pro, w, 1)
        kc.copy_rows(tex_pro, halo, 0, above, above.size().height - 1, 1)
        got = kc.height_to_normal_strip(tex_pro, imgs[r], halo, h).planes()
        for c in range(3):
            assert bits_equal(got[c], want[c][y0:y1]), (r, c)


@pytest.mark.parametrize("h,w,parts", [(64, 64, 2), (100, 128, 3), (256, 512, 4)])
def test_height_to_normal_strips_with_peer_mailboxes(tex_pro, h, w, parts):
    """The halo row read by the kernel from a mailbox (kc_halo_*), every "rank" emulated in one
    process on one stream: publish always precedes the reading kernel in stream order, so
    nothing ever waits (what the multi-process run does across GPUs is the same data path through
    a CUDA-IPC mapping).  Three steps with different data exercise both slots and the ack."""
    from kanter_core_b200 import dist as kdist
    strips = [kdist.strip_rows(h, r, parts) for r in range(parts)]
    boxes = [kc.HaloLink.outbox(tex_pro, w) for _ in range(parts)]
    inboxes = [boxes[(r - 1) % parts].local_inbox() for r in range(parts)]
    for step in (1, 2, 3):
        hgt = rnd(40 + step, h, w)
        want = oracle.height_to_normal(hgt)
        imgs = [kc.SlotImage.from_planes(tex_pro, [hgt[y0:y1]]) for (y0, y1) in strips]
        for r, (y0, y1) in enumerate(strips):
            boxes[r].publish(imgs[r], (y1 - y0) - 1, step)
        for r, (y0, y1) in enumerate(strips):
            got = kc.height_to_normal_strip_peer(tex_pro, imgs[r], inboxes[r], step, h).planes()
            for c in range(3):
                assert bits_equal(got[c], want[c][y0:y1]), (step, r, c)
    assert kc.halo_timeouts(tex_pro) == 0
    for l in inboxes + boxes:
        l.close()


@pytest.mark.parametrize("h,w,parts", [(64, 64, 1), (64, 128, 2), (100, 256, 3), (512, 1024, 4), (4096, 4096, 2)])
def test_height_to_normal_strip_exchange_in_one_launch(h, w, parts):
    """kc_height_to_normal_strip_exchange: publish + stencil + acknowledge in ONE kernel per strip and step.  Every
    "rank" is a context of its own (its own stream) in this process, all launched back to back: strip r waits inside
    its kernel for the row strip r-1 publishes from inside ITS kernel -- rank 0 for the last rank's, launched after it
    (the wrap) -- so the kernels really overlap and the flags really order them.  parts = 1: a GPU reading its own
    mailbox, which only works because the publishing blocks are scheduled first.  Four steps: both slots, the acks."""
    from kanter_core_b200 import dist as kdist
    tps = [kc.TextureProcessor.new() for _ in range(parts)]
    try:
        strips = [kdist.strip_rows(h, r, parts) for r in range(parts)]
        boxes = [kc.HaloLink.outbox(tps[r], w) for r in range(parts)]
        inboxes = [boxes[(r - 1) % parts].local_inbox(tps[r]) for r in range(parts)]    # same process: the mailbox is aliased, not IPC-mapped
        for step in (1, 2, 3, 4):
            hgt = rnd(60 + step, h, w)
            want = oracle.height_to_normal(hgt)
            imgs = [kc.SlotImage.from_planes(tps[r], [hgt[y0:y1]]) for r, (y0, y1) in enumerate(strips)]
            for tp in tps:
                tp.synchronize()
            outs = [kc.height_to_normal_strip_exchange(tps[r], imgs[r], boxes[r], inboxes[r], step, h) for r in range(parts)]
            for r, (y0, y1) in enumerate(strips):
                got = outs[r].planes()
                for c in range(3):
                    assert bits_equal(got[c], want[c][y0:y1]), (step, r, c)
        for tp in tps:
            assert kc.halo_timeouts(tp) == 0
        for l in inboxes + boxes:
            l.close()
    finally:
        for tp in tps:
            tp.close()


def test_a_halo_row_that_never_arrives_is_an_error_not_a_stale_result():
    """A peer that never publishes: the kernel's wait gives up after 2 s (never a hang) and the NEXT synchronising
    call fails -- the strip was computed from whatever the mailbox held (advisor finding, round 1)."""
    tp = kc.TextureProcessor.new()
    try:
        box = kc.HaloLink.outbox(tp, 64)
        inbox = box.local_inbox()
        img = kc.SlotImage.from_planes(tp, [rnd(77, 16, 64)])
        out = kc.height_to_normal_strip_peer(tp, img, inbox, 1, 32)      # nobody published step 1
        with pytest.raises(TexProError):
            tp.synchronize()
        tp.synchronize()                                                 # reported once; the context stays usable
        box.publish(img, 15, 2)
        kc.height_to_normal_strip_peer(tp, img, inbox, 2, 32).planes()
        assert kc.halo_timeouts(tp) == 1
        del out
        inbox.close()
        box.close()
    finally:
        tp.close()


def test_ops_module_functions(tex_pro):
    A, B = rnd(31, 20, 24), rnd(32, 20, 24)
    ia, ib = kc.SlotImage.from_planes(tex_pro, [A]), kc.SlotImage.from_planes(tex_pro, [B])
    assert bits_equal(kc.mix(tex_pro, MixType.Divide, ia, ib).planes()[0], oracle.mix_plane(3, A, B))
    assert bits_equal(kc.mix(tex_pro, MixType.Add, ia, None).planes()[0], oracle.mix_plane(0, A, np.zeros_like(A)))
    assert bits_equal(kc.resize(tex_pro, ia, Size(31, 45), ResizeFilter.CatmullRom).planes()[0], oracle.resize_plane(A, 31, 45, 2))


def test_pipelined_async_read_matches_sync(tex_pro):
    """read_rgba(sync=False) steps enqueued back to back (uploads on the upload stream,
    downloads on the download stream, device buffers recycled between steps) deliver the
    same bytes as the synchronous path, step by step, with different inputs every step."""
    import ctypes as C
    from kanter_core_b200._lib import call
    tp = tex_pro
    S, steps = 256, 6
    ins = [[kc.pinned_empty((S, S)) for _ in range(4)] for _ in range(2 * steps)]
    for i, planes in enumerate(ins):
        for c, p in enumerate(planes):
            p[...] = rnd(1000 + 10 * i + c, S, S)
    outs = [kc.pinned_empty((S, S, 4), np.uint8) for _ in range(steps)]
    lg = tp.new_live_graph()
    lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tp, ins[0])), 0)
    lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tp, ins[1])), 1)
    a = lg.add_node(Node.new(NodeType.Embed(0)))
    b = lg.add_node(Node.new(NodeType.Embed(1)))
    m = lg.add_node(Node.new(NodeType.Mix(MixType.Multiply)))
    o = lg.add_node(Node.new(NodeType.OutputRgba("out")))
    lg.connect(a, m, SlotId(0), SlotId(0))
    lg.connect(b, m, SlotId(0), SlotId(1))
    lg.connect(m, o, SlotId(0), SlotId(0))
    ev = C.c_void_p()
    call("kc_event_create", C.byref(ev))
    for i in range(steps):
        lg.replace_embedded(kc.SlotImage.from_planes(tp, ins[2 * i], sync=False), 0)
        lg.replace_embedded(kc.SlotImage.from_planes(tp, ins[2 * i + 1], sync=False), 1)
        lg.read_rgba(o, SlotId(0), Size(S, S), out=outs[i], sync=False)
    call("kc_event_record_download", tp._ctx._h, ev)
    call("kc_event_synchronize", ev)
    call("kc_event_destroy", ev)
    for i in range(steps):
        planes = [oracle.mix_plane(2, ins[2 * i][c], ins[2 * i + 1][c]) for c in range(3)] + [np.ones((S, S), np.float32)]
        assert np.array_equal(outs[i], oracle.to_u8(planes, False)), "step %d" % i
    for planes in ins:
        for p in planes:
            kc.free_pinned(p)
    for p in outs:
        kc.free_pinned(p)


def test_deferred_host_planes_upload_only_what_is_read(tex_pro):
    """from_planes(deferred=True): planes stay in pinned host memory until an evaluation reads them;
    Mix(rgba) never reads alpha, so 6 of 8 planes are uploaded, and the result is unchanged."""
    tp = tex_pro
    S = 128
    A = [kc.pinned_empty((S, S)) for _ in range(4)]
    B = [kc.pinned_empty((S, S)) for _ in range(4)]
    for i, p in enumerate(A + B):
        p[...] = rnd(500 + i, S, S)
    out = kc.pinned_empty((S, S, 4), np.uint8)
    lg = tp.new_live_graph()
    lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tp, A, deferred=True)), 0)
    lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tp, B, deferred=True)), 1)
    a = lg.add_node(Node.new(NodeType.Embed(0)))
    b = lg.add_node(Node.new(NodeType.Embed(1)))
    m = lg.add_node(Node.new(NodeType.Mix(MixType.Multiply)))
    o = lg.add_node(Node.new(NodeType.OutputRgba("out")))
    lg.connect(a, m, SlotId(0), SlotId(0))
    lg.connect(b, m, SlotId(0), SlotId(1))
    lg.connect(m, o, SlotId(0), SlotId(0))
    x0 = tp.transfer_stats()
    lg.read_rgba(o, SlotId(0), Size(S, S), out=out)
    x1 = tp.transfer_stats()
    assert x1["h2d_bytes"] - x0["h2d_bytes"] == 6 * S * S * 4
    assert x1["d2h_bytes"] - x0["d2h_bytes"] == S * S * 4
    planes = [oracle.mix_plane(2, A[c], B[c]) for c in range(3)] + [np.ones((S, S), np.float32)]
    assert np.array_equal(out, oracle.to_u8(planes, False))
    # reading a deferred plane directly uploads just that plane
    img = kc.SlotImage.from_planes(tp, A, deferred=True)
    assert np.array_equal(img.planes()[3], A[3])
    x2 = tp.transfer_stats()
    assert x2["h2d_bytes"] - x1["h2d_bytes"] == 4 * S * S * 4
    tp.synchronize()
    for p in A + B + [out]:
        kc.free_pinned(p)


@pytest.mark.parametrize("filt", [ResizeFilter.Lanczos3, ResizeFilter.CatmullRom, ResizeFilter.Triangle])
@pytest.mark.parametrize("sizes", [((40, 36), (131, 97)), ((257, 260), (64, 60)), ((1, 1), (32, 32))])
def test_resize_clamp_is_switchable(tex_pro, filt, sizes):
    """The [0,1] clamp of the second resize pass is the one arithmetic step of image-0.24 that no
    reference golden pins, so it can be switched off: the clamped result is then
    exactly clamp(unclamped), overshoot and out-of-range inputs survive, and the default is on."""
    (sh, sw), (dh, dw) = sizes
    src = rnd(123, sh, sw, -0.5, 1.5)
    if src.size > 4:
        src[::5, ::3] = 1.0
        src[1::5, 1::3] = 0.0          # hard edges: Lanczos3 / CatmullRom overshoot
    else:
        src[0, 0] = 1.3
    img = kc.SlotImage.from_planes(tex_pro, [src])
    clamped = kc.resize(tex_pro, img, Size(dw, dh), filt).planes()[0]
    assert bits_equal(clamped, oracle.resize_plane(src, dw, dh, int(filt)))
    tex_pro.set_resize_clamp(False)
    try:
        free = kc.resize(tex_pro, img, Size(dw, dh), filt).planes()[0]
        v = kc.SlotImage.from_value(tex_pro, Size(1, 1), 1.75, False)
        assert kc.resize(tex_pro, v, Size(8, 8), filt).planes()[0][3, 3] == np.float32(1.75)     # the folded broadcast too
    finally:
        tex_pro.set_resize_clamp(True)
    if dh >= sh:                         # an upsample keeps the input's range (and overshoots); a 4x downsample averages it away
        assert free.max() > 1.0 and (src.size == 1 or free.min() < 0.0)
    assert bits_equal(np.clip(free, np.float32(0), np.float32(1)), clamped)
    v = kc.SlotImage.from_value(tex_pro, Size(1, 1), 1.75, False)
    assert kc.resize(tex_pro, v, Size(8, 8), filt).planes()[0][3, 3] == np.float32(1.0)


@pytest.mark.parametrize("seed", range(6))
def test_resize_long_window_random_shapes_bit_exact(tex_pro, seed):
    """Seeded random downsampling shapes through the long-window kernels (TMA-fed vertical march with zero-weight idle slots,
    horizontal pass with two adjacent outputs per thread and block-wise weights): ratios from 1.3 to 40 per axis, independent
    per axis (two cases upsample along one axis and downsample along the other), widths that are and are not multiples of four
    (the latter take the per-output fallbacks), every filter, NaN and inf samples -- bit for bit against the oracle."""
    rng = np.random.default_rng(4000 + seed)
    for case in range(7):
        dw, dh = int(rng.integers(3, 90)), int(rng.integers(3, 70))
        rx, ry = float(rng.uniform(1.3, 12.0)), float(rng.uniform(1.3, 12.0))
        if case == 5:
            rx, ry = float(rng.uniform(20, 40)), float(rng.uniform(2, 4))        # one very long horizontal window
        if case == 4:
            rx = float(rng.uniform(0.3, 0.9))                                     # upsampling along x, a long window along y
        if case == 6:
            ry = float(rng.uniform(0.3, 0.9))                                     # and the other way round
        sw, sh = max(2, int(dw * rx)), max(2, int(dh * ry))
        if case % 3 != 2:
            sw = (sw + 3) & ~3                                                    # the TMA-fed march needs whole float4 columns
        sw, sh = min(sw, 2600), min(sh, 1400)
        filt = list(ResizeFilter)[int(rng.integers(0, len(ResizeFilter)))]
        p = rnd(int(rng.integers(1, 1 << 30)), sh, sw, -0.25, 1.25)
        if case % 2 == 0:
            p[int(rng.integers(0, sh)), int(rng.integers(0, sw))] = np.nan
            p[int(rng.integers(0, sh)), int(rng.integers(0, sw))] = np.inf
            p[int(rng.integers(0, sh))] = -0.0
        img = kc.SlotImage.from_planes(tex_pro, [p])
        got = _resize_direct(tex_pro, img, dw, dh, filt).planes()[0]
        want = oracle.resize_plane(p, dw, dh, int(filt))
        assert bits_equal(got, want), (seed, case, (sw, sh), (dw, dh), filt)
