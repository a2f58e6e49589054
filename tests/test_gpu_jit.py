"""The NVRTC-specialised fused kernels (kc_jit.cu) against the same oracles as the interpreter:
with the policy forced to "always", goldens stay byte-exact, the 32-node graph bit-exact, long
chains / shared sub-expressions / RGBA8 export / FAST tolerance unchanged.  (`KC_JIT=1 pytest -m gpu`
runs the WHOLE suite through the specialised kernels; this file keeps a representative part of
that in the default run.)"""
import ctypes as C

import numpy as np
import pytest

import kanter_core_b200 as kc
import oracle
from kanter_core_b200 import MixType, Node, NodeType, SlotId
from kanter_core_b200._lib import call
from tests import graphs

pytestmark = pytest.mark.gpu


@pytest.fixture()
def jit_always():
    call("kc_debug_set_tuning", b"jit", 1)
    yield
    call("kc_debug_set_tuning", b"jit", 0)


def _specialised_kernel_ran():
    v, c, s = C.c_int32(), C.c_int32(), C.c_int32()
    call("kc_debug_last_tile_config", C.byref(v), C.byref(c), C.byref(s))
    return v.value < 0


@pytest.mark.parametrize("name", ["pow_node_rgba", "divide_node_gray", "invert_graph_node_import", "separate_node", "value_node"])
def test_goldens_byte_exact_through_specialised_kernels(tex_pro, jit_always, name):
    case = graphs.GOLDEN_CASES[name]()
    lg = graphs.run_product(tex_pro, case, request=False)
    got = lg.read_rgba(case.node, SlotId(0), kc.Size(*case.size))
    assert np.array_equal(got, case.expected())
    assert _specialised_kernel_ran()                      # NVRTC is part of the image: no silent fallback here


def test_config5_bit_exact_through_specialised_kernels(tex_pro, jit_always):
    size = 256
    g, out = graphs.config5_graph(size)
    inputs = graphs.config5_inputs(777, size)
    want = graphs.config5_oracle(g, out, inputs)
    got = graphs.config5_product(tex_pro, g, out, inputs).slot_data(out, SlotId(0)).image.planes()
    for c in range(4):
        assert np.array_equal(got[c].view(np.uint32), want[c].view(np.uint32)), c
    assert _specialised_kernel_ran()


def test_long_chain_and_ragged_sizes(tex_pro, jit_always):
    r = np.random.default_rng(12)
    for (h, w) in ((1, 1), (7, 13), (33, 1021), (300, 1000)):       # ragged tails, partial tiles
        A = r.random((h, w), dtype=np.float32) + np.float32(0.25)
        B = r.random((h, w), dtype=np.float32) + np.float32(0.25)
        lg = tex_pro.new_live_graph()
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tex_pro, [A])), 0)
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tex_pro, [B])), 1)
        a = lg.add_node(Node.new(NodeType.Embed(0)))
        b = lg.add_node(Node.new(NodeType.Embed(1)))
        cur, want = a, A
        for i in range(14):
            op = MixType(i % 5)
            m = lg.add_node(Node.new(NodeType.Mix(op)))
            lg.connect(cur, m, SlotId(0), SlotId(0))
            lg.connect(b, m, SlotId(0), SlotId(1))
            with np.errstate(all="ignore"):
                want = oracle.mix_plane(int(op), want, B)
            cur = m
        kc.LiveGraph.await_clean_read(lg, cur)
        got = lg.slot_data(cur, SlotId(0)).image.planes()[0]
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        assert np.array_equal(got[ok].view(np.uint32), want[ok].view(np.uint32)), (h, w)


def test_fast_mode_within_tolerance(tex_pro_fast, jit_always):
    r = np.random.default_rng(13)
    A = [r.random((256, 384), dtype=np.float32) for _ in range(4)]
    B = [r.random((256, 384), dtype=np.float32) for _ in range(4)]
    ia, ib = kc.SlotImage.from_planes(tex_pro_fast, A), kc.SlotImage.from_planes(tex_pro_fast, B)
    got = kc.mix(tex_pro_fast, MixType.Pow, kc.mix(tex_pro_fast, MixType.Multiply, ia, ib), ib).planes()
    for c in range(3):
        want = oracle.mix_plane(4, oracle.mix_plane(2, A[c], B[c]), B[c]).astype(np.float64)
        assert (np.abs(got[c].astype(np.float64) - want) <= 1e-6 + 1e-5 * np.abs(want)).all()
    assert _specialised_kernel_ran()


def test_automatic_policy_compiles_in_the_background(tex_pro_fast):
    """A tape that keeps coming back on a large plane is compiled on a background thread: the
    launches meanwhile are served by the interpreter (no stall), the specialised kernel takes
    over once it is ready, and both give the same pixels."""
    import time
    tp = tex_pro_fast
    s = 1024
    r = np.random.default_rng(11)
    planes = [r.random((s, s), dtype=np.float32) + np.float32(0.25) for _ in range(3)]
    imgs = [kc.SlotImage.from_planes(tp, [p]) for p in planes]

    def evaluate():
        # a tape no other test uses: ((a * b) / c - a) * c  -- five ops, 1 Mpx
        x = kc.mix(tp, MixType.Multiply, imgs[0], imgs[1])
        x = kc.mix(tp, MixType.Divide, x, imgs[2])
        x = kc.mix(tp, MixType.Subtract, x, imgs[0])
        x = kc.mix(tp, MixType.Multiply, x, imgs[2])
        return x.planes()[0]

    t0 = time.perf_counter()
    first = [evaluate() for _ in range(3)]                 # the third sighting starts the compile
    stall = time.perf_counter() - t0
    assert not _specialised_kernel_ran()                   # ... and is still served by the interpreter
    assert stall < 0.5, "the launching thread waited for NVRTC (%.2f s)" % stall
    assert kc.jit_wait(60000) == 0
    after = evaluate()
    assert _specialised_kernel_ran()
    want = ((planes[0] * planes[1]) / planes[2] - planes[0]) * planes[2]
    for got in first + [after]:
        assert np.allclose(got, want, rtol=1e-5, atol=1e-6)
    assert np.array_equal(first[0], after)                 # no pow in this tape: FAST == EXACT arithmetic, same bits
