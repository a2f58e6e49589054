"""The spill queue (TransientBufferQueue, src/transient_buffer.rs:250-411; memory_threshold,
src/texture_processor.rs:19): above the threshold the least recently used planes leave HBM for
pinned host memory, reading them brings them back, and results never change.  The residency
pattern asserted by the reference's `drive_cache` test (tests/integration_tests.rs:142-248) is
reproduced on planes that really hold pixels (the reference's 1x1 Value planes are constant
descriptors here and occupy nothing)."""
import numpy as np
import pytest

import kanter_core_b200 as kc
from kanter_core_b200 import LiveGraph, MixType, Node, NodeType, SlotId
from tests import graphs

pytestmark = pytest.mark.gpu

S = 64
PLANE = S * S * 4


def _chain(tp):
    """embed A -> Mix(Add, A, B) = m1 -> Mix(Multiply, m1, B) = m2, use_cache on (every node keeps its data)"""
    r = np.random.default_rng(5)
    A = r.random((S, S), dtype=np.float32)
    B = r.random((S, S), dtype=np.float32)
    lg = tp.new_live_graph()
    lg.use_cache = True
    lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tp, [A])), 0)
    lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tp, [B])), 1)
    a = lg.add_node(Node.new(NodeType.Embed(0)))
    b = lg.add_node(Node.new(NodeType.Embed(1)))
    m1 = lg.add_node(Node.new(NodeType.Mix(MixType.Add)))
    m2 = lg.add_node(Node.new(NodeType.Mix(MixType.Multiply)))
    lg.connect(a, m1, SlotId(0), SlotId(0))
    lg.connect(b, m1, SlotId(0), SlotId(1))
    lg.connect(m1, m2, SlotId(0), SlotId(0))
    lg.connect(b, m2, SlotId(0), SlotId(1))
    return lg, (a, b, m1, m2), (A, B)


def test_drive_cache_residency(tex_pro):
    tp = tex_pro
    try:
        lg, (a, b, m1, m2), (A, B) = _chain(tp)
        tp.set_memory_threshold(PLANE)                       # room for ONE plane
        LiveGraph.await_clean_read(lg, m2)
        # the newest result is in HBM, everything older has been pushed out (drive_cache :183-194)
        assert lg.slot_in_memory(m2, SlotId(0))
        for n in (a, b, m1):
            assert not lg.slot_in_memory(n, SlotId(0)), int(n)
        st = tp.spill_stats()
        assert st["spills"] >= 3 and st["bytes_spilled"] >= 3 * PLANE
        # values survive the round trip bit for bit (drive_cache :196-222)
        want_m1 = (A + B).astype(np.float32)
        got = lg.slot_data(m1, SlotId(0)).image.planes()[0]  # reading loads it back ...
        assert np.array_equal(got, want_m1)
        assert lg.slot_in_memory(m1, SlotId(0))              # ... (drive_cache :226-247)
        assert not lg.slot_in_memory(m2, SlotId(0))          # and the one plane of room went to it
        assert np.array_equal(lg.slot_data(m2, SlotId(0)).image.planes()[0], (want_m1 * B).astype(np.float32))
        assert tp.spill_stats()["reloads"] >= 2
        tp.set_memory_threshold(0)                            # no limit again: nothing else moves
        n0 = tp.spill_stats()["spills"]
        assert np.array_equal(lg.slot_data(a, SlotId(0)).image.planes()[0], A)
        assert tp.spill_stats()["spills"] == n0
    finally:
        tp.set_memory_threshold(0)


@pytest.mark.parametrize("budget_planes", [1, 3, 8])
def test_results_do_not_depend_on_the_threshold(tex_pro, budget_planes):
    """The 32-node graph (fused groups, HeightToNormal, two resizes, nested graph) evaluated with
    room for only a few planes: inputs and intermediates bounce between HBM and the host, the
    output stays bit-identical to the CPU oracle."""
    tp = tex_pro
    size = 128
    g, out = graphs.config5_graph(size)
    inputs = graphs.config5_inputs(321, size)
    want = graphs.config5_oracle(g, out, inputs)
    try:
        tp.set_memory_threshold(budget_planes * size * size * 4)
        lg = graphs.config5_product(tp, g, out, inputs)
        got = lg.slot_data(out, SlotId(0)).image.planes()
        for c in range(4):
            assert np.array_equal(got[c].view(np.uint32), want[c].view(np.uint32)), c
        assert tp.spill_stats()["spills"] > 0
        # a second evaluation with the inputs partly on the host
        for eid in range(3):
            lg.replace_embedded(kc.SlotImage.from_planes(tp, inputs[eid]), eid)
        assert np.array_equal(lg.read_rgba(out, SlotId(0), kc.Size(size, size)),
                              __import__("oracle").to_u8(want, False))
    finally:
        tp.set_memory_threshold(0)


def test_exposed_device_pointers_are_never_spilled(tex_pro):
    import ctypes as C
    from kanter_core_b200._lib import call
    tp = tex_pro
    try:
        img = kc.SlotImage.from_planes(tp, [np.ones((S, S), np.float32)])
        ptr = C.c_void_p()
        call("kc_plane_device_ptr", img.plane_handles()[0], C.byref(ptr))
        tp.set_memory_threshold(16)
        others = [kc.SlotImage.from_planes(tp, [np.full((S, S), i, np.float32)]) for i in range(4)]
        v = C.c_int32()
        call("kc_plane_in_memory", img.plane_handles()[0], C.byref(v))
        assert v.value == 1
        assert np.array_equal(img.planes()[0], np.ones((S, S), np.float32))
        for i, o in enumerate(others):
            assert np.array_equal(o.planes()[0], np.full((S, S), i, np.float32))
    finally:
        tp.set_memory_threshold(0)
