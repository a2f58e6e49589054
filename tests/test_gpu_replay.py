"""Evaluation replay (kc_live_graph_set_replay): the second identical request over an unchanged graph and the same input
planes is captured into one executable CUDA graph and later ones replay it.  What must hold: results identical to the
ordinary evaluation (and so to the oracle) bit for bit whatever the pixel content, planes somebody still holds are never
overwritten, any change of the graph or of its inputs' identity falls back and re-captures."""
import ctypes as C

import numpy as np
import pytest

import kanter_core_b200 as kc
import oracle
from kanter_core_b200 import MixType, Node, NodeType, SlotId
from kanter_core_b200._lib import call, kc_image
from tests import graphs

pytestmark = pytest.mark.gpu


def bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


class Buffers:
    """Images over planes whose CONTENT is rewritten in place (kc_plane_upload): what a caller streaming frames through
    fixed device buffers does."""

    def __init__(self, tp, shapes):
        self.tp, self.images = tp, []
        for n, (h, w) in shapes:
            im = kc_image()
            im.kind, im.width, im.height = (1 if n == 4 else 0), w, h
            for c in range(n):
                pl = C.c_void_p()
                call("kc_plane_create", tp._ctx._h, w, h, C.byref(pl))
                im.planes[c] = pl
            self.images.append(kc.SlotImage(tp._ctx, im))

    def fill(self, planes_per_image):
        for img, planes in zip(self.images, planes_per_image):
            for c, p in enumerate(planes):
                a = np.ascontiguousarray(p, dtype=np.float32)
                call("kc_plane_upload", img._im.planes[c], a.ctypes.data)


def _config5(tp, size):
    g, out = graphs.config5_graph(size)
    lsize = max(1, size // 4)
    bufs = Buffers(tp, [(4, (size, size)), (1, (size, size)), (4, (lsize, lsize))])
    lg = tp.new_live_graph()
    lg.set_node_graph(g)
    for eid, img in enumerate(bufs.images):
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, img), eid)
    return g, out, bufs, lg


@pytest.mark.parametrize("size", [64, 256])
def test_replay_is_bit_identical_to_the_oracle_frame_after_frame(tex_pro, size):
    g, out, bufs, lg = _config5(tex_pro, size)
    lg.set_replay(True)
    for frame in range(7):
        inputs = graphs.config5_inputs(300 + frame, size)
        bufs.fill(inputs)
        for eid, img in enumerate(bufs.images):            # "new inputs arrived": same planes, new content
            lg.replace_embedded(img, eid)
        lg.request(out)
        got = lg.slot_data(out, SlotId(0)).image.planes()
        want = graphs.config5_oracle(g, out, inputs)
        for c in range(4):
            assert bits_equal(got[c], want[c]), (frame, c)
        del got
    st = lg.replay_stats()
    assert st["captures"] >= 1 and st["replays"] >= 1, st     # (a specialised kernel arriving from another test's compile re-captures)
    assert lg.last_run_stats()["kernels"] > 0


def test_replay_never_overwrites_a_plane_somebody_holds(tex_pro):
    size = 64
    g, out, bufs, lg = _config5(tex_pro, size)
    lg.set_replay(True)
    held, held_want = None, None
    for frame in range(8):
        inputs = graphs.config5_inputs(400 + frame, size)
        bufs.fill(inputs)
        for eid, img in enumerate(bufs.images):
            lg.replace_embedded(img, eid)
        lg.request(out)
        want = graphs.config5_oracle(g, out, inputs)
        img = lg.slot_data(out, SlotId(0)).image
        for c in range(4):
            assert bits_equal(img.planes()[c], want[c]), (frame, c)
        if held is not None:                                # the frame before this one, still in a caller's hands
            for c in range(4):
                assert bits_equal(held.planes()[c], held_want[c]), ("held", frame, c)
        if frame in (3, 4):
            held, held_want = img, want                     # keep it across the next evaluation
        else:
            held, held_want = None, None
        del img
    st = lg.replay_stats()
    assert st["captures"] >= 1 and st["replays"] >= 1, st   # replays resume once the caller lets go


def test_replay_follows_changes_of_the_graph_and_of_the_inputs(tex_pro):
    tp = tex_pro
    size = 96
    r = np.random.default_rng(5)
    a0, b0 = r.random((size, size), dtype=np.float32), r.random((size, size), dtype=np.float32)
    bufs = Buffers(tp, [(1, (size, size)), (1, (size, size))])
    bufs.fill([[a0], [b0]])
    lg = tp.new_live_graph()
    lg.set_replay(True)
    for eid, img in enumerate(bufs.images):
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, img), eid)
    ea = lg.add_node(Node.new(NodeType.Embed(0)))
    eb = lg.add_node(Node.new(NodeType.Embed(1)))
    m = lg.add_node(Node.new(NodeType.Mix(MixType.Multiply)))
    h = lg.add_node(Node.new(NodeType.HeightToNormal))
    o = lg.add_node(Node.new(NodeType.OutputRgba("out")))
    lg.connect(ea, m, SlotId(0), SlotId(0))
    lg.connect(eb, m, SlotId(0), SlotId(1))
    lg.connect(m, h, SlotId(0), SlotId(0))
    lg.connect(h, o, SlotId(0), SlotId(0))

    def frame(op, a, b):
        for eid, img in enumerate(bufs.images):
            lg.replace_embedded(img, eid)
        lg.request(o)
        got = lg.slot_data(o, SlotId(0)).image.planes()
        want = oracle.height_to_normal(oracle.mix_plane(op, a, b))
        for c in range(3):
            assert bits_equal(got[c], want[c]), c

    for _ in range(4):
        frame(2, a0, b0)
    assert lg.replay_stats()["replays"] >= 2
    lg.set_mix_type(m, MixType.Add)                         # the graph changed: the plan is gone
    for _ in range(4):
        frame(0, a0, b0)
    a1 = r.random((size, size), dtype=np.float32)
    other = kc.SlotImage.from_planes(tp, [a1])              # ANOTHER plane embedded under id 0: another plan
    for _ in range(4):
        lg.replace_embedded(other, 0)
        lg.replace_embedded(bufs.images[1], 1)
        lg.request(o)
        got = lg.slot_data(o, SlotId(0)).image.planes()
        want = oracle.height_to_normal(oracle.mix_plane(0, a1, b0))
        for c in range(3):
            assert bits_equal(got[c], want[c]), c
    st = lg.replay_stats()
    assert st["captures"] >= 3, st
    lg.set_replay(False)
    frame(0, a0, b0)                                        # and off again: the ordinary path


@pytest.mark.parametrize("name", sorted(n for n in graphs.GOLDEN_CASES if graphs.GOLDEN_CASES[n]().embeds))
def test_replay_on_the_reference_goldens(tex_pro, name):
    """The reference's golden graphs that take embedded images, evaluated five times with replay on (the images embedded
    again before each): the reference's bytes every time, through the ordinary pass, the capture and the replays."""
    case = graphs.GOLDEN_CASES[name]()
    lg = tex_pro.new_live_graph()
    lg.set_node_graph(case.graph)
    lg.set_replay(True)
    imgs = {eid: kc.SlotImage.from_u8(tex_pro, graphs.decode(path)) for eid, path in case.embeds.items()}
    for eid, img in imgs.items():
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, img), eid)
    want = case.expected()
    for i in range(5):
        for eid, img in imgs.items():
            lg.replace_embedded(img, eid)
        lg.request(case.node)
        got = lg.buffer_rgba(case.node, SlotId(0))
        assert np.array_equal(got, want), (i, int((got != want).sum()))
    assert lg.replay_stats()["replays"] >= 2


@pytest.mark.parametrize("seed", range(16))
def test_replay_on_random_graphs(tex_pro, seed):
    """Seeded random DAGs (every node kind, resize policy and filter, nested graphs; tests/test_gpu_fuzz.py) with ragged
    inputs up to a megapixel: all output nodes requested together six times with the same images embedded again before
    each -- through the ordinary pass, the capture (whose chain of launches is re-wired to the true dependencies, so the
    independent branches of these graphs run side by side) and the replays.  Every output bit-identical to the oracle
    every time."""
    from tests.test_gpu_fuzz import BIG, bits_equal, random_graph        # (that bits_equal lets NaNs differ in sign and payload: x86 and the GPU do)
    graph, embeds = random_graph(7000 + seed, n_ops=8 + seed % 9, sizes=BIG if seed % 4 == 0 else None)
    og = oracle.from_node_graph(graph)
    for eid, planes in embeds.items():
        og.embed(eid, planes)
    og.eval()
    outs = [n.node_id for n in graph.nodes if n.node_type.is_output()]
    if not outs:
        pytest.skip("the generator made no output node")
    lg = tex_pro.new_live_graph()
    lg.set_node_graph(graph)
    lg.set_replay(True)
    imgs = {eid: kc.SlotImage.from_planes(tex_pro, planes) for eid, planes in embeds.items()}
    for eid, img in imgs.items():
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, img), eid)
    for rounds in range(6):
        for eid, img in imgs.items():
            lg.replace_embedded(img, eid)
        lg.request_many(outs)
        for nid in outs:
            for s in og.slot_ids(int(nid)):
                want = og.slot(int(nid), s)
                got = lg.slot_data(nid, SlotId(s)).image.planes()
                assert len(got) == len(want)
                for c in range(len(want)):
                    assert bits_equal(got[c], want[c]), (rounds, int(nid), s, c)
                del got
    lg.close()


def test_replay_at_a_size_the_tensor_map_kernels_take(tex_pro):
    """The 32-node graph at 1024^2 (resizes 256^2 -> 1024^2 through the tensor-map kernel): captured, re-wired, replayed."""
    size = 1024
    kc.jit_wait()          # a specialised kernel arriving from an earlier test's background compile re-captures the plan
    g, out, bufs, lg = _config5(tex_pro, size)
    lg.set_replay(True)
    for frame in range(8):
        inputs = graphs.config5_inputs(500 + frame, size)
        bufs.fill(inputs)
        for eid, img in enumerate(bufs.images):
            lg.replace_embedded(img, eid)
        lg.request(out)
        got = lg.slot_data(out, SlotId(0)).image.planes()
        want = graphs.config5_oracle(g, out, inputs)
        for c in range(4):
            assert bits_equal(got[c], want[c]), (frame, c)
        del got
    assert lg.replay_stats()["replays"] >= 1, lg.replay_stats()


@pytest.mark.parametrize("lanes", [2, 3])
def test_concurrent_section_runs_independent_graphs_on_lanes_bit_identically(tex_pro, lanes):
    """kc_context_concurrent_begin/end: the replays of several live graphs go to side streams.  Frame after frame (inputs
    rewritten in place between sections) every graph's result equals the oracle bit for bit; a download inside a section is
    ordered behind the lanes; a graph that has no plan yet simply evaluates on the context's stream."""
    size, n = 128, 4
    sets = [_config5(tex_pro, size) for _ in range(n)]
    for g, out, bufs, lg in sets:
        lg.set_replay(True)
    for frame in range(6):
        inputs = [graphs.config5_inputs(900 + 10 * frame + i, size) for i in range(n)]
        for (g, out, bufs, lg), inp in zip(sets, inputs):
            bufs.fill(inp)
        with tex_pro.concurrent(lanes):
            for rep in range(2):                                  # twice: the second round replays over the first one's arenas
                for g, out, bufs, lg in sets:
                    for eid, img in enumerate(bufs.images):
                        lg.replace_embedded(img, eid)
                    lg.request(out)
            if frame == 3:                                        # reading inside the section: the download waits for the lanes
                g, out, bufs, lg = sets[1]
                got = lg.slot_data(out, SlotId(0)).image.planes()
                want = graphs.config5_oracle(g, out, inputs[1])
                assert all(bits_equal(got[c], want[c]) for c in range(4))
                del got
        for i, (g, out, bufs, lg) in enumerate(sets):
            got = lg.slot_data(out, SlotId(0)).image.planes()
            want = graphs.config5_oracle(g, out, inputs[i])
            for c in range(4):
                assert bits_equal(got[c], want[c]), (frame, i, c)
            del got
    assert sum(lg.replay_stats()["replays"] for _, _, _, lg in sets) >= n * 4
    with pytest.raises(Exception):
        call("kc_context_concurrent_begin", tex_pro._ctx._h, 99)


def test_concurrent_section_survives_teardown_while_lanes_are_busy():
    """A plan dropped (graph replaced), a live graph released and the context destroyed while replays are still running on
    the lanes: every teardown path waits for the lanes before it frees what they write."""
    tp = kc.TextureProcessor()
    size, n = 256, 3
    sets = [_config5(tp, size) for _ in range(n)]
    inputs = [graphs.config5_inputs(950 + i, size) for i in range(n)]
    for (g, out, bufs, lg), inp in zip(sets, inputs):
        lg.set_replay(True)
        bufs.fill(inp)
    for _ in range(3):                                            # ordinary pass, capture, first replay
        with tp.concurrent(3):
            for g, out, bufs, lg in sets:
                for eid, img in enumerate(bufs.images):
                    lg.replace_embedded(img, eid)
                lg.request(out)
    call("kc_context_concurrent_begin", tp._ctx._h, 3)
    for g, out, bufs, lg in sets:
        for eid, img in enumerate(bufs.images):
            lg.replace_embedded(img, eid)
        lg.request(out)
    g1, out1, bufs1, lg1 = sets[1]
    lg1.set_node_graph(graphs.config5_graph(size)[0])           # a new graph: the plan goes while its replay may still run
    g2, out2, bufs2, lg2 = sets.pop(2)
    del lg2, bufs2                                                # a live graph and its inputs released inside the section
    g0, out0, bufs0, lg0 = sets[0]
    got = lg0.slot_data(out0, SlotId(0)).image.planes()           # (a download: ordered behind the lanes)
    want = graphs.config5_oracle(g0, out0, inputs[0])
    assert all(bits_equal(got[c], want[c]) for c in range(4))
    del got, lg0, lg1, bufs0, bufs1, sets
    del tp                                                        # the context goes with the section still open
