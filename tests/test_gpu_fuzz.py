"""Seeded random node graphs, every slot of every node compared bit for bit with
the CPU oracle (EXACT mode), and within the stated tolerance in FAST mode.

The hand-written tests pin each operator; this one pins the planner: which nodes
fuse into one tape, how many sources a tape reads, temporaries, broadcast of 1x1
values, implicit resizes in the middle of a fused chain, Gray/Rgba coercion, fan-out
of one plane into several consumers, nested evaluation order.
"""
import numpy as np
import pytest

import kanter_core_b200 as kc
import oracle
from kanter_core_b200 import MixType, Node, NodeType, ResizeFilter, ResizePolicy, Size, SlotId
from kanter_core_b200._lib import call

from .test_gpu_ops import ABS, REL, bits_equal

pytestmark = pytest.mark.gpu

SIZES = [(64, 64), (96, 64), (50, 70), (33, 17), (128, 128), (1, 1), (7, 200)]
POLICIES = [ResizePolicy.MostPixels, ResizePolicy.LeastPixels, ResizePolicy.LargestAxes, ResizePolicy.SmallestAxes]
MIX = list(MixType)


def random_graph(seed, n_ops, h2n=True, typed=True, inputs=None, nested=True, sizes=None):
    """-> (NodeGraph, {embed id: planes}).  Every op node takes its inputs from earlier
    nodes, so the graph is a DAG.  With `typed` the generator tracks which outputs carry
    Gray and which Rgba data and only makes connections the operators accept at run time
    (HeightToNormal and CombineRgba read Gray, SeparateRgba reads Rgba, outputs read their
    own kind); without it only the static slot types are respected and some graphs fail
    to evaluate - in the reference as well.  `inputs` (a list of "gray" | "rgba") makes the
    graph an inner one: InputGray / InputRgba nodes "i0", "i1", .. instead of Embed nodes.
    With `nested`, some operators are Graph nodes holding such an inner graph.  The named
    outputs and their kinds are left in `graph.fuzz_outputs`."""
    r = np.random.default_rng(seed)
    SIZES = sizes or globals()["SIZES"]
    g = kc.NodeGraph.new()
    embeds = {}
    outs = []  # (node id, output slot, kind of the data "gray" | "rgba", static slot type "gray" | "rgba" | "any")

    def add(nt, policy=None, filt=None):
        n = Node.new(nt)
        if policy is not None:
            n.resize_policy = policy
        if filt is not None:
            n.resize_filter = filt
        return g.add_node(n)

    for k, kind in enumerate(inputs or []):
        nt = NodeType.InputRgba("i%d" % k) if kind == "rgba" else NodeType.InputGray("i%d" % k)
        outs.append((add(nt), 0, kind, kind))
    for eid in range(int(r.integers(2, 5)) if inputs is None else 0):
        w, h = [z for z in SIZES if z != (1, 1)][int(r.integers(len(SIZES) - 1))]
        nplanes = 4 if r.random() < 0.5 else 1
        lo, hi = (-0.5, 1.5) if r.random() < 0.3 else (0.0, 1.0)
        embeds[eid] = [(r.random((h, w), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)
                       for _ in range(nplanes)]
        outs.append((add(NodeType.Embed(eid)), 0, "rgba" if nplanes == 4 else "gray", "rgba"))  # node_type.rs:186-188
    for _ in range(int(r.integers(1, 3))):
        outs.append((add(NodeType.Value(float(np.float32(r.random() * 2.0)))), 0, "gray", "gray"))

    def policy(connected):
        p = r.random()
        if p < 0.55 or not connected:
            return POLICIES[int(r.integers(len(POLICIES)))]
        if p < 0.75:
            return ResizePolicy.SpecificSlot(SlotId(connected[int(r.integers(len(connected)))]))
        w, h = SIZES[int(r.integers(len(SIZES)))]
        return ResizePolicy.SpecificSize(Size.new(w, h))

    def pick(want=None):
        pool = [o for o in outs if not typed or want is None or (o[2] == want and o[3] in (want, "any"))]
        return pool[int(r.integers(len(pool)))] if pool else None

    def build(nt, filt, sources):
        """sources: {input slot: picked output}.  The policy is drawn knowing which slots
        are connected; edges the static slot types refuse (untyped mode) are left out."""
        n = add(nt, policy(sorted(sources)), filt)
        kinds = {}
        for in_slot, (src, s, kind, _) in sources.items():
            try:
                g.connect(src, n, SlotId(s), SlotId(in_slot))
                kinds[in_slot] = kind
            except kc.TexProError:
                assert not typed
        return n, kinds

    for _ in range(n_ops):
        p = r.random()
        filt = list(ResizeFilter)[int(r.integers(len(ResizeFilter)))]
        if nested and typed and inputs is None and r.random() < 0.12:
            # a Graph node: inner inputs take outer sources of their kind, inner outputs become sources
            want = ["rgba" if r.random() < 0.5 else "gray" for _ in range(int(r.integers(1, 4)))]
            srcs = [pick(k) for k in want]
            if all(srcs):
                inner = random_graph(int(r.integers(1 << 30)), int(r.integers(2, 7)), h2n=h2n, inputs=want, sizes=sizes)
                if inner.fuzz_outputs:
                    n = add(NodeType.Graph(inner), None, filt)
                    for k, (src, sl, _, _) in enumerate(srcs):
                        g.connect(src, n, SlotId(sl), inner.input_slot_id_with_name("i%d" % k))
                    for name, kind in inner.fuzz_outputs:
                        outs.append((n, int(inner.output_slot_id_with_name(name)), kind, kind))
                    continue
        if p < 0.55:
            src = {0: pick()}
            if r.random() < 0.9:
                src[1] = pick()
            if r.random() < 0.05:
                del src[0]           # right side only: zeros op right (mix.rs:77-83)
            n, kinds = build(NodeType.Mix(MIX[int(r.integers(len(MIX)))]), filt, src)
            if kinds:
                outs.append((n, 0, kinds.get(0) or kinds[1], "any"))
        elif p < 0.70:
            one = pick("rgba")
            if one:
                n, kinds = build(NodeType.SeparateRgba, filt, {0: one})
                if kinds:
                    outs.extend((n, s, "gray", "gray") for s in range(4))
        elif p < 0.85:
            src = {s: pick("gray") for s in range(4) if r.random() < 0.8}
            src = {s: o for s, o in src.items() if o}
            n, kinds = build(NodeType.CombineRgba, filt, src)
            outs.append((n, 0, "rgba", "rgba"))
        elif h2n:
            one = pick("gray")
            n, kinds = build(NodeType.HeightToNormal, filt, {0: one}) if one else (None, None)
            if kinds:
                outs.append((n, 0, "rgba", "rgba"))
    g.fuzz_outputs = []
    for k in range(int(r.integers(1, 4))):
        rgba = r.random() < 0.5
        one = pick("rgba" if rgba else "gray")
        if one:
            nt = NodeType.OutputRgba("o%d" % k) if rgba else NodeType.OutputGray("o%d" % k)
            n = add(nt)
            try:
                g.connect(one[0], n, SlotId(one[1]), SlotId(0))
                g.fuzz_outputs.append(("o%d" % k, "rgba" if rgba else "gray"))
            except kc.TexProError:
                assert not typed
    return (g, embeds) if inputs is None else g


def evaluate_both(tp, graph, embeds, use_cache=True):
    """use_cache (src/live_graph.rs:72) keeps every node's slot data; without it the live
    graph drops a parent's data once its children are clean, as the reference does."""
    og = oracle.from_node_graph(graph)
    for eid, planes in embeds.items():
        og.embed(eid, planes)
    og.eval()
    lg = tp.new_live_graph()
    lg.use_cache = use_cache
    lg.set_node_graph(graph)
    for eid, planes in embeds.items():
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tp, planes)), eid)
    if use_cache:
        for n in graph.nodes:        # nodes no output depends on are evaluated only when asked for
            kc.LiveGraph.await_clean_read(lg, n.node_id)
    return og, lg


def each_slot(og, lg, graph):
    for n in graph.nodes:
        nid = int(n.node_id)
        if not lg.use_cache:         # ask node by node; whatever was dropped is evaluated again
            kc.LiveGraph.await_clean_read(lg, n.node_id)
        for s in og.slot_ids(nid):
            yield nid, n, s, og.slot(nid, s), lg.slot_data(n.node_id, SlotId(s)).image.planes()


def describe(graph, nid):
    ins = [(int(e.output_id), int(e.output_slot), int(e.input_slot)) for e in graph.edges if int(e.input_id) == nid]
    return "node %d inputs(out node, out slot, in slot)=%s" % (nid, ins)


@pytest.mark.parametrize("seed", range(40))
def test_random_graph_exact(tex_pro, seed):
    graph, embeds = random_graph(1000 + seed, n_ops=6 + seed % 17)
    og, lg = evaluate_both(tex_pro, graph, embeds, use_cache=seed % 2 == 0)
    checked = 0
    for nid, n, s, want, got in each_slot(og, lg, graph):
        assert len(got) == len(want), (describe(graph, nid), n.node_type, s)
        for c in range(len(want)):
            assert bits_equal(got[c], want[c]), (describe(graph, nid), n.node_type, s, c, got[c].shape, want[c].shape)
        checked += 1
    assert checked >= len(embeds)


@pytest.mark.parametrize("seed", range(12))
def test_random_graph_exact_specialised_kernels(tex_pro, seed):
    """The same comparison with every tape compiled to its own kernel (KC_JIT=1 policy)."""
    call("kc_debug_set_tuning", b"jit", 1)
    try:
        graph, embeds = random_graph(2000 + seed, n_ops=8 + seed)
        og, lg = evaluate_both(tex_pro, graph, embeds)
        for nid, n, s, want, got in each_slot(og, lg, graph):
            for c in range(len(want)):
                assert bits_equal(got[c], want[c]), (describe(graph, nid), n.node_type, s, c)
    finally:
        call("kc_debug_set_tuning", b"jit", 0)


@pytest.mark.parametrize("seed", range(24))
def test_random_graph_that_may_not_evaluate(tex_pro, seed):
    """Connections checked against the static slot types only: Gray data behind an Rgba
    slot, Rgba into HeightToNormal, SpecificSlot naming an empty slot ...  Where the
    reference's evaluation fails, this one has to fail as well (no made-up pixels); where
    it succeeds, every slot matches bit for bit."""
    graph, embeds = random_graph(4000 + seed, n_ops=6 + seed % 17, typed=False)
    og = oracle.from_node_graph(graph)
    for eid, planes in embeds.items():
        og.embed(eid, planes)
    try:
        og.eval()
    except oracle.OracleError:
        lg = tex_pro.new_live_graph()
        lg.set_node_graph(graph)
        for eid, planes in embeds.items():
            lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tex_pro, planes)), eid)
        with pytest.raises(kc.TexProError):
            for n in graph.nodes:
                kc.LiveGraph.await_clean_read(lg, n.node_id)
        return
    og, lg = evaluate_both(tex_pro, graph, embeds)
    for nid, n, s, want, got in each_slot(og, lg, graph):
        assert len(got) == len(want), (describe(graph, nid), n.node_type, s)
        for c in range(len(want)):
            assert bits_equal(got[c], want[c]), (describe(graph, nid), n.node_type, s, c)


@pytest.mark.parametrize("seed", range(16))
def test_random_graph_fast_tracks_the_oracle(tex_pro_fast, seed):
    """FAST mode through whole graphs.  The per-operator tolerance (1e-5 rel / 1e-6 abs,
    test_gpu_ops.py) cannot hold sample for sample across a chain: one Subtract that
    cancels followed by a Divide amplifies a 4e-7 difference without bound.  What a
    planner mistake in the FAST instantiations would look like is a gross error over
    whole planes, so the check is: at least 99% of every plane's finite samples are
    within 100x the per-operator tolerance, relative to the larger of the sample and the
    node's largest input magnitude.  No HeightToNormal (test_config5 bounds it)."""
    graph, embeds = random_graph(3000 + seed, n_ops=6 + seed, h2n=False)
    og, lg = evaluate_both(tex_pro_fast, graph, embeds)
    mags = {}
    for nid, n, s, want, got in each_slot(og, lg, graph):
        fin = [np.abs(p[np.isfinite(p)]) for p in want]
        mags[(nid, s)] = max([float(f.max()) for f in fin if f.size] or [0.0])
    for nid, n, s, want, got in each_slot(og, lg, graph):
        scale = max([mags.get((int(e.output_id), int(e.output_slot)), 0.0) for e in graph.edges
                     if int(e.input_id) == nid] or [0.0])
        scale = min(scale, 1e30)
        for c in range(len(want)):
            a, b = got[c].astype(np.float64), want[c].astype(np.float64)
            assert a.shape == b.shape
            fin = np.isfinite(a) & np.isfinite(b)
            err = np.abs(a[fin] - b[fin])
            bound = 100.0 * (ABS + REL * np.maximum(np.abs(b[fin]), scale))
            bad = int((err > bound).sum())
            assert bad <= max(1, int(fin.sum()) // 100), (describe(graph, nid), n.node_type, s, c, bad, float(err.max()))
            assert int((np.isfinite(a) != np.isfinite(b)).sum()) <= max(1, a.size // 100)  # inf/NaN in the same places


def test_threads_share_one_context(tex_pro):
    """The reference's engine evaluates live graphs on worker threads (src/engine.rs:200-307)
    that share the TextureProcessor.  Here: eight threads, each with its own live graph on the
    same context (ctypes drops the GIL around every call), three rounds each; every slot still
    matches the oracle bit for bit."""
    import threading
    cases = [random_graph(5000 + i, n_ops=8 + i) for i in range(8)]
    wants = []
    for graph, embeds in cases:
        og = oracle.from_node_graph(graph)
        for eid, planes in embeds.items():
            og.embed(eid, planes)
        og.eval()
        wants.append({(int(n.node_id), s): og.slot(int(n.node_id), s) for n in graph.nodes for s in og.slot_ids(int(n.node_id))})
    errors = []

    def work(i):
        try:
            graph, embeds = cases[i]
            for _ in range(3):
                lg = tex_pro.new_live_graph()
                lg.use_cache = True
                lg.set_node_graph(graph)
                for eid, planes in embeds.items():
                    lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tex_pro, planes)), eid)
                for n in graph.nodes:
                    kc.LiveGraph.await_clean_read(lg, n.node_id)
                for (nid, s), want in wants[i].items():
                    got = lg.slot_data(kc.NodeId(nid), SlotId(s)).image.planes()
                    assert len(got) == len(want)
                    for c in range(len(want)):
                        assert bits_equal(got[c], want[c]), (i, nid, s, c)
                lg.close()
        except BaseException as e:  # noqa: BLE001 - reported in the main thread
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(cases))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("threshold", [16, 64 * 1024, 1 << 20])
@pytest.mark.parametrize("seed", range(10))
def test_random_graph_exact_under_memory_pressure(tex_pro, seed, threshold):
    """The same graphs with the spill queue squeezed (src/transient_buffer.rs:250-411): with a
    threshold of 16 B, 64 KiB or 1 MiB nearly every plane leaves HBM as soon as nothing pins it and
    comes back when a kernel needs it.  Results do not change, and planes really moved."""
    graph, embeds = random_graph(8000 + seed, n_ops=8 + seed)
    before = tex_pro.spill_stats()
    tex_pro.set_memory_threshold(threshold)
    try:
        og, lg = evaluate_both(tex_pro, graph, embeds, use_cache=seed % 2 == 0)
        for nid, n, s, want, got in each_slot(og, lg, graph):
            assert len(got) == len(want)
            for c in range(len(want)):
                assert bits_equal(got[c], want[c]), (describe(graph, nid), n.node_type, s, c)
        lg.close()
    finally:
        tex_pro.set_memory_threshold(0)            # 0: no limit
    after = tex_pro.spill_stats()
    if threshold == 16:
        assert after["spills"] > before["spills"] and after["reloads"] > before["reloads"]


BIG = [(1500, 1100), (2048, 1024), (1031, 997), (640, 2200), (1024, 1024), (1, 1), (3000, 500)]


@pytest.mark.parametrize("seed", range(10))
def test_random_graph_exact_at_megapixel_sizes(tex_pro, seed):
    """Planes of 1-2 Mpx with ragged edges: many tiles per plane, tails that are not a multiple
    of a float4 or of a tile, resizes between unrelated sizes in both directions (short- and
    long-window kernels), and tapes hot enough for the automatic specialisation to kick in
    (the graph is evaluated four times; the background compile is waited for in between)."""
    graph, embeds = random_graph(9000 + seed, n_ops=7 + seed, sizes=BIG)
    og = oracle.from_node_graph(graph)
    for eid, planes in embeds.items():
        og.embed(eid, planes)
    og.eval()
    lg = tex_pro.new_live_graph()
    lg.use_cache = True
    lg.set_node_graph(graph)
    for eid, planes in embeds.items():
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tex_pro, planes)), eid)
    for rounds in range(4):
        if rounds:
            for eid, planes in embeds.items():            # "new" pixels: everything downstream is dirty again
                lg.replace_embedded(kc.SlotImage.from_planes(tex_pro, planes), eid)
        if rounds == 3:
            kc.jit_wait()
        for n in graph.nodes:
            kc.LiveGraph.await_clean_read(lg, n.node_id)
    for nid, n, s, want, got in each_slot(og, lg, graph):
        assert len(got) == len(want)
        for c in range(len(want)):
            assert bits_equal(got[c], want[c]), (describe(graph, nid), n.node_type, s, c, got[c].shape)
    lg.close()


def test_contexts_on_separate_threads_do_not_interfere():
    """Four contexts, one thread each, no lock in common: what they share inside the library (the
    cache of specialised kernels, occupancy and attribute caches, the planner's stamps) is
    process-wide state that has to hold up."""
    import threading
    errors = []

    def work(i):
        tp = None
        try:
            tp = kc.TextureProcessor.new(math_mode=kc.MATH_EXACT)
            for rep in range(3):
                graph, embeds = random_graph(9500 + 10 * i + rep, n_ops=8 + i)
                og, lg = evaluate_both(tp, graph, embeds)
                for nid, n, s, want, got in each_slot(og, lg, graph):
                    assert len(got) == len(want)
                    for c in range(len(want)):
                        assert bits_equal(got[c], want[c]), (i, rep, nid, s, c)
                lg.close()
        except BaseException as e:  # noqa: BLE001 - reported in the main thread
            errors.append((i, repr(e)))
        finally:
            if tp is not None:
                tp.close()

    threads = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
