"""Incremental evaluation under random edits.  A live graph is evaluated, then edited
again and again - operator changed, constant changed, embedded pixels replaced (also
with another size), an input rewired, a node added, a node removed - and after every
edit a few nodes are requested and compared, bit for bit, with a from-scratch CPU
oracle evaluation of the graph as it now stands.  What this pins is the dirty
propagation of src/live_graph.rs (set_state / propagate to children, :243-320,
:426-600): a stale plane left behind by a missed invalidation shows up as a mismatch.
"""
import numpy as np
import pytest

import kanter_core_b200 as kc
import oracle
from kanter_core_b200 import MixType, Node, NodeType, ResizeFilter, ResizePolicy, Side, SlotId

from .test_gpu_fuzz import SIZES, random_graph
from .test_gpu_ops import bits_equal

pytestmark = pytest.mark.gpu

def data_kinds(lg, embeds):
    """{(node, out slot): "gray" | "rgba"} of the graph as it stands (sources before consumers)."""
    kinds = {}
    ins = {}
    for e in lg.edges:
        ins.setdefault(int(e.input_id), {})[int(e.input_slot)] = (int(e.output_id), int(e.output_slot))
    todo = list(lg.nodes)
    while todo:
        later = []
        for n in todo:
            nid, k = int(n.node_id), n.node_type.kind
            if any(o not in kinds for o in ins.get(nid, {}).values()):
                later.append(n)
                continue
            src = {s: kinds[o] for s, o in ins.get(nid, {}).items()}
            if k == 6:
                kinds[(nid, 0)] = "rgba" if len(embeds[int(n.node_type.payload)]) == 4 else "gray"
            elif k == 8:
                kinds[(nid, 0)] = "gray"
            elif k == 9:
                kinds[(nid, 0)] = src.get(0) or src.get(1) or "gray"
            elif k in (10, 12):
                kinds[(nid, 0)] = "rgba"
            elif k == 11:
                for s in range(4):
                    kinds[(nid, s)] = "gray"
            elif k == 4:     # a nested graph: one output slot per inner Output node, typed by it
                inner = n.node_type.payload
                for m in inner.nodes:
                    if m.node_type.kind in (2, 3):
                        slot = int(inner.output_slot_id_with_name(m.node_type.payload))
                        kinds[(nid, slot)] = "gray" if m.node_type.kind == 2 else "rgba"
            else:
                kinds[(nid, -1)] = "sink"      # outputs: nothing reads them
        assert len(later) < len(todo), "cycle"
        todo = later
    return {k: v for k, v in kinds.items() if v != "sink"}


def descendants(edges, nid):
    out, work = {int(nid)}, [int(nid)]
    while work:
        x = work.pop()
        for e in edges:
            if int(e.output_id) == x and int(e.input_id) not in out:
                out.add(int(e.input_id))
                work.append(int(e.input_id))
    return out


def static_type(lg, nid, slot, kinds):
    k = lg.node(nid).node_type.kind
    if k == 4:
        return kinds[(nid, slot)]
    return {6: "rgba", 8: "gray", 9: "any", 10: "rgba", 11: "gray", 12: "rgba"}.get(k)


def check(tp, lg, embeds, r, how_many):
    og = oracle.from_node_graph(lg)
    for eid, planes in embeds.items():
        og.embed(eid, planes)
    og.eval()
    nodes = lg.nodes
    picks = [nodes[int(i)] for i in r.choice(len(nodes), size=min(how_many, len(nodes)), replace=False)]
    for n in picks:
        kc.LiveGraph.await_clean_read(lg, n.node_id)
        for s in og.slot_ids(int(n.node_id)):
            want = og.slot(int(n.node_id), s)
            got = lg.slot_data(n.node_id, SlotId(s)).image.planes()
            assert len(got) == len(want), (int(n.node_id), n.node_type, s)
            for c in range(len(want)):
                assert bits_equal(got[c], want[c]), (int(n.node_id), n.node_type, s, c, got[c].shape, want[c].shape)


def random_planes(r, n, size=None):
    w, h = size or [z for z in SIZES if z != (1, 1)][int(r.integers(len(SIZES) - 1))]
    return [r.random((h, w), dtype=np.float32) for _ in range(n)]


@pytest.mark.parametrize("seed", range(48))
def test_edits_keep_every_requested_slot_exact(tex_pro, seed):
    r = np.random.default_rng(7000 + seed)
    graph, embeds = random_graph(6000 + seed, n_ops=8 + seed % 9)
    lg = tex_pro.new_live_graph()
    lg.use_cache = seed % 3 != 0
    lg.set_node_graph(graph)

    def embed(eid):
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tex_pro, embeds[eid])), eid)

    for eid in embeds:
        embed(eid)
    check(tex_pro, lg, embeds, r, 6)
    done = {}
    for step in range(14):
        nodes = lg.nodes
        edges = lg.edges
        kinds = data_kinds(lg, embeds)
        op = int(r.integers(7))
        if op == 0:      # another operator on a Mix node
            mixes = [n for n in nodes if n.node_type.kind == 9]
            if not mixes:
                continue
            n = mixes[int(r.integers(len(mixes)))]
            lg.set_mix_type(n.node_id, list(MixType)[int(r.integers(5))])
        elif op == 1:    # another constant
            vals = [n for n in nodes if n.node_type.kind == 8]
            if not vals:
                continue
            n = vals[int(r.integers(len(vals)))]
            n.node_type = NodeType.Value(float(np.float32(r.random() * 3.0 - 0.5)))
            lg.set_node(n)
        elif op == 2:    # new pixels behind an Embed node: same kind, same or another size
            eid = int(r.integers(len(embeds)))
            same = r.random() < 0.5
            old = embeds[eid]
            embeds[eid] = random_planes(r, len(old), size=(old[0].shape[1], old[0].shape[0]) if same else None)
            lg.replace_embedded(kc.SlotImage.from_planes(tex_pro, embeds[eid]), eid)   # an id in use cannot be embedded again
        elif op == 3:    # rewire one input of a Mix / CombineRgba to another source of the same kind of data
            cands = [e for e in edges if lg.node(e.input_id).node_type.kind in (9, 12)]
            if not cands:
                continue
            e = cands[int(r.integers(len(cands)))]
            want = kinds.get((int(e.output_id), int(e.output_slot)))
            need_static = "gray" if lg.node(e.input_id).node_type.kind == 12 else None
            below = descendants(edges, e.input_id)
            pool = [(nid, s) for (nid, s), k in kinds.items()
                    if k == want and nid not in below and (need_static is None or static_type(lg, nid, s, kinds) in ("gray", "any"))]
            if not pool:
                continue
            src = pool[int(r.integers(len(pool)))]
            lg.disconnect_slot(e.input_id, Side.Input, e.input_slot)
            lg.connect(kc.NodeId(src[0]), e.input_id, SlotId(src[1]), e.input_slot)
        elif op == 4:    # a new Mix over two existing outputs, with a policy and a filter of its own
            pool = sorted(kinds)
            a, b = pool[int(r.integers(len(pool)))], pool[int(r.integers(len(pool)))]
            n = Node.new(NodeType.Mix(list(MixType)[int(r.integers(5))]))
            n.resize_filter = list(ResizeFilter)[int(r.integers(5))]
            n.resize_policy = [ResizePolicy.MostPixels, ResizePolicy.LeastPixels, ResizePolicy.SpecificSlot(SlotId(1))][int(r.integers(3))]
            nid = lg.add_node(n)
            lg.connect(kc.NodeId(a[0]), nid, SlotId(a[1]), SlotId(0))
            lg.connect(kc.NodeId(b[0]), nid, SlotId(b[1]), SlotId(1))
        elif op == 5:    # remove a node nothing depends on (never an Embed: the ids stay dense)
            used = {int(e.output_id) for e in edges}
            leaves = [n for n in nodes if int(n.node_id) not in used and n.node_type.kind not in (6, 8)]
            if len(leaves) < 2:
                continue
            lg.remove_node(leaves[int(r.integers(len(leaves)))].node_id)
        else:            # another filter / policy on a node that resizes
            cands = [n for n in nodes if n.node_type.kind in (9, 10, 11, 12)]
            if not cands:
                continue
            n = cands[int(r.integers(len(cands)))]
            n.resize_filter = list(ResizeFilter)[int(r.integers(5))]
            if r.random() < 0.5:
                w, h = SIZES[int(r.integers(len(SIZES)))]
                n.resize_policy = ResizePolicy.SpecificSize(kc.Size.new(w, h))
            lg.set_node(n)
        done[op] = done.get(op, 0) + 1
        check(tex_pro, lg, embeds, r, 4)
    assert sum(done.values()) >= 6, done
    lg.close()
