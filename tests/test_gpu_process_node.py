"""`process_node` (src/node/node_type.rs:213-248) through the C ABI, one node at a time, the way
the reference's engine calls it (src/engine.rs:288-296): the SlotData of the connected inputs in
edge order, the graph's embedded / input slot data, the node's outputs back -- compared with the
CPU oracle's per-op functions."""
import numpy as np
import pytest

import kanter_core_b200 as kc
import oracle
from kanter_core_b200 import (Edge, MixType, Node, NodeGraph, NodeType, ResizeFilter, ResizePolicy, Size, SlotData, SlotId)
from tests import graphs

pytestmark = pytest.mark.gpu


def rnd(seed, h, w):
    return np.random.default_rng(seed).random((h, w), dtype=np.float32)


def img(tp, planes):
    return kc.SlotImage.from_planes(tp, planes)


def node(nt, nid, **kw):
    n = Node.with_id(nt, nid)
    for k, v in kw.items():
        setattr(n, k, v)
    return n


def test_mix_node_with_two_inputs_and_an_implicit_resize(tex_pro):
    A = [rnd(c, 48, 64) for c in range(4)]
    B = [rnd(10 + c, 24, 32) for c in range(4)]                    # smaller: MostPixels resizes it with the node's filter
    n = node(NodeType.Mix(MixType.Multiply), 7, resize_filter=ResizeFilter.CatmullRom)
    sds = [SlotData.new(1, 0, img(tex_pro, A)), SlotData.new(2, 0, img(tex_pro, B))]
    edges = [Edge(1, 7, 0, 0), Edge(2, 7, 0, 1)]
    out = kc.process_node(tex_pro, n, sds, edges)
    assert len(out) == 1 and int(out[0].node_id) == 7 and int(out[0].slot_id) == 0
    got = out[0].image.planes()
    for c in range(3):
        want = oracle.mix_plane(2, A[c], oracle.resize_plane(B[c], 64, 48, int(ResizeFilter.CatmullRom)))
        assert np.array_equal(got[c].view(np.uint32), want.view(np.uint32))
    assert np.array_equal(got[3], np.ones((48, 64), np.float32))
    # edge order decides which SlotData feeds which slot (assign_slot_ids, node_type.rs:250-267)
    swapped = kc.process_node(tex_pro, node(NodeType.Mix(MixType.Subtract), 7), sds, [Edge(1, 7, 0, 1), Edge(2, 7, 0, 0)])
    a_minus = kc.process_node(tex_pro, node(NodeType.Mix(MixType.Subtract), 7), sds[::-1], [Edge(2, 7, 0, 0), Edge(1, 7, 0, 1)])
    for c in range(3):
        assert np.array_equal(swapped[0].image.planes()[c], a_minus[0].image.planes()[c])


def test_separate_combine_output_are_aliases(tex_pro):
    A = [rnd(20 + c, 16, 16) for c in range(4)]
    src = SlotData.new(1, 0, img(tex_pro, A))
    sep = kc.process_node(tex_pro, node(NodeType.SeparateRgba, 2), [src], [Edge(1, 2, 0, 0)])
    assert [int(s.slot_id) for s in sep] == [0, 1, 2, 3]
    for c in range(4):
        assert not sep[c].image.is_rgba() and np.array_equal(sep[c].image.planes()[0], A[c])
    k0 = tex_pro.stats()["kernel_launches"]
    comb = kc.process_node(tex_pro, node(NodeType.CombineRgba, 3), [sep[2], sep[0]], [Edge(2, 3, 2, 0), Edge(2, 3, 0, 2)])
    outn = kc.process_node(tex_pro, node(NodeType.OutputRgba("out"), 4), comb, [Edge(3, 4, 0, 0)])
    assert tex_pro.stats()["kernel_launches"] == k0               # aliasing only: nothing ran on the device
    got = outn[0].image.planes()
    assert np.array_equal(got[0], A[2]) and np.array_equal(got[2], A[0])
    assert not got[1].any() and np.array_equal(got[3], np.ones((16, 16), np.float32))   # missing G -> 0, missing A -> 1


def test_value_embed_input_and_height_to_normal(tex_pro):
    v = kc.process_node(tex_pro, node(NodeType.Value(0.25), 1), [], [])
    assert v[0].image.size() == Size(1, 1) and v[0].image.planes()[0][0, 0] == np.float32(0.25)
    H = rnd(5, 40, 56)
    emb = SlotData.new(0, 0, img(tex_pro, [H]))
    e = kc.process_node(tex_pro, node(NodeType.Embed(9), 2), [], [], embedded_slot_datas=[(9, emb)])
    assert np.array_equal(e[0].image.planes()[0], H)
    with pytest.raises(kc.TexProError):                            # unknown embedded id
        kc.process_node(tex_pro, node(NodeType.Embed(8), 2), [], [], embedded_slot_datas=[(9, emb)])
    i = kc.process_node(tex_pro, node(NodeType.InputGray("in"), 3), [], [], input_slot_datas=[SlotData.new(3, 0, img(tex_pro, [H]))])
    assert np.array_equal(i[0].image.planes()[0], H)
    n = kc.process_node(tex_pro, node(NodeType.HeightToNormal, 4), e, [Edge(2, 4, 0, 0)])
    want = oracle.height_to_normal(H)
    for c in range(3):
        assert np.array_equal(n[0].image.planes()[c].view(np.uint32), want[c].view(np.uint32))
    with pytest.raises(kc.TexProError) as err:                     # no (Gray) input -> no buffers -> InvalidBufferCount
        kc.process_node(tex_pro, node(NodeType.HeightToNormal, 4), [], [])
    assert err.value.kind == "InvalidBufferCount"


def test_specific_size_policy_and_nested_graph(tex_pro):
    L = rnd(7, 16, 16)
    n = node(NodeType.Mix(MixType.Add), 5, resize_policy=ResizePolicy.SpecificSize(Size.new(40, 24)), resize_filter=ResizeFilter.Lanczos3)
    out = kc.process_node(tex_pro, n, [SlotData.new(1, 0, img(tex_pro, [L]))], [Edge(1, 5, 0, 0)])
    assert out[0].image.size() == Size(40, 24)
    want = oracle.mix_plane(0, oracle.resize_plane(L, 40, 24, int(ResizeFilter.Lanczos3)), np.zeros((24, 40), np.float32))
    assert np.array_equal(out[0].image.planes()[0].view(np.uint32), want.view(np.uint32))
    inner = NodeGraph.from_path(graphs.INVERT_JSON)
    g = node(NodeType.Graph(inner), 6)
    res = kc.process_node(tex_pro, g, [SlotData.new(1, 0, img(tex_pro, [L]))],
                          [Edge(1, 6, 0, int(inner.input_slot_id_with_name("in")))])
    assert [int(r.slot_id) for r in res] == [int(inner.output_slot_id_with_name("out"))]
    assert np.array_equal(res[0].image.planes()[0], (np.float32(1.0) - L).astype(np.float32))


def test_argument_errors(tex_pro):
    with pytest.raises(kc.TexProError):                            # assert_eq!(edges.len(), slot_datas.len()), node_type.rs:221-226
        kc.process_node(tex_pro, node(NodeType.Mix(MixType.Add), 1), [], [Edge(0, 1, 0, 0)])
