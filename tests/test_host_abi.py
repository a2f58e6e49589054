"""CPU-only checks of the host side: the shared library loads and exports every
symbol include/kanter_b200.h declares, the NodeGraph model behaves like
src/node_graph.rs, JSON import/export follows the serde schema, and the error
codes are the TexProError discriminants.  No CUDA device is touched."""
import ctypes as C
import json
import os
import re

import pytest

import kanter_core_b200 as kc
from kanter_core_b200 import (MixType, Node, NodeGraph, NodeType, ResizeFilter, ResizePolicy, Side, Size, SlotId,
                              SlotType, TexProError)
from kanter_core_b200 import _lib
from tests import graphs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "kanter_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = declared_symbols()
    assert len(syms) > 90
    lib = C.CDLL(_lib.LIB_PATH)
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_abi_version_and_error_strings():
    assert _lib.lib.kc_abi_version() == 1
    # Display for TexProError, src/error.rs:37-64
    assert _lib.lib.kc_error_string(4) == b"Invalid number of channels"
    assert _lib.lib.kc_error_string(10) == b"Could not find a `SlotData`"
    assert _lib.lib.kc_error_string(19).startswith(b"Invalid name")


def test_no_gpu_means_loud_failure_not_fallback():
    import subprocess, sys
    # hide every device from a fresh process: context creation must fail with a CUDA error
    code = ("import kanter_core_b200 as kc\n"
            "try:\n    kc.TextureProcessor.new()\n    print('CREATED')\n"
            "except kc.TexProError as e:\n    print('ERR', e.kind)\n")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True).stdout
    assert "ERR Cuda" in out, out


def test_entry_points_that_need_no_device():
    """Argument checking and the process-wide helpers answer without a GPU."""
    from kanter_core_b200._lib import TexProError, call, lib
    n = C.c_int32(-1)
    call("kc_debug_jit_wait", 10, C.byref(n))                     # nothing is compiling
    assert n.value == 0
    for fn, args in (("kc_context_set_resize_unclamped", (None, 1)), ("kc_context_set_math_mode", (None, 0)),
                     ("kc_context_synchronize", (None,)), ("kc_live_graph_create", (None, None))):
        with pytest.raises(TexProError) as e:
            call(fn, *args)
        assert e.value.code == 101, fn                           # KC_ERR_INVALID_ARGUMENT, not a crash
    ctx = C.c_void_p()
    with pytest.raises(TexProError) as e:                         # a caller-owned stream does not conjure a device either
        call("kc_context_create_on_stream", 0, None, None, C.byref(ctx))
    assert e.value.kind == "Cuda" and not ctx.value
    assert lib.kc_context_destroy(None) == 0                      # destroying nothing is fine


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "kanter_core_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("CPU oracle", "").replace("the oracle", "") or f == "_lib.py", (dirpath, f)


# ---- NodeGraph --------------------------------------------------------------------
def test_slot_tables_match_the_reference():
    # src/node/node_type.rs:141-211
    mix = Node.new(NodeType.Mix(MixType.Add))
    assert [(s.name, int(s.slot_id), s.slot_type) for s in mix.input_slots()] == [
        ("left", 0, SlotType.GrayOrRgba), ("right", 1, SlotType.GrayOrRgba)]
    assert [(s.name, int(s.slot_id), s.slot_type) for s in mix.output_slots()] == [("output", 0, SlotType.GrayOrRgba)]
    comb = Node.new(NodeType.CombineRgba)
    assert [s.name for s in comb.input_slots()] == ["red", "green", "blue", "alpha"]
    sep = Node.new(NodeType.SeparateRgba)
    assert [int(s.slot_id) for s in sep.output_slots()] == [0, 1, 2, 3]
    assert Node.new(NodeType.OutputGray("o")).output_slots() == []
    assert Node.new(NodeType.Value(1.0)).input_slots() == []
    assert Node.new(NodeType.HeightToNormal).input_slots()[0].slot_type == SlotType.Gray
    assert Node.new(NodeType.HeightToNormal).output_slots()[0].slot_type == SlotType.Rgba
    assert mix.input_slot_with_name("right").slot_id == 1
    with pytest.raises(TexProError) as e:
        mix.input_slot_with_name("nope")
    assert e.value.kind == "InvalidName"


def test_graph_node_slots_are_inner_node_ids():
    # NodeGraph::input_slots/output_slots, src/node_graph.rs:299-330
    inner = NodeGraph.from_path(graphs.INVERT_JSON)
    n = Node.new(NodeType.Graph(inner))
    assert [(s.name, int(s.slot_id), s.slot_type) for s in n.input_slots()] == [("in", 808182335, SlotType.Gray)]
    assert [(s.name, int(s.slot_id), s.slot_type) for s in n.output_slots()] == [("out", 3948812722, SlotType.Gray)]
    assert inner.input_slot_id_with_name("in") == 808182335
    assert inner.output_slot_id_with_name("out") == 3948812722


def test_connect_rules():
    g = NodeGraph.new()
    v = g.add_node(Node.new(NodeType.Value(0.0)))
    m = g.add_node(Node.new(NodeType.Mix(MixType.default())))
    # connect_invalid_slot, tests/integration_tests.rs:786-810
    g.connect(v, m, SlotId(0), SlotId(0))
    g.connect(v, m, SlotId(0), SlotId(1))
    with pytest.raises(TexProError) as e:
        g.connect(v, m, SlotId(0), SlotId(2))
    assert e.value.kind == "InvalidSlotId"
    # connecting to an occupied input replaces the edge (src/node_graph.rs:435)
    v2 = g.add_node(Node.new(NodeType.Value(1.0)))
    g.connect(v2, m, SlotId(0), SlotId(0))
    assert [e_._tuple() for e_ in g.edges if e_.input_slot == 0] == [(int(v2), int(m), 0, 0)]
    # try_connect refuses an occupied slot
    with pytest.raises(TexProError) as e:
        g.try_connect(v, m, SlotId(0), SlotId(0))
    assert e.value.kind == "SlotOccupied"
    # the same edge twice is an InvalidEdge only if it survives the implicit disconnect: it does not
    g.connect(v2, m, SlotId(0), SlotId(0))
    # unknown node
    with pytest.raises(TexProError) as e:
        g.connect(999, m, SlotId(0), SlotId(0))
    assert e.value.kind == "InvalidNodeId"


def test_wrong_slot_type():
    # tests/integration_tests.rs:1330-1347 (Rgba output into a Gray input)
    g = NodeGraph.new()
    i = g.add_node(Node.new(NodeType.Image("x.png")))
    o = g.add_node(Node.new(NodeType.OutputGray("out")))
    with pytest.raises(TexProError) as e:
        g.connect(i, o, SlotId(0), SlotId(0))
    assert e.value.kind == "InvalidSlotType"
    # GrayOrRgba fits both (a Mix output into OutputGray: mix_node_single_input)
    m = g.add_node(Node.new(NodeType.Mix(MixType.Add)))
    g.connect(m, o, SlotId(0), SlotId(0))


def test_ids_names_and_removal():
    g = NodeGraph.new()
    a = g.add_node(Node.new(NodeType.OutputRgba("out")))
    b = g.add_node(Node.new(NodeType.OutputRgba("out")))
    c = g.add_node(Node.new(NodeType.OutputGray("out")))
    d = g.add_node(Node.new(NodeType.InputGray("")))
    assert (a, b, c, d) == (0, 1, 2, 3)
    names = {int(n.node_id): n.node_type.payload for n in g.nodes}
    # avoid_name_collision, src/node_graph.rs:141-164; empty names become "untitled"
    assert names == {0: "out", 1: "out_0", 2: "out_1", 3: "untitled"}
    assert g.output_ids() == [0, 1, 2] and g.input_ids() == [3]
    with pytest.raises(TexProError):
        g.add_node_with_id(Node.with_id(NodeType.Value(1.0), 2))
    g.add_node_with_id(Node.with_id(NodeType.Value(1.0), 10))
    v = g.add_node(Node.new(NodeType.Value(2.0)))
    assert v == 4  # new_id keeps counting from where it was and skips ids in use
    # remove_node drops the edges too (remove_node test, :773-784)
    g.connect(10, 2, SlotId(0), SlotId(0))
    g.remove_node(10)
    assert not g.edges and 10 not in g.node_ids()
    with pytest.raises(TexProError) as e:
        g.disconnect_slot(2, Side.Input, SlotId(0))
    assert e.value.kind == "SlotNotOccupied"


def test_json_import_matches_fixture_and_roundtrips():
    g = NodeGraph.from_path(graphs.INVERT_JSON)
    want = json.load(open(graphs.INVERT_JSON))
    assert json.loads(g.export_json_string()) == want
    # byte-for-byte the serde_json pretty printing of the fixture
    assert g.export_json_string() == open(graphs.INVERT_JSON).read().rstrip("\n")
    # from_path restarts the id counter after the largest id (src/node_graph.rs:36-43)
    assert g.add_node(Node.new(NodeType.Value(0.0))) == 3948812723


def test_json_every_variant_roundtrips(tmp_path):
    inner = NodeGraph.from_path(graphs.INVERT_JSON)
    g = NodeGraph.new()
    nodes = [NodeType.InputGray("a"), NodeType.InputRgba("b"), NodeType.OutputGray("c"), NodeType.OutputRgba("d"),
             NodeType.Graph(inner), NodeType.Image("dir/some \"file\".png"), NodeType.Embed(7), NodeType.Write("out.png"),
             NodeType.Value(0.33), NodeType.Mix(MixType.Pow), NodeType.HeightToNormal, NodeType.SeparateRgba,
             NodeType.CombineRgba]
    ids = []
    for i, t in enumerate(nodes):
        n = Node.new(t)
        n.resize_policy = [ResizePolicy.MostPixels, ResizePolicy.LeastPixels, ResizePolicy.LargestAxes,
                           ResizePolicy.SmallestAxes, ResizePolicy.SpecificSlot(SlotId(2)),
                           ResizePolicy.SpecificSize(Size(640, 480))][i % 6]
        n.resize_filter = ResizeFilter(i % 5)
        ids.append(g.add_node(n))
    g.connect(ids[8], ids[9], SlotId(0), SlotId(1))
    p = tmp_path / "g.json"
    g.export_json(str(p))
    d = json.load(open(p))
    assert d["nodes"][8]["node_type"] == {"Value": 0.33}
    assert d["nodes"][10]["node_type"] == "HeightToNormal"
    assert d["nodes"][4]["node_type"]["Graph"] == json.load(open(graphs.INVERT_JSON))
    assert d["nodes"][4]["resize_policy"] == {"SpecificSlot": 2}
    assert d["nodes"][5]["resize_policy"] == {"SpecificSize": {"width": 640, "height": 480}}
    assert d["edges"] == [{"output_id": 8, "input_id": 9, "output_slot": 0, "input_slot": 1}]
    g2 = NodeGraph.from_path(str(p))
    assert g2.export_json_string() == g.export_json_string()
    assert [repr(n.resize_policy) for n in g2.nodes] == [repr(n.resize_policy) for n in g.nodes]


def test_json_errors():
    with pytest.raises(TexProError) as e:
        NodeGraph.from_json("{\"nodes\": [")
    assert e.value.kind == "Io"
    with pytest.raises(TexProError):
        NodeGraph.from_json('{"nodes":[{"node_id":1,"node_type":"Bogus","resize_policy":"MostPixels","resize_filter":"Triangle"}],"edges":[]}')
    with pytest.raises(TexProError):
        NodeGraph.from_path("/nonexistent/graph.json")
    # what serde_json refuses is refused here: numbers outside f64, tokens that are not JSON numbers,
    # u32 fields that are negative / fractional / too large, nesting deeper than 128
    node = '{"node_id":%s,"node_type":{"Value":%s},"resize_policy":"MostPixels","resize_filter":"Triangle"}'
    doc = '{"nodes":[' + node + '],"edges":[%s]}'
    assert len(NodeGraph.from_json(doc % ("7", "0.5", "")).nodes) == 1
    for nid, val, edge in (("7", "1e999", ""), ("7", "inf", ""), ("7", "+1", ""), ("7", ".5", ""), ("7", "0x10", ""),
                           ("-1", "0.5", ""), ("1.5", "0.5", ""), ("4294967296", "0.5", ""), ('"7"', "0.5", ""),
                           ("7", "0.5", '{"output_id":7,"input_id":"x","output_slot":0,"input_slot":0}'),
                           ("7", "0.5", "7")):
        with pytest.raises(TexProError) as e:
            NodeGraph.from_json(doc % (nid, val, edge))
        assert e.value.kind == "Io", (nid, val, edge)
    with pytest.raises(TexProError):
        NodeGraph.from_json("[" * 100000)


def test_set_mix_type_and_clone_independence():
    g = NodeGraph.new()
    m = g.add_node(Node.new(NodeType.Mix(MixType.Add)))
    v = g.add_node(Node.new(NodeType.Value(1.0)))
    c = g.clone()
    g.set_mix_type(m, MixType.Divide)
    assert g.node(m).node_type.payload == MixType.Divide and c.node(m).node_type.payload == MixType.Add
    with pytest.raises(TexProError):
        g.set_mix_type(v, MixType.Add)


def test_calculate_size_policies_without_pixels():
    """calculate_size (src/shared.rs:61-139) through the ABI, on constant descriptors
    (no device work): ties in MostPixels go to the LAST input, in LeastPixels to the FIRST."""
    from kanter_core_b200._lib import call, kc_edge, kc_image, kc_slot_data
    lib = _lib.lib

    def img(w, h):
        im = kc_image()
        # a descriptor that is only measured belongs to no context
        pl = C.c_void_p()
        call("kc_plane_from_value", None, w, h, 0.0, C.byref(pl))
        im.kind = 0
        im.width, im.height = w, h
        im.planes[0] = pl
        return im

    sizes = [(128, 64), (64, 128), (32, 32)]
    sds = (kc_slot_data * 3)()
    edges = (kc_edge * 3)()
    for i, (w, h) in enumerate(sizes):
        sds[i].node_id, sds[i].slot_id, sds[i].image = 10 + i, 0, img(w, h)
        edges[i] = kc_edge(10 + i, 99, 0, i)

    def size(policy, slot=0, pw=0, ph=0, n=3):
        w, h = C.c_uint32(), C.c_uint32()
        call("kc_calculate_size", sds, n, edges, n, policy, slot, pw, ph, C.byref(w), C.byref(h))
        return (w.value, h.value)

    assert size(0) == (64, 128)            # MostPixels: tie -> last
    assert size(1) == (32, 32)             # LeastPixels
    assert size(1, n=2) == (128, 64)       # LeastPixels: tie -> first
    assert size(2) == (128, 128)           # LargestAxes
    assert size(3) == (32, 32)             # SmallestAxes
    assert size(4, slot=1) == (64, 128)    # SpecificSlot(1)
    assert size(4, slot=7) == (128, 64)    # absent slot -> lowest connected slot
    assert size(5, pw=300, ph=200) == (300, 200)
    assert size(0, n=0) == (1, 1)          # MostPixels with no inputs


def test_node_graph_bookkeeping_surface():
    """can_connect / connected_edges / new_id / rename_output_node, src/node_graph.rs:86-96,232-270,376-393,518-537."""
    import kanter_core_b200 as kc
    from kanter_core_b200 import MixType, Node, NodeGraph, NodeType, Side, SlotId
    g = NodeGraph.new()
    a = g.add_node(Node.new(NodeType.Value(0.5)))
    b = g.add_node(Node.new(NodeType.Value(0.25)))
    m = g.add_node(Node.new(NodeType.Mix(MixType.Add)))
    o1 = g.add_node(Node.new(NodeType.OutputGray("out")))
    o2 = g.add_node(Node.new(NodeType.OutputGray("out")))        # de-collided on insert
    assert g.output_names() == ["out", "out_0"]
    g.can_connect(a, m, SlotId(0), SlotId(0))
    g.connect(a, m, SlotId(0), SlotId(0))
    with pytest.raises(kc.TexProError) as e:
        g.can_connect(b, m, SlotId(0), SlotId(0))
    assert e.value.kind == "SlotOccupied"
    with pytest.raises(kc.TexProError) as e:
        g.can_connect(b, m, SlotId(0), SlotId(7))
    assert e.value.kind == "InvalidSlotId"
    g.connect(b, m, SlotId(0), SlotId(1))
    g.connect(m, o1, SlotId(0), SlotId(0))
    g.connect(m, o2, SlotId(0), SlotId(0))
    out_edges = g.connected_edges(m, Side.Output, SlotId(0))
    assert sorted(int(x.input_id) for x in out_edges) == sorted([int(o1), int(o2)])
    assert [int(x.output_id) for x in g.connected_edges(m, Side.Input, SlotId(1))] == [int(b)]
    with pytest.raises(kc.TexProError) as e:
        g.connected_edges(a, Side.Input, SlotId(0))
    assert e.value.kind == "SlotNotOccupied"
    assert g.rename_output_node(o2, "out") == "out_0"            # collides with o1's name again
    assert g.output_names() == ["out", "out_0"]
    assert g.rename_output_node(o2, "normal") == "out_0"
    assert g.output_names() == ["out", "normal"]
    with pytest.raises(kc.TexProError) as e:
        g.rename_output_node(m, "x")
    assert e.value.kind == "InvalidNodeType"
    nid = g.new_id()
    assert int(nid) not in [int(n) for n in g.node_ids()]
    assert sorted(int(x) for x in g.get_children_recursive(a)) == sorted([int(m), int(o1), int(o2)])


def test_kernel_generator_compiles_with_nvrtc():
    """kc_jit.cu: the device code embedded in the library (kc_tape.h + kc_tile_vm.cuh) plus the
    straight-line program generated from a tape compile for sm_100a with NVRTC -- no GPU involved.
    Tapes: every arithmetic op on source / temporary / immediate operands, stores, both exports."""
    import ctypes as C
    from kanter_core_b200._lib import call
    LD, ADD, SUB, RSUB, MUL, DIV, RDIV, POW, RPOW, ST_TMP, ST_OUT, PACK_RGBA, PACK_GRAY = range(13)
    S = lambda k: k          # source k
    T = lambda j: 8 + j      # temporary j
    IMM = 14
    w = lambda op, arg=0: op | (arg << 8)
    tapes = {
        "all_ops": [w(LD, S(0)), w(ADD, S(1)), w(SUB, IMM), w(RSUB, S(2)), w(ST_TMP, 0), w(MUL, S(3)), w(DIV, T(0)), w(RDIV, IMM),
                    w(POW, S(1)), w(RPOW, T(0)), w(ST_OUT, 0), w(LD, T(0)), w(ST_OUT, 1)],
        "export_rgba": [w(LD, S(0)), w(ST_TMP, 0), w(LD, S(1)), w(ST_TMP, 1), w(LD, S(2)), w(MUL, S(3)), w(ST_TMP, 2), w(LD, IMM), w(PACK_RGBA, 1)],
        "export_gray": [w(LD, S(0)), w(POW, IMM), w(PACK_GRAY, 0)],
    }
    for name, tape in tapes.items():
        arr = (C.c_uint32 * len(tape))(*tape)
        for exact, v, ctas in ((1, 4, 3), (0, 1, 2)):
            n = C.c_size_t()
            call("kc_debug_jit_compile", arr, len(tape), exact, v, ctas, C.byref(n))
            assert n.value > 10000, (name, exact, v)


def test_graph_queries_of_the_reference_surface():
    """edge_indices_node / edge_indices_slot / slot_occupied / input_nodes / output_nodes /
    set_image_node_path (src/node_graph.rs:65-83,191-203,351-374,449-460)."""
    g = NodeGraph.new()
    i = g.add_node(Node.new(NodeType.InputGray("in")))
    im = g.add_node(Node.new(NodeType.Image("a.png")))
    m = g.add_node(Node.new(NodeType.Mix(MixType.Add)))
    o = g.add_node(Node.new(NodeType.OutputGray("out")))
    g.connect(i, m, SlotId(0), SlotId(0))
    g.connect(im, m, SlotId(0), SlotId(1))
    g.connect(m, o, SlotId(0), SlotId(0))
    assert g.edge_indices_node(m) == [0, 1, 2] and g.edge_indices_node(i) == [0]
    assert g.edge_indices_slot(m, Side.Input, SlotId(1)) == [1] and g.edge_indices_slot(m, Side.Output, SlotId(0)) == [2]
    assert g.slot_occupied(m, Side.Input, SlotId(0)) and not g.slot_occupied(o, Side.Output, SlotId(0))
    assert [int(n.node_id) for n in g.input_nodes()] == [int(i)] and [int(n.node_id) for n in g.output_nodes()] == [int(o)]
    with pytest.raises(TexProError):
        g.edge_indices_node(kc.NodeId(77))
    g.set_image_node_path(im, "b.png")
    assert g.node(im).node_type.payload == "b.png"
    with pytest.raises(TexProError):
        g.set_image_node_path(m, "c.png")
    assert kc.NodeId(5).as_usize() == 5 and SlotId(3).as_usize() == 3


def test_image_handles_without_a_device():
    """SlotImage::from_self / SlotData::from_self / in_memory (src/slot_image.rs:104-114, src/slot_data.rs:62-78) on
    constant descriptors, which need no device: a copy is another reference on the same immutable planes."""
    from kanter_core_b200._lib import call, kc_image
    im = kc_image()
    im.kind, im.width, im.height = 1, 7, 5
    for c in range(4):
        pl = C.c_void_p()
        call("kc_plane_from_value", None, 7, 5, 0.25 * c, C.byref(pl))
        im.planes[c] = pl
    img = kc.SlotImage(None, im)
    assert img.is_rgba() and img.size() == Size(7, 5) and img.in_memory()
    twin = img.from_self()
    assert twin._im.planes[2] == img._im.planes[2]              # the same plane object, one more reference
    sd = kc.SlotData.new(4, 2, img).from_self()
    assert (int(sd.node_id), int(sd.slot_id)) == (4, 2) and sd.in_memory() and sd.size() == Size(7, 5)
    v, is_c = C.c_float(), C.c_int32()
    del img, twin                                               # two of the three references go; the plane must survive
    call("kc_plane_is_constant", sd.image._im.planes[3], C.byref(is_c), C.byref(v))
    assert is_c.value == 1 and v.value == 0.75
    import numpy as np
    a = np.zeros((2, 2), np.float32)
    for make, n in ((kc.SlotImage.from_buffers_rgb, 4), (kc.SlotImage.from_buffers_rgba, 3)):   # src/slot_image.rs:66-69,90-93
        with pytest.raises(TexProError) as e:
            make(None, [a] * n)
        assert e.value.kind == "InvalidBufferCount"


def test_context_less_constant_planes_can_be_released():
    """kc_plane_from_value(ctx = NULL) is documented as legal; releasing such a plane (or an image made only of
    them) used to dereference the NULL context / return early without releasing (advisor finding, round 1)."""
    from kanter_core_b200._lib import call, kc_image
    pl = C.c_void_p()
    call("kc_plane_from_value", None, 3, 2, 0.5, C.byref(pl))
    call("kc_plane_retain", pl)
    call("kc_plane_release", pl)
    call("kc_plane_release", pl)            # last reference: the descriptor is gone, no context touched
    call("kc_plane_release", None)          # NULL stays a no-op
    im = kc_image()
    im.kind, im.width, im.height = 1, 3, 2
    for c in range(4):
        p = C.c_void_p()
        call("kc_plane_from_value", None, 3, 2, 0.25 * c, C.byref(p))
        im.planes[c] = p
    call("kc_image_release", C.byref(im))
    assert all(not im.planes[c] for c in range(4)), "the handles must be cleared once released"
    call("kc_image_release", C.byref(im))   # releasing an empty image is a no-op


def test_propagate_priority():
    """The reference's unit test `propagate_priority` (src/priority.rs:181-245) on the C++ graph model:

         1---2---+
                 4---5          own priorities 3, -10, 8, 5, 0
             3---+

    a node that has a high priority needs all its parents to have the same high priority."""
    from kanter_core_b200 import MixType, Node, NodeGraph, NodeType, SlotId
    g = NodeGraph.new()
    own = [3, -10, 8, 5, 0]
    ids = []
    for v in own:
        nid = g.add_node(Node.new(NodeType.Mix(MixType.default())))
        g.node(nid).priority.set_priority(v)
        ids.append(nid)
    n1, n2, n3, n4, n5 = ids
    g.connect(n1, n2, SlotId(0), SlotId(0))
    g.connect(n2, n4, SlotId(0), SlotId(0))
    g.connect(n3, n4, SlotId(0), SlotId(1))
    g.connect(n4, n5, SlotId(0), SlotId(0))
    want = {n1: 5, n2: 5, n3: 8, n4: 5, n5: 0}      # assert_priority(node, expected) x 5, :215-238
    for nid, v in zip(ids, own):
        p = g.node(nid).priority
        assert p.priority() == v
        assert p.propagated_priority() == want[nid], nid
    # the reference pops its priority-sorted list: 3 (8), then 4, 1, 2 by propagated 5, then 5 (0): same order here
    order = sorted(ids, key=lambda n: g.node(n).priority.propagated_priority(), reverse=True)
    assert order[0] == n3 and order[-1] == n5
    # lowering node 4 lets its ancestors fall back to their own priorities; a priority set on a free node travels
    g.node(n4).priority.set_priority(-20)
    assert [g.node(n).priority.propagated_priority() for n in ids] == [3, 0, 8, 0, 0]
    fresh = Node.new(NodeType.Value(1.0))
    fresh.priority.set_priority(7)
    f = g.add_node(fresh)
    assert g.node(f).priority.priority() == 7
    # set_node replaces the node's data, not its scheduling state; clones keep it; JSON does not carry it (#[serde(skip)])
    node = g.node(n3)
    node.resize_filter = kc.ResizeFilter.Nearest
    assert g.clone().node(n3).priority.priority() == 8
    assert NodeGraph.from_json(g.export_json_string()).node(n3).priority.priority() == 0


def test_concurrent_section_entry_points_reject_a_null_context():
    """kc_context_concurrent_begin / _end without a context: an error code and a message, no crash (no GPU needed)."""
    from kanter_core_b200._lib import call
    for name, args in (("kc_context_concurrent_begin", (None, 2)), ("kc_context_concurrent_end", (None,))):
        with pytest.raises(Exception) as e:
            call(name, *args)
        assert "NULL" in str(e.value)
