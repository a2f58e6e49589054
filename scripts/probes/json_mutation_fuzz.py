"""Mutation fuzz of the graph JSON loader (kc_graph.cu: kc_graph_from_json): damaged documents
must come back as an error or as a graph, never as a crash.  Every graph that does load is
exported and loaded again (the export must be loadable).  Run against the AddressSanitizer
build like scripts/probes/png_mutation_fuzz.py (no GPU needed)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kanter_core_b200._lib import lib  # noqa: E402


def load(text):
    g = C.c_void_p()
    rc = lib.kc_graph_from_json(text, C.byref(g))
    if rc != 0:
        return rc
    out = C.c_void_p()
    assert lib.kc_graph_export_json(g, C.byref(out)) == 0
    again = C.string_at(out)
    lib.kc_free(out)
    lib.kc_graph_destroy(g)
    g2 = C.c_void_p()
    if b"null" in again:        # a Value that overflowed f32 to inf is written as null, as serde_json does; null does not load
        return 0
    if lib.kc_graph_from_json(again, C.byref(g2)) != 0:
        raise AssertionError("export of a loaded graph does not load: %s\n--- input\n%s\n--- export\n%s" % (
            lib.kc_last_error().decode(), text.decode("utf-8", "replace"), again.decode("utf-8", "replace")))
    lib.kc_graph_destroy(g2)
    return 0


def seeds():
    out = [open(os.path.join(ROOT, "tests", "golden", "data", "invert_graph.json"), "rb").read()]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import kanter_core_b200 as kc
    from kanter_core_b200 import MixType, Node, NodeType, ResizeFilter, ResizePolicy, Size, SlotId
    g = kc.NodeGraph.new()
    inner = kc.NodeGraph.from_path(os.path.join(ROOT, "tests", "golden", "data", "invert_graph.json"))
    ids = [g.add_node(Node.new(t)) for t in (
        NodeType.InputGray("a"), NodeType.InputRgba("b"), NodeType.Value(0.25), NodeType.Mix(MixType.Pow),
        NodeType.HeightToNormal, NodeType.SeparateRgba, NodeType.CombineRgba, NodeType.Embed(3),
        NodeType.Image("some/path.png"), NodeType.Write("out.png"), NodeType.Graph(inner),
        NodeType.OutputGray("g"), NodeType.OutputRgba("c"))]
    n = Node.new(NodeType.Mix(MixType.Add))
    n.resize_policy = ResizePolicy.SpecificSize(Size.new(12, 34))
    n.resize_filter = ResizeFilter.Lanczos3
    m = g.add_node(n)
    g.connect(ids[2], m, SlotId(0), SlotId(0))
    g.connect(ids[0], ids[3], SlotId(0), SlotId(1))
    p = C.c_void_p()
    assert lib.kc_graph_export_json(g._h, C.byref(p)) == 0
    out.append(C.string_at(p))
    lib.kc_free(p)
    return out


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    r = np.random.default_rng(2)
    base = seeds()
    tokens = [b"{", b"}", b"[", b"]", b",", b":", b'"', b"null", b"-1", b"1e999", b"4294967296", b'"Mix"', b'"node_type"',
              b"\\", b"\x00", b"\xff", b'"\\u12"', b"[[[[[[[[[[[[[[[[[[[[[[[[[[[[[[[[", b"0.0000000000000000000000000000001"]
    ok = err = 0
    for it in range(iters):
        d = bytearray(base[int(r.integers(len(base)))])
        how = int(r.integers(5))
        if how == 0:
            for _ in range(int(r.integers(1, 5))):
                d[int(r.integers(len(d)))] = int(r.integers(256))
        elif how == 1:
            d = d[:int(r.integers(len(d)))]
        elif how == 2:
            for _ in range(int(r.integers(1, 4))):
                at = int(r.integers(len(d)))
                d[at:at] = tokens[int(r.integers(len(tokens)))]
        elif how == 3:
            a, b = sorted(int(x) for x in r.integers(len(d), size=2))
            del d[a:b]
        else:
            a, b = sorted(int(x) for x in r.integers(len(d), size=2))
            d[a:a] = d[a:b] * int(r.integers(1, 4))
        text = bytes(d).replace(b"\x00", b" ")            # the ABI takes a C string
        rc = load(text)
        ok += rc == 0
        err += rc != 0
    deep = b"[" * 200000                                  # nesting deeper than any stack
    assert load(deep) != 0
    deep = b'{"nodes":' * 100000
    assert load(deep) != 0
    print("json mutation fuzz: %d inputs, %d loaded (and round-tripped), %d rejected, no crash" % (iters, ok, err))


if __name__ == "__main__":
    main()
