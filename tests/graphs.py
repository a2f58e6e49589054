"""The graphs of the reference's own integration tests (tests/integration_tests.rs),
rebuilt through the mirrored NodeGraph API.  Each builder returns a Case: the
NodeGraph, the node whose slot 0 the test reads, the golden it is compared with
and the expected size.  Both the CPU oracle tests and the GPU parity tests
consume these, so the two sides are guaranteed to evaluate the same graph.
"""
import os

import numpy as np
from PIL import Image as PILImage

import kanter_core_b200 as kc
from kanter_core_b200 import MixType, Node, NodeGraph, NodeType, ResizePolicy, Size, SlotId

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")
CMP = os.path.join(DATA, "test_compare")
IMAGE_1 = os.path.join(DATA, "image_1.png")
IMAGE_2 = os.path.join(DATA, "image_2.png")
HEART_128 = os.path.join(DATA, "heart_128.png")
HEART_256 = os.path.join(DATA, "heart_256.png")
HEART_WIDE = os.path.join(DATA, "heart_wide.png")
HEART_TALL = os.path.join(DATA, "heart_tall.png")
HEART_110 = os.path.join(DATA, "heart_110.png")
CLOUDS = os.path.join(DATA, "clouds.png")
INVERT_JSON = os.path.join(DATA, "invert_graph.json")


def decode(path):
    """8-bit samples as image::open(..).as_flat_samples_u8() yields them."""
    im = PILImage.open(path)
    assert im.mode in ("L", "LA", "RGB", "RGBA"), im.mode
    a = np.asarray(im, dtype=np.uint8)
    return a if a.ndim == 3 else a[:, :, None]


def golden(name):
    a = decode(os.path.join(CMP, name))
    assert a.shape[2] == 4
    return a


class Case:
    def __init__(self, graph, node, golden_name=None, size=(256, 256), embeds=None, golden_path=None):
        self.graph, self.node, self.golden_name, self.size = graph, node, golden_name, size
        self.embeds = embeds or {}       # embed id -> path of the image embedded (as an Rgba image)
        self.golden_path = golden_path

    def expected(self):
        if self.golden_path:
            return decode(self.golden_path)
        return golden(self.golden_name)

    def image_nodes(self):
        return {int(n.node_id): n.node_type.payload for n in self.graph.nodes if n.node_type.kind == 5}


def _img(g, path):
    return g.add_node(Node.new(NodeType.Image(path)))


def input_output():  # :52-95, compared against the INPUT
    g = NodeGraph.new()
    i = _img(g, IMAGE_2)
    o = g.add_node(Node.new(NodeType.OutputRgba("out")))
    g.connect(i, o, SlotId(0), SlotId(0))
    return Case(g, o, golden_path=IMAGE_2)


def mix_node_single_input():  # :494-523
    g = NodeGraph.new()
    v = _img(g, IMAGE_2)
    m = g.add_node(Node.new(NodeType.Mix(MixType.Add)))
    o = g.add_node(Node.new(NodeType.OutputGray("out")))
    g.connect(v, m, SlotId(0), SlotId(0))
    g.connect(m, o, SlotId(0), SlotId(0))
    return Case(g, o, "mix_node_single_input.png")


def mix_node_single_input_2():  # :525-553
    g = NodeGraph.new()
    v = _img(g, IMAGE_2)
    m = g.add_node(Node.new(NodeType.Mix(MixType.Subtract)))
    o = g.add_node(Node.new(NodeType.OutputGray("out")))
    g.connect(v, m, SlotId(0), SlotId(1))
    g.connect(m, o, SlotId(0), SlotId(0))
    return Case(g, o, "mix_node_single_input_2.png")


def embedded_node_data():  # :567-617 (second graph; embed 0 holds image_1 passed through a first graph)
    g = NodeGraph.new()
    o = g.add_node(Node.new(NodeType.OutputRgba("out")))
    i = g.add_node(Node.new(NodeType.Embed(0)))
    g.connect(i, o, SlotId(0), SlotId(0))
    return Case(g, o, "embedded_node_data.png", embeds={0: IMAGE_1})


def separate_node():  # :619-674
    g = NodeGraph.new()
    i1 = _img(g, IMAGE_1)
    s1 = g.add_node(Node.new(NodeType.SeparateRgba))
    i2 = _img(g, IMAGE_2)
    s2 = g.add_node(Node.new(NodeType.SeparateRgba))
    o = g.add_node(Node.new(NodeType.OutputRgba("out")))
    c = g.add_node(Node.new(NodeType.CombineRgba))
    g.connect(i1, s1, SlotId(0), SlotId(0))
    g.connect(i2, s2, SlotId(0), SlotId(0))
    g.connect(s1, c, SlotId(3), SlotId(0))
    g.connect(s1, c, SlotId(1), SlotId(1))
    g.connect(s2, c, SlotId(2), SlotId(2))
    g.connect(s2, c, SlotId(3), SlotId(3))
    g.connect(c, o, SlotId(0), SlotId(0))
    return Case(g, o, "mix_images.png")


def irregular_sizes():  # :676-738
    g = NodeGraph.new()
    i1 = _img(g, HEART_128)
    i2 = _img(g, HEART_110)
    m = g.add_node(Node.new(NodeType.Mix(MixType.default())))
    o = g.add_node(Node.new(NodeType.OutputRgba("out")))
    g.connect(i1, m, SlotId(0), SlotId(0))
    g.connect(i2, m, SlotId(0), SlotId(1))
    g.connect(m, o, SlotId(0), SlotId(0))
    return Case(g, o, "irregular_sizes.png", size=(128, 128))


def value_node():  # :812-846 (reads the Combine node itself)
    g = NodeGraph.new()
    vals = [g.add_node(Node.new(NodeType.Value(v))) for v in (0.0, 0.33, 0.66, 1.0)]
    n = Node.new(NodeType.CombineRgba)
    n.resize_policy = ResizePolicy.SpecificSize(Size.new(256, 256))
    c = g.add_node(n)
    for i, v in enumerate(vals):
        g.connect(v, c, SlotId(0), SlotId(i))
    return Case(g, c, "value_node.png")


def _invert_graph():  # :996-1023
    ig = NodeGraph.new()
    white = ig.add_node(Node.new(NodeType.Value(1.0)))
    nin = ig.add_node(Node.new(NodeType.InputGray("in")))
    sub = ig.add_node(Node.new(NodeType.Mix(MixType.Subtract)))
    nout = ig.add_node(Node.new(NodeType.OutputGray("out")))
    ig.connect(white, sub, SlotId(0), SlotId(0))
    ig.connect(nin, sub, SlotId(0), SlotId(1))
    ig.connect(sub, nout, SlotId(0), SlotId(0))
    return ig


def _graph_node_case(inner, golden_name, gray):
    in_slot = inner.input_slot_id_with_name("in")
    out_slot = inner.output_slot_id_with_name("out")
    g = NodeGraph.new()
    img = _img(g, IMAGE_2)
    if gray:
        gn = g.add_node(Node.new(NodeType.Graph(inner)))
        sep = g.add_node(Node.new(NodeType.SeparateRgba))
        o = g.add_node(Node.new(NodeType.OutputGray("out")))
        g.connect(img, sep, SlotId(0), SlotId(0))
        g.connect(sep, gn, SlotId(0), in_slot)
    else:
        gn = g.add_node(Node.new(NodeType.Graph(inner)))
        o = g.add_node(Node.new(NodeType.OutputRgba("out")))
        g.connect(img, gn, SlotId(0), in_slot)
    g.connect(gn, o, out_slot, SlotId(0))
    return Case(g, o, golden_name)


def invert_graph_node():  # :993-1071
    return _graph_node_case(_invert_graph(), "invert_graph_node.png", gray=True)


def invert_graph_node_import():  # :1108-1160
    return _graph_node_case(NodeGraph.from_path(INVERT_JSON), "invert_graph_node_import.png", gray=True)


def graph_node_rgba():  # :1207-1262
    ng = NodeGraph.new()
    i = ng.add_node(Node.new(NodeType.InputRgba("in")))
    o = ng.add_node(Node.new(NodeType.OutputRgba("out")))
    ng.connect(i, o, SlotId(0), SlotId(0))
    return _graph_node_case(ng, "graph_node_rgba.png", gray=False)


def graph_node_gray():  # :1264-1328
    ng = NodeGraph.new()
    i = ng.add_node(Node.new(NodeType.InputGray("in")))
    o = ng.add_node(Node.new(NodeType.OutputGray("out")))
    ng.connect(i, o, SlotId(0), SlotId(0))
    return _graph_node_case(ng, "graph_node_gray.png", gray=True)


def height_to_normal_node():  # :1349-1384
    g = NodeGraph.new()
    i = _img(g, CLOUDS)
    s = g.add_node(Node.new(NodeType.SeparateRgba))
    h = g.add_node(Node.new(NodeType.HeightToNormal))
    o = g.add_node(Node.new(NodeType.OutputRgba("out")))
    g.connect(i, s, SlotId(0), SlotId(0))
    g.connect(s, h, SlotId(0), SlotId(0))
    g.connect(h, o, SlotId(0), SlotId(0))
    return Case(g, o, "height_to_normal_node.png")


def mix_node_gray(mix_type, name):  # :1439-1475
    g = NodeGraph.new()
    i = _img(g, IMAGE_2)
    s = g.add_node(Node.new(NodeType.SeparateRgba))
    m = g.add_node(Node.new(NodeType.Mix(mix_type)))
    o = g.add_node(Node.new(NodeType.OutputGray("out")))
    g.connect(i, s, SlotId(0), SlotId(0))
    g.connect(s, m, SlotId(0), SlotId(0))
    g.connect(s, m, SlotId(1), SlotId(1))
    g.connect(m, o, SlotId(0), SlotId(0))
    return Case(g, o, name)


def mix_node_rgba(mix_type, name):  # :1477-1510
    g = NodeGraph.new()
    i1 = _img(g, IMAGE_1)
    i2 = _img(g, IMAGE_2)
    m = g.add_node(Node.new(NodeType.Mix(mix_type)))
    o = g.add_node(Node.new(NodeType.OutputRgba("out")))
    g.connect(i1, m, SlotId(0), SlotId(0))
    g.connect(i2, m, SlotId(0), SlotId(1))
    g.connect(m, o, SlotId(0), SlotId(0))
    return Case(g, o, name)


def resize_policy_case(policy, path_1, path_2):  # :848-892 (size only; reads the Mix node)
    g = NodeGraph.new()
    i1 = _img(g, path_1)
    i2 = _img(g, path_2)
    n = Node.new(NodeType.Mix(MixType.default()))
    n.resize_policy = policy
    m = g.add_node(n)
    g.connect(i1, m, SlotId(0), SlotId(0))
    g.connect(i2, m, SlotId(0), SlotId(1))
    return Case(g, m)


RESIZE_POLICY_CASES = [  # :894-949
    ("least_pixels", ResizePolicy.LeastPixels, HEART_128, HEART_256, (128, 128)),
    ("largest_axes", ResizePolicy.LargestAxes, HEART_WIDE, HEART_TALL, (128, 128)),
    ("smallest_axes", ResizePolicy.SmallestAxes, HEART_WIDE, HEART_TALL, (64, 64)),
    ("most_pixels", ResizePolicy.MostPixels, HEART_128, HEART_256, (256, 256)),
    ("specific_size", ResizePolicy.SpecificSize(Size.new(256, 256)), HEART_128, HEART_WIDE, (256, 256)),
    ("specific_slot_1", ResizePolicy.SpecificSlot(SlotId(1)), HEART_128, HEART_WIDE, (128, 64)),
    ("specific_slot_2", ResizePolicy.SpecificSlot(SlotId(2)), HEART_128, HEART_WIDE, (128, 128)),
]

MIX_NAMES = [(MixType.Add, "add"), (MixType.Subtract, "subtract"), (MixType.Multiply, "multiply"),
             (MixType.Divide, "divide"), (MixType.Pow, "pow")]

# name -> builder, for every golden check of the reference's test-suite
GOLDEN_CASES = {
    "input_output": input_output,
    "mix_node_single_input": mix_node_single_input,
    "mix_node_single_input_2": mix_node_single_input_2,
    "embedded_node_data": embedded_node_data,
    "separate_node": separate_node,
    "irregular_sizes": irregular_sizes,
    "value_node": value_node,
    "invert_graph_node": invert_graph_node,
    "invert_graph_node_import": invert_graph_node_import,
    "graph_node_rgba": graph_node_rgba,
    "graph_node_gray": graph_node_gray,
    "height_to_normal_node": height_to_normal_node,
}
for _mt, _nm in MIX_NAMES:
    GOLDEN_CASES["%s_node_gray" % _nm] = (lambda mt=_mt, nm=_nm: mix_node_gray(mt, "%s_node_gray.png" % nm))
    GOLDEN_CASES["%s_node_rgba" % _nm] = (lambda mt=_mt, nm=_nm: mix_node_rgba(mt, "%s_node_rgba.png" % nm))
assert len(GOLDEN_CASES) == 22


def run_oracle(case):
    """Evaluate a Case on the CPU oracle; returns the oracle Graph."""
    import oracle
    og = oracle.from_node_graph(case.graph, images={nid: decode(p) for nid, p in case.image_nodes().items()})
    for eid, path in case.embeds.items():
        og.embed(eid, oracle.deconstruct_u8(decode(path)))
    og.eval()
    return og


def run_product(tex_pro, case, request=True):
    """Evaluate a Case on the GPU backend; returns the LiveGraph."""
    lg = tex_pro.new_live_graph()
    lg.set_node_graph(case.graph)
    for eid, path in case.embeds.items():
        img = kc.SlotImage.from_u8(tex_pro, decode(path))
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, img), eid)
    if request:
        kc.LiveGraph.await_clean_read(lg, case.node)
    return lg


# ---------------------------------------------------------------------------
# BASELINE.json configs[4]: the 32-node batch graph
# ---------------------------------------------------------------------------
def config5_graph(size, lsize=None):
    """32 nodes: 3 Embed (A: Rgba size^2, H: Gray size^2, L: Rgba lsize^2), 3 Value,
    2 SeparateRgba, 12 Mix-gray in three chains of four (all five ops), 1 HeightToNormal,
    1 SeparateRgba of the normal map, 1 nested Graph (data/invert_graph.json), 2 Mix with
    SpecificSize(size^2) that upsample planes of L (Lanczos3, Gaussian), 2 CombineRgba,
    4 Mix-rgba, 1 OutputRgba.  Returns (NodeGraph, output node id)."""
    from kanter_core_b200 import ResizeFilter
    lsize = lsize or max(1, size // 4)
    g = NodeGraph.new()
    add = lambda nt: g.add_node(Node.new(nt))
    mix = lambda op: add(NodeType.Mix(op))

    def con(src, dst, out_slot, in_slot):
        g.connect(src, dst, SlotId(out_slot), SlotId(in_slot))

    eA, eH, eL = add(NodeType.Embed(0)), add(NodeType.Embed(1)), add(NodeType.Embed(2))
    v1, v2, v3 = add(NodeType.Value(0.5)), add(NodeType.Value(2.0)), add(NodeType.Value(0.25))
    sepA, sepL = add(NodeType.SeparateRgba), add(NodeType.SeparateRgba)
    con(eA, sepA, 0, 0)
    con(eL, sepL, 0, 0)
    # chain 1: m4 = pow(Ar*Ag + 0.5, 2) - Ab
    m1, m2, m3, m4 = mix(MixType.Multiply), mix(MixType.Add), mix(MixType.Pow), mix(MixType.Subtract)
    con(sepA, m1, 0, 0); con(sepA, m1, 1, 1)
    con(m1, m2, 0, 0); con(v1, m2, 0, 1)
    con(m2, m3, 0, 0); con(v2, m3, 0, 1)
    con(m3, m4, 0, 0); con(sepA, m4, 2, 1)
    # chain 2: n4 = ((Ab + H) * 0.25 - Ar) / 2
    n1, n2, n3, n4 = mix(MixType.Add), mix(MixType.Multiply), mix(MixType.Subtract), mix(MixType.Divide)
    con(sepA, n1, 2, 0); con(eH, n1, 0, 1)
    con(n1, n2, 0, 0); con(v3, n2, 0, 1)
    con(n2, n3, 0, 0); con(sepA, n3, 0, 1)
    con(n3, n4, 0, 0); con(v2, n4, 0, 1)
    # chain 3 (the height field): h4 = pow(H*0.5 + m3, 0.25) * Ag
    h1, h2, h3, h4 = mix(MixType.Multiply), mix(MixType.Add), mix(MixType.Pow), mix(MixType.Multiply)
    con(eH, h1, 0, 0); con(v1, h1, 0, 1)
    con(h1, h2, 0, 0); con(m3, h2, 0, 1)
    con(h2, h3, 0, 0); con(v3, h3, 0, 1)
    con(h3, h4, 0, 0); con(sepA, h4, 1, 1)
    h2n = add(NodeType.HeightToNormal)
    con(h4, h2n, 0, 0)
    sepN = add(NodeType.SeparateRgba)
    con(h2n, sepN, 0, 0)
    inner = NodeGraph.from_path(INVERT_JSON)
    gn = add(NodeType.Graph(inner))
    g.connect(m4, gn, SlotId(0), inner.input_slot_id_with_name("in"))
    # implicit upsample of L's planes: u1 = up_lanczos3(Lr) * Nr ; u2 = up_gaussian(Lg) + Ng
    nu1 = Node.new(NodeType.Mix(MixType.Multiply))
    nu1.resize_policy = ResizePolicy.SpecificSize(Size.new(size, size))
    nu1.resize_filter = ResizeFilter.Lanczos3
    u1 = g.add_node(nu1)
    nu2 = Node.new(NodeType.Mix(MixType.Add))
    nu2.resize_policy = ResizePolicy.SpecificSize(Size.new(size, size))
    nu2.resize_filter = ResizeFilter.Gaussian
    u2 = g.add_node(nu2)
    con(sepL, u1, 0, 0); con(sepN, u1, 0, 1)
    con(sepL, u2, 1, 0); con(sepN, u2, 1, 1)
    c1, c2 = add(NodeType.CombineRgba), add(NodeType.CombineRgba)
    con(m4, c1, 0, 0); con(n4, c1, 0, 1)
    g.connect(gn, c1, inner.output_slot_id_with_name("out"), SlotId(2))
    con(u1, c2, 0, 0); con(u2, c2, 0, 1); con(sepN, c2, 2, 2)
    r1, r2, r3, r4 = mix(MixType.Multiply), mix(MixType.Add), mix(MixType.Subtract), mix(MixType.Multiply)
    con(eA, r1, 0, 0); con(c1, r1, 0, 1)
    con(r1, r2, 0, 0); con(c2, r2, 0, 1)
    con(r2, r3, 0, 0); con(c1, r3, 0, 1)
    con(r3, r4, 0, 0); con(eA, r4, 0, 1)
    out = add(NodeType.OutputRgba("out"))
    con(r4, out, 0, 0)
    assert len(g.nodes) == 32, len(g.nodes)
    return g, out


def smooth_plane(seed, h, w):
    """A smooth height map in [0.05, 0.95]: a sum of five low-frequency sines (SURVEY.md section 8d, config 3's
    "sum-of-sines variant").  Neighbouring pixels differ by ~1e-3 / (w / 256) or less, the case in which HeightToNormal
    amplifies any error of its input by ~w / 2."""
    r = np.random.default_rng(seed)
    y, x = np.meshgrid(np.arange(h, dtype=np.float64) / h, np.arange(w, dtype=np.float64) / w, indexing="ij")
    acc = np.zeros((h, w), np.float64)
    for _ in range(5):
        fx, fy = r.integers(1, 5, 2)
        acc += r.uniform(0.3, 1.0) * np.sin(2 * np.pi * (fx * x + fy * y) + r.uniform(0, 2 * np.pi))
    acc = (acc - acc.min()) / (acc.max() - acc.min())
    return (0.05 + 0.9 * acc).astype(np.float32)


def config5_inputs(seed, size, lsize=None, smooth=False):
    """Synthetic inputs of graph `seed`: A (4 planes), H (1 plane), L (4 planes), uniform [0,1);
    smooth=True: A and H are smooth sum-of-sines maps (so is the height field the graph derives from them)."""
    lsize = lsize or max(1, size // 4)
    r = np.random.default_rng(seed)
    if smooth:
        A = [smooth_plane(seed * 10 + c, size, size) for c in range(4)]
        H = [smooth_plane(seed * 10 + 4, size, size)]
        L = [r.random((lsize, lsize), dtype=np.float32) for _ in range(4)]
        return A, H, L
    A = [r.random((size, size), dtype=np.float32) for _ in range(4)]
    H = [r.random((size, size), dtype=np.float32)]
    L = [r.random((lsize, lsize), dtype=np.float32) for _ in range(4)]
    return A, H, L


def config5_oracle(graph, out, inputs):
    import oracle
    og = oracle.from_node_graph(graph)
    for eid, planes in enumerate(inputs):
        og.embed(eid, planes)
    og.eval()
    return og.slot(int(out), 0)


def config5_product(tex_pro, graph, out, inputs, read=True):
    lg = tex_pro.new_live_graph()
    lg.set_node_graph(graph)
    for eid, planes in enumerate(inputs):
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tex_pro, planes)), eid)
    if read:
        kc.LiveGraph.await_clean_read(lg, out)
    return lg
