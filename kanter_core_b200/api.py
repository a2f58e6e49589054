"""Host-side mirror of the vismut_core (kanter_core) public interface, on top of
the C ABI.  Same names, argument meaning and error behaviour as the crate, so a
test written against the reference reads the same here:

    tex_pro = TextureProcessor.new()
    live_graph = tex_pro.new_live_graph()
    mix = live_graph.add_node(Node.new(NodeType.Mix(MixType.Add)))
    live_graph.connect(a, mix, SlotId(0), SlotId(0))
    rgba8 = LiveGraph.await_clean_read(live_graph, out).buffer_rgba(out, SlotId(0))

Reference: src/lib.rs:1-13 and the modules it exports.  All pixel work happens
in libkanter_b200.so on the GPU; this file only moves handles around.
"""
import ctypes as C
import enum
import json as _json

import numpy as np

from . import _lib
from ._lib import TexProError, call, kc_edge, kc_image, kc_node_desc, kc_options, kc_slot


# ---- ids, sizes, enums (src/node_graph.rs:592-624, src/slot_data.rs:5-30) -------
class NodeId(int):
    def __repr__(self):
        return "NodeId(%d)" % int(self)

    def as_usize(self):
        return int(self)


class SlotId(int):
    def __repr__(self):
        return "SlotId(%d)" % int(self)

    def as_usize(self):
        return int(self)


class EmbeddedSlotDataId(int):  # src/node/embed.rs:13-14
    pass


class Size:
    def __init__(self, width, height):
        self.width, self.height = int(width), int(height)

    @staticmethod
    def new(width, height):
        return Size(width, height)

    def pixel_count(self):
        return (self.width * self.height) & 0xFFFFFFFF  # u32 multiply, src/slot_data.rs:27-29

    def __eq__(self, o):
        return isinstance(o, Size) and (self.width, self.height) == (o.width, o.height)

    def __iter__(self):
        return iter((self.width, self.height))

    def __repr__(self):
        return "%dx%d" % (self.width, self.height)


class MixType(enum.IntEnum):  # src/node/mix.rs:20-26
    Add = 0
    Subtract = 1
    Multiply = 2
    Divide = 3
    Pow = 4

    @staticmethod
    def default():
        return MixType.Add


class ResizeFilter(enum.IntEnum):  # src/node/mod.rs:63-69
    Nearest = 0
    Triangle = 1
    CatmullRom = 2
    Gaussian = 3
    Lanczos3 = 4

    @staticmethod
    def default():
        return ResizeFilter.Triangle


class ResizePolicy:  # src/node/mod.rs:33-40
    _NAMES = ["MostPixels", "LeastPixels", "LargestAxes", "SmallestAxes", "SpecificSlot", "SpecificSize"]

    def __init__(self, kind, slot=0, size=None):
        self.kind, self.slot, self.size = kind, SlotId(slot), size or Size(0, 0)

    @staticmethod
    def SpecificSlot(slot_id):
        return ResizePolicy(4, slot=slot_id)

    @staticmethod
    def SpecificSize(size):
        return ResizePolicy(5, size=size)

    @staticmethod
    def default():
        return ResizePolicy.MostPixels

    def __eq__(self, o):
        return isinstance(o, ResizePolicy) and (self.kind, int(self.slot), tuple(self.size)) == (o.kind, int(o.slot), tuple(o.size))

    def __repr__(self):
        return self._NAMES[self.kind]


ResizePolicy.MostPixels = ResizePolicy(0)
ResizePolicy.LeastPixels = ResizePolicy(1)
ResizePolicy.LargestAxes = ResizePolicy(2)
ResizePolicy.SmallestAxes = ResizePolicy(3)


class Side(enum.IntEnum):  # src/node/mod.rs:101-105
    Input = 0
    Output = 1


class SlotType(enum.IntEnum):  # src/node/mod.rs:197-202
    Gray = 0
    Rgba = 1
    GrayOrRgba = 2


class NodeState(enum.IntEnum):  # src/live_graph.rs:23-37
    Clean = 0
    Dirty = 1
    Requested = 2
    Prioritised = 3
    Processing = 4
    ProcessingDirty = 5


class NodeType:
    """`enum NodeType`, src/node/node_type.rs:14-28."""
    _NAMES = ["InputGray", "InputRgba", "OutputGray", "OutputRgba", "Graph", "Image", "Embed", "Write",
              "Value", "Mix", "HeightToNormal", "SeparateRgba", "CombineRgba"]

    def __init__(self, kind, payload=None):
        self.kind, self.payload = kind, payload

    InputGray = staticmethod(lambda name: NodeType(_lib.NODE_INPUT_GRAY, str(name)))
    InputRgba = staticmethod(lambda name: NodeType(_lib.NODE_INPUT_RGBA, str(name)))
    OutputGray = staticmethod(lambda name: NodeType(_lib.NODE_OUTPUT_GRAY, str(name)))
    OutputRgba = staticmethod(lambda name: NodeType(_lib.NODE_OUTPUT_RGBA, str(name)))
    Graph = staticmethod(lambda graph: NodeType(_lib.NODE_GRAPH, graph))
    Image = staticmethod(lambda path: NodeType(_lib.NODE_IMAGE, str(path)))
    Embed = staticmethod(lambda esd_id: NodeType(_lib.NODE_EMBED, EmbeddedSlotDataId(esd_id)))
    Write = staticmethod(lambda path: NodeType(_lib.NODE_WRITE, str(path)))
    Value = staticmethod(lambda v: NodeType(_lib.NODE_VALUE, float(np.float32(v))))
    Mix = staticmethod(lambda mix_type: NodeType(_lib.NODE_MIX, MixType(mix_type)))

    def is_input(self):
        return self.kind in (_lib.NODE_INPUT_GRAY, _lib.NODE_INPUT_RGBA)

    def is_output(self):
        return self.kind in (_lib.NODE_OUTPUT_GRAY, _lib.NODE_OUTPUT_RGBA)

    def name(self):
        return self.payload if self.kind <= _lib.NODE_OUTPUT_RGBA else None

    def __eq__(self, o):  # discriminant comparison, node_type.rs:50-54
        return isinstance(o, NodeType) and self.kind == o.kind

    def __repr__(self):
        return self._NAMES[self.kind] if self.payload is None else "%s(%r)" % (self._NAMES[self.kind], self.payload)


NodeType.HeightToNormal = NodeType(_lib.NODE_HEIGHT_TO_NORMAL)
NodeType.SeparateRgba = NodeType(_lib.NODE_SEPARATE_RGBA)
NodeType.CombineRgba = NodeType(_lib.NODE_COMBINE_RGBA)


class Slot:  # src/node/mod.rs:223-238
    def __init__(self, name, slot_id, slot_type):
        self.name, self.slot_id, self.slot_type = name, SlotId(slot_id), SlotType(slot_type)

    def __repr__(self):
        return "Slot(%r, %d, %s)" % (self.name, self.slot_id, self.slot_type.name)


class Priority:
    """`struct Priority`, src/priority.rs:11-45.  On a node that lives in a graph the value is the graph's (the
    reference shares an Arc<Priority> between the node and the engine); on a free-standing node it travels with
    the node into add_node."""

    def __init__(self, value=0, owner=None, node_id=None):
        self._value, self._owner, self._node_id = int(value), owner, node_id

    def set_priority(self, val):  # :33-37
        val = int(val)
        if not -128 <= val <= 127:
            raise ValueError("priority is an i8")
        self._value = val
        if self._owner is not None:
            self._owner._set_priority(self._node_id, val)

    def priority(self):  # :43-45
        if self._owner is not None:
            return self._owner._priority(self._node_id)[0]
        return self._value

    def propagated_priority(self):  # :39-41, after PriorityPropagator::update (:101-127)
        if self._owner is not None:
            return self._owner._priority(self._node_id)[1]
        return self._value


class Node:
    """`struct Node`, src/node/mod.rs:114-123."""

    def __init__(self, node_type, node_id=0):
        self.node_id = NodeId(node_id)
        self.node_type = node_type
        self.resize_policy = ResizePolicy.default()
        self.resize_filter = ResizeFilter.default()
        self.priority = Priority()

    new = staticmethod(lambda node_type: Node(node_type))
    with_id = staticmethod(lambda node_type, node_id: Node(node_type, node_id))

    def _desc(self):
        """(kc_node_desc, keepalive) for the C ABI."""
        d = kc_node_desc()
        t = self.node_type
        keep = []
        d.node_id = int(self.node_id)
        d.node_type = t.kind
        if t.kind == _lib.NODE_VALUE:
            d.value = t.payload
        elif t.kind == _lib.NODE_MIX:
            d.mix_type = int(t.payload)
        elif t.kind == _lib.NODE_EMBED:
            d.embed_id = int(t.payload)
        elif t.kind == _lib.NODE_GRAPH:
            d.graph = t.payload._h
            keep.append(t.payload)
        elif t.payload is not None:
            b = t.payload.encode("utf-8")
            keep.append(b)
            d.name = b
        d.resize_policy = self.resize_policy.kind
        d.policy_slot = int(self.resize_policy.slot)
        d.policy_width, d.policy_height = self.resize_policy.size.width, self.resize_policy.size.height
        d.resize_filter = int(self.resize_filter)
        return d, keep

    @staticmethod
    def _from_desc(d):
        k = d.node_type
        if k == _lib.NODE_VALUE:
            t = NodeType(k, float(d.value))
        elif k == _lib.NODE_MIX:
            t = NodeType(k, MixType(d.mix_type))
        elif k == _lib.NODE_EMBED:
            t = NodeType(k, EmbeddedSlotDataId(d.embed_id))
        elif k == _lib.NODE_GRAPH:
            t = NodeType(k, NodeGraph._clone_of(d.graph))
        elif k in (_lib.NODE_HEIGHT_TO_NORMAL, _lib.NODE_SEPARATE_RGBA, _lib.NODE_COMBINE_RGBA):
            t = NodeType(k)
        else:
            t = NodeType(k, (d.name or b"").decode("utf-8"))
        n = Node(t, d.node_id)
        if d.resize_policy == 4:
            n.resize_policy = ResizePolicy.SpecificSlot(SlotId(d.policy_slot))
        elif d.resize_policy == 5:
            n.resize_policy = ResizePolicy.SpecificSize(Size(d.policy_width, d.policy_height))
        else:
            n.resize_policy = ResizePolicy(d.resize_policy)
        n.resize_filter = ResizeFilter(d.resize_filter)
        return n

    def _slots(self, fn):
        d, keep = self._desc()
        arr = (kc_slot * 64)()
        n = C.c_size_t()
        call(fn, C.byref(d), arr, 64, C.byref(n))
        return [Slot(arr[i].name.decode(), arr[i].slot_id, arr[i].slot_type) for i in range(n.value)]

    def input_slots(self):  # node_type.rs:141-175
        return self._slots("kc_node_input_slots")

    def output_slots(self):  # node_type.rs:177-211
        return self._slots("kc_node_output_slots")

    def input_slot_with_name(self, name):
        for s in self.input_slots():
            if s.name == name:
                return s
        raise TexProError(19, "no input slot named %r" % name)

    def output_slot_with_name(self, name):
        for s in self.output_slots():
            if s.name == name:
                return s
        raise TexProError(19, "no output slot named %r" % name)

    def __repr__(self):
        return "Node(%d, %r)" % (self.node_id, self.node_type)


class Edge:  # src/edge.rs:9-14
    def __init__(self, output_id, input_id, output_slot, input_slot):
        self.output_id, self.input_id = NodeId(output_id), NodeId(input_id)
        self.output_slot, self.input_slot = SlotId(output_slot), SlotId(input_slot)

    new = staticmethod(lambda a, b, c, d: Edge(a, b, c, d))

    def _tuple(self):
        return (int(self.output_id), int(self.input_id), int(self.output_slot), int(self.input_slot))

    def __eq__(self, o):
        return isinstance(o, Edge) and self._tuple() == o._tuple()

    def __repr__(self):
        return "Edge(%d:%d -> %d:%d)" % (self.output_id, self.output_slot, self.input_id, self.input_slot)


class _GraphView:
    """Read-side of a kc_graph handle; shared by NodeGraph and LiveGraph."""

    def _graph_handle(self):
        raise NotImplementedError

    @property
    def nodes(self):
        h = self._graph_handle()
        n = C.c_size_t()
        call("kc_graph_node_count", h, C.byref(n))
        out = []
        for i in range(n.value):
            d = kc_node_desc()
            call("kc_graph_node_at", h, i, C.byref(d))
            out.append(Node._from_desc(d))
        return out

    @property
    def edges(self):
        h = self._graph_handle()
        n = C.c_size_t()
        call("kc_graph_edge_count", h, C.byref(n))
        out = []
        for i in range(n.value):
            e = kc_edge()
            call("kc_graph_edge_at", h, i, C.byref(e))
            out.append(Edge(e.output_id, e.input_id, e.output_slot, e.input_slot))
        return out

    def node(self, node_id):
        d = kc_node_desc()
        call("kc_graph_node", self._graph_handle(), int(node_id), C.byref(d))
        n = Node._from_desc(d)
        n.priority = Priority(0, self, int(node_id))   # node(id)?.priority.set_priority(v) reaches the graph's node
        return n

    def _priority(self, node_id):
        own, prop = C.c_int8(), C.c_int8()
        call("kc_graph_node_priority", self._graph_handle(), int(node_id), C.byref(own), C.byref(prop))
        return own.value, prop.value

    def _set_priority(self, node_id, val):
        call("kc_graph_set_node_priority", self._graph_handle(), int(node_id), int(val))

    def has_node_with_id(self, node_id):
        self.node(node_id)

    def node_ids(self):
        return [n.node_id for n in self.nodes]

    def _ids(self, fn):
        arr = (C.c_uint32 * 4096)()
        n = C.c_size_t()
        call(fn, self._graph_handle(), arr, 4096, C.byref(n))
        return [NodeId(arr[i]) for i in range(n.value)]

    def output_ids(self):
        return self._ids("kc_graph_output_ids")

    def input_ids(self):
        return self._ids("kc_graph_input_ids")

    def input_slot_id_with_name(self, name):
        s = C.c_uint32()
        call("kc_graph_input_slot_id_with_name", self._graph_handle(), name.encode(), C.byref(s))
        return SlotId(s.value)

    def output_slot_id_with_name(self, name):
        s = C.c_uint32()
        call("kc_graph_output_slot_id_with_name", self._graph_handle(), name.encode(), C.byref(s))
        return SlotId(s.value)

    def input_edges(self, node_id):
        return [e for e in self.edges if e.input_id == node_id]

    def get_parents(self, node_id):
        return sorted({e.output_id for e in self.edges if e.input_id == node_id})

    def get_children(self, node_id):
        self.has_node_with_id(node_id)
        return sorted({e.input_id for e in self.edges if e.output_id == node_id})

    def can_connect(self, output_node, input_node, output_slot, input_slot):  # node_graph.rs:376-393
        call("kc_graph_can_connect", self._graph_handle(), int(output_node), int(input_node), int(output_slot), int(input_slot))

    def connected_edges(self, node_id, side, slot_id):  # node_graph.rs:518-537
        arr = (kc_edge * 256)()
        n = C.c_size_t()
        call("kc_graph_connected_edges", self._graph_handle(), int(node_id), int(side), int(slot_id), arr, 256, C.byref(n))
        return [Edge(e.output_id, e.input_id, e.output_slot, e.input_slot) for e in arr[:n.value]]

    def get_children_recursive(self, node_id):  # node_graph.rs:566-575 (duplicates kept, like the reference)
        kids = self.get_children(node_id)
        out = list(kids)
        for c in kids:
            out += self.get_children_recursive(c)
        return out

    def output_names(self):
        return [n.node_type.payload for n in self.nodes if n.node_type.is_output()]

    def input_names(self):
        return [n.node_type.payload for n in self.nodes if n.node_type.is_input()]

    def input_nodes(self):  # node_graph.rs:191-196
        return [n for n in self.nodes if n.node_type.is_input()]

    def output_nodes(self):  # :198-203
        return [n for n in self.nodes if n.node_type.is_output()]

    def edge_indices_node(self, node_id):  # :351-361
        self.has_node_with_id(node_id)
        return [i for i, e in enumerate(self.edges) if e.output_id == node_id or e.input_id == node_id]

    def edge_indices_slot(self, node_id, side, slot_id):  # :364-374
        if int(side) == int(Side.Input):
            return [i for i, e in enumerate(self.edges) if e.input_id == node_id and e.input_slot == slot_id]
        return [i for i, e in enumerate(self.edges) if e.output_id == node_id and e.output_slot == slot_id]

    def slot_occupied(self, node_id, side, slot_id):  # :449-460
        return bool(self.edge_indices_slot(node_id, side, slot_id))

    def export_json_string(self):
        p = C.c_void_p()
        call("kc_graph_export_json", self._graph_handle(), C.byref(p))
        try:
            return C.string_at(p).decode("utf-8")
        finally:
            _lib.lib.kc_free(p)


class NodeGraph(_GraphView):
    """`struct NodeGraph`, src/node_graph.rs:17-22."""

    def __init__(self, handle=None):
        if handle is None:
            h = C.c_void_p()
            call("kc_graph_create", C.byref(h))
            handle = h
        self._h = handle

    new = staticmethod(lambda: NodeGraph())

    def __del__(self):
        if getattr(self, "_h", None):
            _lib.lib.kc_graph_destroy(self._h)
            self._h = None

    def _graph_handle(self):
        return self._h

    @staticmethod
    def _clone_of(raw_handle):
        h = C.c_void_p()
        call("kc_graph_clone", raw_handle, C.byref(h))
        return NodeGraph(h)

    def clone(self):
        return NodeGraph._clone_of(self._h)

    @staticmethod
    def from_path(path):  # :33-46
        h = C.c_void_p()
        call("kc_graph_from_path", str(path).encode(), C.byref(h))
        return NodeGraph(h)

    @staticmethod
    def from_json(text):
        h = C.c_void_p()
        call("kc_graph_from_json", text.encode("utf-8"), C.byref(h))
        return NodeGraph(h)

    def export_json(self, path):  # :98-102
        call("kc_graph_export_json_path", self._h, str(path).encode())

    def add_node(self, node):  # :332-337
        d, keep = node._desc()
        out = C.c_uint32()
        call("kc_graph_add_node", self._h, C.byref(d), C.byref(out))
        if node.priority.priority():
            self._set_priority(out.value, node.priority.priority())
        return NodeId(out.value)

    def add_node_with_id(self, node):  # :339-348
        d, keep = node._desc()
        call("kc_graph_add_node_with_id", self._h, C.byref(d))
        if node.priority.priority():
            self._set_priority(int(node.node_id), node.priority.priority())

    def remove_node(self, node_id):
        call("kc_graph_remove_node", self._h, int(node_id))

    def connect(self, output_node, input_node, output_slot, input_slot):  # :416-446
        call("kc_graph_connect", self._h, int(output_node), int(input_node), int(output_slot), int(input_slot))
        return Edge(output_node, input_node, output_slot, input_slot)

    def try_connect(self, output_node, input_node, output_slot, input_slot):  # :394-413
        call("kc_graph_try_connect", self._h, int(output_node), int(input_node), int(output_slot), int(input_slot))

    def disconnect_slot(self, node_id, side, slot_id):  # :500-520
        call("kc_graph_disconnect_slot", self._h, int(node_id), int(side), int(slot_id))

    def remove_edge(self, edge):
        e = kc_edge(*edge._tuple())
        call("kc_graph_remove_edge", self._h, C.byref(e))

    def new_id(self):  # :86-96
        out = C.c_uint32()
        call("kc_graph_new_id", self._h, C.byref(out))
        return NodeId(out.value)

    def rename_output_node(self, node_id, new_name):  # :232-270, returns the old name
        p = C.c_void_p()
        call("kc_graph_rename_output_node", self._h, int(node_id), new_name.encode(), C.byref(p))
        try:
            return C.string_at(p).decode("utf-8")
        finally:
            _lib.lib.kc_free(p)

    def set_image_node_path(self, node_id, path):  # :65-83: only an Image node takes a path
        n = self.node(node_id)
        if n.node_type.kind != _lib.NODE_IMAGE:
            raise TexProError(5, "node %d is not an Image node" % node_id)
        n.node_type = NodeType.Image(str(path))
        d, keep = n._desc()
        call("kc_graph_set_node", self._h, C.byref(d))

    def set_mix_type(self, node_id, mix_type):  # :48-63
        n = self.node(node_id)
        if n.node_type.kind != _lib.NODE_MIX:
            raise TexProError(5, "node %d is not a Mix node" % node_id)
        n.node_type = NodeType.Mix(mix_type)
        d, keep = n._desc()
        call("kc_graph_set_node", self._h, C.byref(d))


# ---- pixel data -------------------------------------------------------------------
class SlotImage:
    """`enum SlotImage { Gray(plane), Rgba([plane; 4]) }`, src/slot_image.rs:16-19.
    Owns one reference on each device plane."""

    def __init__(self, ctx, raw):
        self._ctx = ctx          # _Context
        self._im = raw           # kc_image (owned references)

    def __del__(self):
        im = getattr(self, "_im", None)
        if im is not None and _lib is not None and _lib.lib is not None:
            _lib.lib.kc_image_release(C.byref(im))   # legal after the context was closed: planes keep its bookkeeping alive
            self._im = None

    # constructors
    @staticmethod
    def from_value(tex_pro, size, value, rgba):  # :28-64
        im = kc_image()
        call("kc_image_from_value", tex_pro._ctx._h, size.width, size.height, float(value), int(bool(rgba)), C.byref(im))
        return SlotImage(tex_pro._ctx, im)

    @staticmethod
    def from_planes(tex_pro, planes, sync=True, deferred=False):
        """planes: one (Gray) or four (Rgba) float32 arrays of shape (h, w).  With
        sync=False the copies are only enqueued: the arrays (pinned, see
        pinned_empty) must stay alive and unchanged until the stream has passed them.
        deferred=True: nothing is copied now; each plane goes up (on the upload stream) when
        something first reads it, and a plane nobody reads never crosses PCIe."""
        arrs = [np.ascontiguousarray(p, dtype=np.float32) for p in planes]
        if len(arrs) not in (1, 4):
            raise TexProError(4, "need 1 or 4 planes")
        h, w = arrs[0].shape
        ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
        im = kc_image()
        if deferred:
            call("kc_image_from_host_planes_deferred", tex_pro._ctx._h, 1 if len(arrs) == 4 else 0, w, h, ptrs, C.byref(im))
            img = SlotImage(tex_pro._ctx, im)
            img._keep = arrs                                 # the planes read from these arrays later
            return img
        call("kc_image_from_host_planes", tex_pro._ctx._h, 1 if len(arrs) == 4 else 0, w, h, ptrs, C.byref(im))
        if sync:
            call("kc_context_synchronize", tex_pro._ctx._h)  # the host arrays may go away after this returns
        return SlotImage(tex_pro._ctx, im)

    @staticmethod
    def from_u8(tex_pro, samples, sync=True):
        """Decoded interleaved u8 samples, shape (h, w) or (h, w, c): deconstruct_image, src/shared.rs:16-56.
        sync=False: returns once the copy and the conversion are enqueued (the samples must then be in
        pinned memory and stay untouched until the context has been synchronised)."""
        a = np.ascontiguousarray(samples, dtype=np.uint8)
        if a.ndim == 2:
            a = a[:, :, None]
        h, w, c = a.shape
        im = kc_image()
        call("kc_image_from_u8", tex_pro._ctx._h, a.ctypes.data, w, h, c, C.byref(im))
        if sync:
            call("kc_context_synchronize", tex_pro._ctx._h)
        return SlotImage(tex_pro._ctx, im)

    @staticmethod
    def from_buffers_rgba(tex_pro, buffers):  # :66-88: exactly four planes
        if len(buffers) != 4:
            raise TexProError(4, "from_buffers_rgba needs 4 buffers, got %d" % len(buffers))
        return SlotImage.from_planes(tex_pro, list(buffers))

    @staticmethod
    def from_buffers_rgb(tex_pro, buffers):  # :90-102: three planes and an alpha of 1.0
        if len(buffers) != 3:
            raise TexProError(4, "from_buffers_rgb needs 3 buffers, got %d" % len(buffers))
        b = [np.ascontiguousarray(x, dtype=np.float32) for x in buffers]
        return SlotImage.from_planes(tex_pro, b + [np.ones_like(b[0])])

    def from_self(self):  # :104-114: the planes are immutable, so a copy is another reference
        im = kc_image()
        C.memmove(C.byref(im), C.byref(self._im), C.sizeof(kc_image))
        call("kc_image_retain", C.byref(im))
        return SlotImage(self._ctx, im)

    def in_memory(self):  # SlotData::in_memory, src/slot_data.rs:70-78: no plane sits in the spill queue's host memory
        for c in range(4 if self.is_rgba() else 1):
            v = C.c_int32()
            call("kc_plane_in_memory", self._im.planes[c], C.byref(v))
            if not v.value:
                return False
        return True

    def is_rgba(self):
        return self._im.kind == _lib.IMAGE_RGBA

    def size(self):  # :116-121
        w, h = C.c_uint32(), C.c_uint32()
        call("kc_plane_size", self._im.planes[0], C.byref(w), C.byref(h))
        return Size(w.value, h.value)

    def _n(self):
        return 4 if self.is_rgba() else 1

    def to_u8(self, srgb=False):  # :142-207
        s = self.size()
        out = np.empty((s.height, s.width, 4), dtype=np.uint8)
        call("kc_image_to_u8", self._ctx._h, C.byref(self._im), int(bool(srgb)), out.ctypes.data)
        return out

    def to_u8_srgb(self):
        return self.to_u8(True)

    def as_type(self, rgba):  # :212-256
        im = kc_image()
        call("kc_image_as_type", self._ctx._h, C.byref(self._im), int(bool(rgba)), C.byref(im))
        return SlotImage(self._ctx, im)

    def planes(self):
        """The f32 planes on the host, one (h, w) array per plane."""
        out = []
        for c in range(self._n()):
            w, h = C.c_uint32(), C.c_uint32()
            call("kc_plane_size", self._im.planes[c], C.byref(w), C.byref(h))
            a = np.empty((h.value, w.value), dtype=np.float32)
            call("kc_plane_download", self._im.planes[c], a.ctypes.data)
            out.append(a)
        return out

    bufs = planes

    def plane_is_constant(self, c):
        k, v = C.c_int32(), C.c_float()
        call("kc_plane_is_constant", self._im.planes[c], C.byref(k), C.byref(v))
        return bool(k.value), v.value

    def plane_handles(self):
        return [self._im.planes[c] for c in range(self._n())]

    def same_plane(self, c, other, oc):
        return self._im.planes[c] == other._im.planes[oc]


class SlotData:  # src/slot_data.rs:35-39
    def __init__(self, node_id, slot_id, image):
        self.node_id, self.slot_id, self.image = NodeId(node_id), SlotId(slot_id), image

    new = staticmethod(lambda n, s, i: SlotData(n, s, i))

    def size(self):
        return self.image.size()

    def from_self(self):  # :62-64: a new SlotData over the same (immutable, shared) planes
        return SlotData(self.node_id, self.slot_id, self.image.from_self())

    def in_memory(self):  # :70-78
        return self.image.in_memory()


class _Context:
    def __init__(self, device, math_mode, fuse, cuda_stream=None):
        o = kc_options()
        _lib.lib.kc_options_default(C.byref(o))
        o.math_mode = int(math_mode)
        o.fuse = int(bool(fuse))
        self._h = C.c_void_p()
        if cuda_stream is None:
            call("kc_context_create", int(device), C.byref(o), C.byref(self._h))
        else:   # the caller's stream, e.g. torch.cuda.Stream().cuda_stream
            call("kc_context_create_on_stream", int(device), C.byref(o), C.c_void_p(int(cuda_stream)), C.byref(self._h))

    def close(self):
        if self._h:
            _lib.lib.kc_context_destroy(self._h)
            self._h = None


def jit_wait(timeout_ms=60000):
    """Hot tapes are compiled in the background (kc_jit.cu); wait until no compile is running.
    Returns the number still running.  Benchmarks call it at the end of their warm-up, then
    run one more untimed step so the finished kernel is loaded before anything is timed."""
    n = C.c_int32()
    call("kc_debug_jit_wait", int(timeout_ms), C.byref(n))
    return n.value


def _is_png(path):
    try:
        with open(path, "rb") as f:
            return f.read(8) == b"\x89PNG\r\n\x1a\n"
    except OSError:
        return False


def _decode_image(path):
    """Host-side codec for Image nodes (image::open + as_flat_samples_u8,
    src/shared.rs:16-56,241): 8-bit samples, 1..4 channels."""
    from PIL import Image as PILImage
    im = PILImage.open(path)
    if im.mode in ("1", "I", "F", "I;16"):
        im = im.convert("L")
    elif im.mode == "P":
        im = im.convert("RGBA" if "transparency" in im.info else "RGB")
    elif im.mode not in ("L", "LA", "RGB", "RGBA"):
        im = im.convert("RGBA")
    return np.asarray(im, dtype=np.uint8)


class TextureProcessor:
    """`TextureProcessor`, src/texture_processor.rs:17-115.  One per device: owns
    the CUDA context/stream/plane pool that replace the engine and
    transient-buffer threads."""

    def __init__(self, memory_threshold=None, device=0, math_mode=_lib.MATH_EXACT, fuse=True, cuda_stream=None):
        self.memory_threshold = memory_threshold
        self._ctx = _Context(device, math_mode, fuse, cuda_stream)
        self._live_graphs = []

    @staticmethod
    def new(memory_threshold=None, **kw):
        return TextureProcessor(memory_threshold, **kw)

    def new_live_graph(self):  # :58-63
        lg = LiveGraph(self)
        self._live_graphs.append(lg)
        return lg

    def push_live_graph(self, live_graph):  # :65-69
        self._live_graphs.append(live_graph)

    def live_graph(self):  # :71-73
        return self._live_graphs

    @staticmethod
    def buffer_rgba(live_graph, node_id, slot_id):  # :75-81: await_clean_write(..).buffer_rgba(..)
        return LiveGraph.await_clean_write(live_graph, node_id).buffer_rgba(node_id, slot_id)

    @staticmethod
    def await_slot_data_size(live_graph, node_id, slot_id):  # :91-105: prioritise the node, wait for its slot, return the size
        return LiveGraph.await_clean_read(live_graph, node_id).slot_data_size(node_id, slot_id)

    def processing_node_count(self):  # :107-109: nodes being processed right now; evaluation here is synchronous per request
        return 0

    def set_max_processing_nodes(self, count):  # :111-114 -> ProcessPackManager::max_count: nodes admitted per engine turn
        call("kc_context_set_max_processing_nodes", self._ctx._h, int(count))

    def max_processing_nodes(self):
        n = C.c_size_t()
        call("kc_context_max_processing_nodes", self._ctx._h, C.byref(n))
        return n.value

    def concurrent(self, lanes=2):
        """Concurrent section (`with tex_pro.concurrent(2): ...`): live graphs evaluated inside it and served by evaluation
        replay run on `lanes` side streams, so independent graphs overlap on the device -- what the reference's thread
        pool does with ready nodes (src/process_pack.rs:27, src/engine.rs:288).  The graphs of one section must not consume
        each other's results; read results after the section (include/kanter_b200.h, kc_context_concurrent_begin)."""
        tp = self

        class _Section:
            def __enter__(self_inner):
                call("kc_context_concurrent_begin", tp._ctx._h, int(lanes))
                return tp

            def __exit__(self_inner, *exc):
                call("kc_context_concurrent_end", tp._ctx._h)
                return False
        return _Section()

    def set_math_mode(self, mode):
        call("kc_context_set_math_mode", self._ctx._h, int(mode))

    def set_resize_clamp(self, clamp):
        """The [0,1] clamp of a resize's second pass (image-0.24 semantics, on by default; no reference golden pins it)."""
        call("kc_context_set_resize_unclamped", self._ctx._h, int(not clamp))

    def set_fuse(self, fuse):
        call("kc_context_set_fuse", self._ctx._h, int(bool(fuse)))

    def synchronize(self):
        call("kc_context_synchronize", self._ctx._h)

    def set_memory_threshold(self, nbytes):
        """TextureProcessor::memory_threshold (src/texture_processor.rs:19): above this many bytes of live
        planes the least recently used ones are spilled (here: to pinned host memory); 0 = no limit."""
        call("kc_context_set_memory_threshold", self._ctx._h, int(nbytes))

    def transfer_stats(self):
        a, b = C.c_uint64(), C.c_uint64()
        call("kc_context_transfer_stats", self._ctx._h, C.byref(a), C.byref(b))
        return {"h2d_bytes": a.value, "d2h_bytes": b.value}

    def spill_stats(self):
        b, s, r = C.c_uint64(), C.c_uint64(), C.c_uint64()
        call("kc_context_spill_stats", self._ctx._h, C.byref(b), C.byref(s), C.byref(r))
        return {"bytes_spilled": b.value, "spills": s.value, "reloads": r.value}

    def stats(self):
        k, b = C.c_uint64(), C.c_uint64()
        call("kc_context_stats", self._ctx._h, C.byref(k), C.byref(b))
        return {"kernel_launches": k.value, "bytes_memory": b.value}

    def bytes_memory(self):  # TransientBufferQueue::bytes_memory, src/transient_buffer.rs:413-420
        return self.stats()["bytes_memory"]

    def close(self):
        for lg in self._live_graphs:
            lg.close()
        self._live_graphs = []
        self._ctx.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class LiveGraph(_GraphView):
    """`LiveGraph`, src/live_graph.rs:63-74."""

    def __init__(self, tex_pro):
        self._tp = tex_pro
        self._ctx = tex_pro._ctx
        self._h = C.c_void_p()
        call("kc_live_graph_create", self._ctx._h, C.byref(self._h))
        self._use_cache = False
        self._auto_update = False
        self._images_loaded = {}
        self._scan_images = False   # set when the graph may have gained an Image node

    def close(self):
        if getattr(self, "_h", None) and _lib is not None and _lib.lib is not None:
            _lib.lib.kc_live_graph_destroy(self._h)   # legal after the context was closed
        self._h = None

    def __del__(self):
        self.close()

    # RwLock-style accessors so code written for Arc<RwLock<LiveGraph>> reads the same
    def write(self):
        return self

    def read(self):
        return self

    def unwrap(self):
        return self

    def _graph_handle(self):
        g = C.c_void_p()
        call("kc_live_graph_node_graph", self._h, C.byref(g))
        return g

    use_cache = property(lambda s: s._use_cache, lambda s, v: (setattr(s, "_use_cache", bool(v)), call("kc_live_graph_set_use_cache", s._h, int(bool(v))))[0])
    auto_update = property(lambda s: s._auto_update, lambda s, v: (setattr(s, "_auto_update", bool(v)), call("kc_live_graph_set_auto_update", s._h, int(bool(v))))[0])

    def set_node_graph(self, node_graph):
        call("kc_live_graph_set_node_graph", self._h, node_graph._h)
        self._images_loaded = {}
        self._scan_images = True

    def add_node(self, node):
        d, keep = node._desc()
        out = C.c_uint32()
        call("kc_live_graph_add_node", self._h, C.byref(d), C.byref(out))
        self._scan_images |= node.node_type.kind == _lib.NODE_IMAGE
        if node.priority.priority():
            self._set_priority(out.value, node.priority.priority())
        return NodeId(out.value)

    def add_node_with_id(self, node):
        d, keep = node._desc()
        call("kc_live_graph_add_node_with_id", self._h, C.byref(d))
        self._scan_images |= node.node_type.kind == _lib.NODE_IMAGE
        if node.priority.priority():
            self._set_priority(int(node.node_id), node.priority.priority())

    def _set_priority(self, node_id, val):
        call("kc_live_graph_set_priority", self._h, int(node_id), int(val))

    def remove_node(self, node_id):
        call("kc_live_graph_remove_node", self._h, int(node_id))

    def connect(self, output_node, input_node, output_slot, input_slot):
        call("kc_live_graph_connect", self._h, int(output_node), int(input_node), int(output_slot), int(input_slot))
        return Edge(output_node, input_node, output_slot, input_slot)

    def disconnect_slot(self, node_id, side, slot_id):
        call("kc_live_graph_disconnect_slot", self._h, int(node_id), int(side), int(slot_id))

    def set_node(self, node):
        d, keep = node._desc()
        call("kc_live_graph_set_node", self._h, C.byref(d))
        self._scan_images |= node.node_type.kind == _lib.NODE_IMAGE

    def set_mix_type(self, node_id, mix_type):  # node_mut(..).node_type = Mix(..), :369-374: the node becomes dirty
        n = self.node(node_id)
        if n.node_type.kind != _lib.NODE_MIX:
            raise TexProError(5, "node %d is not a Mix node" % node_id)
        n.node_type = NodeType.Mix(mix_type)
        self.set_node(n)

    def add_input_slot_data(self, slot_data):  # :347-350
        call("kc_live_graph_add_input_slot_data", self._h, int(slot_data.node_id), int(slot_data.slot_id), C.byref(slot_data.image._im))

    def embed_slot_data_with_id(self, slot_data, esd_id):  # :324-341
        call("kc_live_graph_embed_slot_data_with_id", self._h, C.byref(slot_data.image._im), int(slot_data.slot_id), int(esd_id))
        return EmbeddedSlotDataId(esd_id)

    def replace_embedded(self, image, esd_id):
        call("kc_live_graph_replace_embedded", self._h, C.byref(image._im), int(esd_id))

    def _load_images(self):
        # PNG files are decoded by the library itself (kc_png.cu) when the Image node runs; any
        # other format the `image` crate would have opened is decoded here with Pillow and handed
        # over as u8 samples
        if not self._scan_images:
            return
        self._scan_images = False
        for n in self.nodes:
            if n.node_type.kind == _lib.NODE_IMAGE and self._images_loaded.get(int(n.node_id)) != n.node_type.payload:
                if _is_png(n.node_type.payload):
                    self._images_loaded[int(n.node_id)] = n.node_type.payload
                    continue
                try:
                    px = _decode_image(n.node_type.payload)
                except Exception:
                    self._images_loaded[int(n.node_id)] = n.node_type.payload
                    continue  # unreadable => the node yields 1x1 magenta (src/node/image.rs:13-18)
                if px.ndim == 2:
                    px = px[:, :, None]
                px = np.ascontiguousarray(px)
                call("kc_live_graph_set_image_data_u8", self._h, int(n.node_id), px.ctypes.data, px.shape[1], px.shape[0], px.shape[2])
                self._images_loaded[int(n.node_id)] = n.node_type.payload

    def request(self, node_id):  # :219-227 (+ the engine's work)
        self._load_images()
        ids = (C.c_uint32 * 1)(int(node_id))
        call("kc_live_graph_request", self._h, ids, 1)

    def request_many(self, node_ids):
        self._load_images()
        ids = (C.c_uint32 * len(node_ids))(*[int(i) for i in node_ids])
        call("kc_live_graph_request", self._h, ids, len(node_ids))

    # -- the reference's request()/prioritise() only change the node's state and leave the work to
    #    the engine thread; `request` above does both at once.  These are the two halves:
    def mark_requested(self, node_id):  # request, :219-227 (state change only)
        call("kc_live_graph_mark", self._h, int(node_id), int(NodeState.Requested))

    def prioritise(self, node_id):  # :229-237 (state change only)
        call("kc_live_graph_mark", self._h, int(node_id), int(NodeState.Prioritised))

    def update(self):
        """One turn of the engine for this graph (src/engine.rs:128-183): evaluates the Requested /
        Prioritised nodes -- every non-clean node when auto_update is set -- with their dirty
        ancestors.  Returns how many nodes were wanted."""
        self._load_images()
        n = C.c_size_t()
        call("kc_live_graph_update", self._h, C.byref(n))
        return n.value

    def set_replay(self, on=True):
        """Evaluation replay: the second identical request over an unchanged graph and the same input planes is captured into
        one CUDA graph, later ones replay it (kc_live_graph_set_replay)."""
        call("kc_live_graph_set_replay", self._h, int(bool(on)))

    def replay_stats(self):
        c, r = C.c_uint64(), C.c_uint64()
        call("kc_live_graph_replay_stats", self._h, C.byref(c), C.byref(r))
        return {"captures": c.value, "replays": r.value}

    def update_turn(self):
        """ONE engine turn with priority admission (src/engine.rs:128-307, src/process_pack.rs:33-96): at most
        `set_max_processing_nodes` closest-processable nodes run, highest propagated priority first.  Returns
        their ids in the order they ran."""
        self._load_images()
        arr = (C.c_uint32 * 4096)()
        n = C.c_size_t()
        call("kc_live_graph_update_turn", self._h, arr, 4096, C.byref(n))
        return [NodeId(arr[i]) for i in range(min(n.value, 4096))]

    @staticmethod
    def await_clean_read(live_graph, node_id):  # :181-195
        live_graph._load_images()
        call("kc_live_graph_await_clean", live_graph._h, int(node_id))
        return live_graph

    await_clean_write = await_clean_read  # :164-179: same wait, a write guard in the reference


    def _id_list(self, fn, *args):
        n = C.c_size_t()
        call(fn, self._h, *args, None, 0, C.byref(n))
        arr = (C.c_uint32 * max(1, n.value))()
        call(fn, self._h, *args, arr, n.value, C.byref(n))
        return [NodeId(arr[i]) for i in range(n.value)]

    def changed_consume(self):  # :156-160
        return self._id_list("kc_live_graph_changed_consume")

    def node_ids_with_state(self, node_state):  # :270-277
        return self._id_list("kc_live_graph_node_ids_with_state", int(node_state), 0)

    def node_ids_without_state(self, node_state):  # :261-268
        return self._id_list("kc_live_graph_node_ids_with_state", int(node_state), 1)

    def get_closest_processable(self, node_id):  # :279-311
        return self._id_list("kc_live_graph_get_closest_processable", int(node_id))

    def remove_edge(self, edge):  # :551-566
        e = kc_edge(int(edge.output_id), int(edge.input_id), int(edge.output_slot), int(edge.input_slot))
        call("kc_live_graph_remove_edge", self._h, C.byref(e))
        return edge

    def rename_output_node(self, node_id, new_name):  # :625-627, returns the old name
        p = C.c_void_p()
        call("kc_live_graph_rename_output_node", self._h, int(node_id), new_name.encode(), C.byref(p))
        try:
            return C.string_at(p).decode("utf-8")
        finally:
            _lib.lib.kc_free(p)

    def new_id(self):  # :422-424
        out = C.c_uint32()
        call("kc_live_graph_new_id", self._h, C.byref(out))
        return NodeId(out.value)

    def has_node(self, node_id):  # :361-363
        self.node(node_id)

    def set_node_with_id(self, node_id, node):  # :376-387
        node.node_id = NodeId(node_id)
        self.set_node(node)

    def slot_in_memory(self, node_id, slot_id):  # :410-412: False while a plane sits in the spill queue's host memory
        v = C.c_int32()
        call("kc_live_graph_slot_in_memory", self._h, int(node_id), int(slot_id), C.byref(v))
        return bool(v.value)

    def try_buffer_rgba(self, node_id, slot_id):  # :98-125 (never blocks here: no spill queue)
        return self.buffer_rgba(node_id, slot_id)

    def try_buffer_srgba(self, node_id, slot_id):  # :127-153
        return self.buffer_srgba(node_id, slot_id)

    def cancel(self):
        call("kc_live_graph_cancel", self._h)

    def node_state(self, node_id):  # :243-249
        s = C.c_int32()
        call("kc_live_graph_node_state", self._h, int(node_id), C.byref(s))
        return NodeState(s.value)

    def slot_data(self, node_id, slot_id):  # :415-420
        im = kc_image()
        call("kc_live_graph_slot_data", self._h, int(node_id), int(slot_id), C.byref(im))
        return SlotData(node_id, slot_id, SlotImage(self._ctx, im))

    def node_slot_datas(self, node_id):  # :389-405
        arr = (C.c_uint32 * 64)()
        n = C.c_size_t()
        call("kc_live_graph_node_slot_ids", self._h, int(node_id), arr, 64, C.byref(n))
        return [self.slot_data(node_id, arr[i]) for i in range(n.value)]

    def slot_data_size(self, node_id, slot_id):  # :407-409
        w, h = C.c_uint32(), C.c_uint32()
        call("kc_live_graph_slot_data_size", self._h, int(node_id), int(slot_id), C.byref(w), C.byref(h))
        return Size(w.value, h.value)

    def buffer_rgba(self, node_id, slot_id):  # :93-95
        s = self.slot_data_size(node_id, slot_id)
        out = np.empty((s.height, s.width, 4), dtype=np.uint8)
        call("kc_live_graph_buffer_rgba", self._h, int(node_id), int(slot_id), out.ctypes.data, out.nbytes)
        return out

    def buffer_srgba(self, node_id, slot_id):  # :127-153
        s = self.slot_data_size(node_id, slot_id)
        out = np.empty((s.height, s.width, 4), dtype=np.uint8)
        call("kc_live_graph_buffer_srgba", self._h, int(node_id), int(slot_id), out.ctypes.data, out.nbytes)
        return out

    def read_rgba(self, node_id, slot_id, size, srgb=False, out=None, sync=True):
        """await_clean_read + buffer_rgba as one call; the f32->RGBA8 conversion is
        fused into the kernel that computes the node (no f32 round trip).
        sync=False (with a pinned `out`): return once everything is enqueued; the bytes
        are valid after TextureProcessor.synchronize().  The copy to the host runs on the
        context's download stream, so the next evaluation's uploads overlap it."""
        self._load_images()
        if out is None:
            out = np.empty((size.height, size.width, 4), dtype=np.uint8)
        call("kc_live_graph_read_rgba" if sync else "kc_live_graph_read_rgba_async", self._h, int(node_id), int(slot_id),
             int(bool(srgb)), out.ctypes.data, out.nbytes)
        return out

    def last_run_stats(self):
        k, g, b = C.c_uint64(), C.c_uint64(), C.c_uint64()
        call("kc_live_graph_last_run_stats", self._h, C.byref(k), C.byref(g), C.byref(b))
        return {"kernels": k.value, "fused_groups": g.value, "algorithmic_bytes": b.value}


# ---- the node bodies as plain operators (src/node/*.rs) -------------------------------
def mix(tex_pro, mix_type, left, right):
    """mix::process, src/node/mix.rs:51-134; left/right: SlotImage or None."""
    out = kc_image()
    call("kc_mix", tex_pro._ctx._h, int(mix_type), C.byref(left._im) if left is not None else None,
         C.byref(right._im) if right is not None else None, C.byref(out))
    return SlotImage(tex_pro._ctx, out)


def process_node(tex_pro, node, slot_datas, edges, embedded_slot_datas=(), input_slot_datas=()):
    """`process_node`, src/node/node_type.rs:213-248 -- THE seam the backend sits behind: one node,
    the SlotData of its connected inputs (slot_datas[i] arrived over edges[i]), the graph's embedded
    and input slot data; returns the node's output SlotDatas.  `embedded_slot_datas`: iterable of
    (embedded id, SlotData)."""
    d, keep = node._desc()
    n = len(slot_datas)
    if n != len(edges):
        raise TexProError(101, "edges.len() != slot_datas.len()")
    sd = (_lib.kc_slot_data * max(1, n))()
    ed = (kc_edge * max(1, n))()
    for i, (s_, e_) in enumerate(zip(slot_datas, edges)):
        sd[i].node_id, sd[i].slot_id, sd[i].image = int(s_.node_id), int(s_.slot_id), s_.image._im
        ed[i] = kc_edge(int(e_.output_id), int(e_.input_id), int(e_.output_slot), int(e_.input_slot))
    emb = list(embedded_slot_datas)
    em = (_lib.kc_embedded_slot_data * max(1, len(emb)))()
    for i, (eid, s_) in enumerate(emb):
        em[i].slot_data_id, em[i].slot_id, em[i].image = int(eid), int(s_.slot_id), s_.image._im
    ins = list(input_slot_datas)
    isd = (_lib.kc_slot_data * max(1, len(ins)))()
    for i, s_ in enumerate(ins):
        isd[i].node_id, isd[i].slot_id, isd[i].image = int(s_.node_id), int(s_.slot_id), s_.image._im
    out = (_lib.kc_slot_data * 8)()
    n_out = C.c_size_t()
    call("kc_process_node", tex_pro._ctx._h, C.byref(d), sd, n, em, len(emb), isd, len(ins), ed, n, out, 8, C.byref(n_out))
    return [SlotData(NodeId(out[i].node_id), SlotId(out[i].slot_id), SlotImage(tex_pro._ctx, out[i].image)) for i in range(n_out.value)]


def height_to_normal(tex_pro, image):
    """height_to_normal::process, src/node/height_to_normal.rs:16-77."""
    out = kc_image()
    call("kc_height_to_normal", tex_pro._ctx._h, C.byref(image._im), C.byref(out))
    return SlotImage(tex_pro._ctx, out)


def height_to_normal_strip(tex_pro, strip, halo_row, full_height):
    """HeightToNormal on rows [y0, y0+h) of a taller image; halo_row: Gray w x 1 image holding
    row (y0-1) mod full_height."""
    out = kc_image()
    call("kc_height_to_normal_strip", tex_pro._ctx._h, C.byref(strip._im), halo_row._im.planes[0], int(full_height), C.byref(out))
    return SlotImage(tex_pro._ctx, out)


class HaloLink:
    """A halo mailbox in peer memory (kc_halo_*): `outbox(tp, width)` on the GPU that owns the strip
    above, `.handle()` -> 64 bytes for the rank below, `HaloLink.open(tp, handle, width)` there."""

    def __init__(self, tex_pro, handle):
        self._tp, self._h = tex_pro, handle

    @staticmethod
    def outbox(tex_pro, width):
        h = C.c_void_p()
        call("kc_halo_outbox_create", tex_pro._ctx._h, int(width), C.byref(h))
        return HaloLink(tex_pro, h)

    def handle(self):
        buf = (C.c_uint8 * 64)()
        call("kc_halo_outbox_handle", self._h, buf)
        return bytes(buf)

    @staticmethod
    def open(tex_pro, handle_bytes, width):
        buf = (C.c_uint8 * 64).from_buffer_copy(handle_bytes)
        h = C.c_void_p()
        call("kc_halo_inbox_open", tex_pro._ctx._h, buf, int(width), C.byref(h))
        return HaloLink(tex_pro, h)

    def local_inbox(self, tex_pro=None):
        """The same mailbox seen from the reading side inside one process; tex_pro: the reader's context (default: the owner's)."""
        tp = tex_pro or self._tp
        h = C.c_void_p()
        call("kc_halo_inbox_local", tp._ctx._h, self._h, C.byref(h))
        return HaloLink(tp, h)

    def publish(self, image, row, step, plane=0):
        call("kc_halo_publish", self._h, image._im.planes[plane], int(row), int(step))

    def close(self):
        if self._h:
            _lib.lib.kc_halo_link_destroy(self._h)
            self._h = None


def height_to_normal_strip_peer(tex_pro, strip, inbox, step, full_height):
    """HeightToNormal on a strip whose halo row the kernel reads from the mailbox of the GPU above."""
    out = kc_image()
    call("kc_height_to_normal_strip_peer", tex_pro._ctx._h, C.byref(strip._im), inbox._h, int(step), int(full_height), C.byref(out))
    return SlotImage(tex_pro._ctx, out)


def height_to_normal_strip_exchange(tex_pro, strip, outbox, inbox, step, full_height):
    """One launch per step: the stencil kernel publishes the strip's last row into `outbox`, reads the row above out of
    `inbox` (the neighbour's mailbox, peer memory) and acknowledges it there."""
    out = kc_image()
    call("kc_height_to_normal_strip_exchange", tex_pro._ctx._h, C.byref(strip._im), outbox._h, inbox._h, int(step), int(full_height), C.byref(out))
    return SlotImage(tex_pro._ctx, out)


def halo_timeouts(tex_pro):
    n = C.c_uint32()
    call("kc_halo_timeouts", tex_pro._ctx._h, C.byref(n))
    return n.value


def resize(tex_pro, image, size, resize_filter):
    """imageops::resize per plane, src/shared.rs:155-201."""
    out = kc_image()
    call("kc_resize", tex_pro._ctx._h, C.byref(image._im), size.width, size.height, int(resize_filter), C.byref(out))
    return SlotImage(tex_pro._ctx, out)


def copy_rows(tex_pro, dst, dst_row, src, src_row, rows, plane=0):
    """Device-to-device copy of whole rows between planes of equal width (halo rows)."""
    call("kc_plane_copy_rows", tex_pro._ctx._h, dst._im.planes[plane], int(dst_row), src._im.planes[plane], int(src_row), int(rows))


def wrap_device_plane(tex_pro, device_ptr, width, height):
    """A Gray SlotImage over caller-owned device memory (e.g. torch_tensor.data_ptr())."""
    pl = C.c_void_p()
    call("kc_plane_wrap_device", tex_pro._ctx._h, int(width), int(height), C.c_void_p(int(device_ptr)), C.byref(pl))
    im = kc_image()
    im.kind, im.width, im.height = _lib.IMAGE_GRAY, int(width), int(height)
    im.planes[0] = pl
    return SlotImage(tex_pro._ctx, im)


def empty_gray(tex_pro, width, height):
    """An uninitialised Gray image in HBM (a destination for copy_rows / uploads)."""
    pl = C.c_void_p()
    call("kc_plane_create", tex_pro._ctx._h, int(width), int(height), C.byref(pl))
    im = kc_image()
    im.kind, im.width, im.height = _lib.IMAGE_GRAY, int(width), int(height)
    im.planes[0] = pl
    return SlotImage(tex_pro._ctx, im)


def pinned_empty(shape, dtype=np.float32):
    """A numpy array over page-locked host memory (kc_host_alloc), for asynchronous
    uploads/downloads.  The memory lives until free_pinned(array)."""
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    call("kc_host_alloc", max(n, 1), C.byref(p))
    buf = (C.c_char * max(n, 1)).from_address(p.value)
    a = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[a.ctypes.data] = p
    return a


def free_pinned(a):
    p = _PINNED.pop(a.ctypes.data, None)
    if p is not None:
        _lib.lib.kc_host_free(p)


_PINNED = {}


def graph_to_dict(graph_view):
    """serde-shaped dict of a NodeGraph/LiveGraph (for comparisons in tests)."""
    return _json.loads(graph_view.export_json_string())
