"""ctypes binding of the CPU oracle (oracle/libkanter_oracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkanter_oracle.so")


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("kanter_oracle.cpp", "kanter_oracle.h", "Makefile")]
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", HERE, "-B", "libkanter_oracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


build()
_lib = C.CDLL(LIB_PATH)
_f = C.POINTER(C.c_float)
vp, u32, i32, u64 = C.c_void_p, C.c_uint32, C.c_int, C.c_uint64
_lib.ko_deconstruct_u8.argtypes = [vp, u32, u32, u32, vp, vp, vp, vp]
_lib.ko_to_u8.argtypes = [vp, vp, vp, vp, u64, i32, vp]
_lib.ko_mix_plane.argtypes = [i32, vp, vp, u64, vp]
_lib.ko_rgb_to_gray.argtypes = [vp, vp, vp, u64, vp]
_lib.ko_height_to_normal.argtypes = [vp, u32, u32, vp, vp, vp]
_lib.ko_height_to_normal_strip.argtypes = [vp, u32, u32, u32, vp, vp, vp, vp]
_lib.ko_resize_plane.argtypes = [vp, u32, u32, vp, u32, u32, i32]
_lib.ko_resize_weights.argtypes = [u32, u32, i32, vp, vp, vp, u32]
_lib.ko_resize_weights.restype = u32
_lib.ko_graph_new.restype = vp
_lib.ko_graph_free.argtypes = [vp]
_lib.ko_graph_add_node.argtypes = [vp, u32, i32, C.c_float, i32, C.c_char_p, vp, u32, i32, u32, u32, u32, i32]
_lib.ko_graph_add_edge.argtypes = [vp, u32, u32, u32, u32]
_lib.ko_graph_set_image_u8.argtypes = [vp, u32, vp, u32, u32, u32]
_lib.ko_graph_add_input_f32.argtypes = [vp, u32, i32, u32, u32, C.POINTER(vp)]
_lib.ko_graph_embed_f32.argtypes = [vp, u32, i32, u32, u32, C.POINTER(vp)]
_lib.ko_graph_eval.argtypes = [vp, i32]
_lib.ko_graph_slot.argtypes = [vp, u32, u32, C.POINTER(i32), C.POINTER(u32), C.POINTER(u32), C.POINTER(vp)]
_lib.ko_graph_slot_ids.argtypes = [vp, u32, C.POINTER(u32), i32]
_lib.ko_graph_eval_batch.argtypes = [vp, i32, i32]
_lib.ko_graph_eval_batch.restype = C.c_double

ERR = {1: "Generic", 2: "Canceled", 3: "Image", 4: "InvalidBufferCount", 5: "InvalidNodeId", 6: "InvalidNodeType",
       7: "InvalidSlotId", 8: "InvalidSlotType", 9: "InvalidEdge", 10: "NoSlotData", 11: "SlotOccupied",
       12: "SlotNotOccupied", 13: "UnableToLock", 14: "NodeProcessing", 15: "PoisonError", 16: "TryLockError",
       17: "NodeDirty", 18: "Io", 19: "InvalidName"}


class OracleError(Exception):
    def __init__(self, code):
        self.code = code
        self.kind = ERR.get(code, str(code))
        super().__init__(self.kind)


def _c(a, dt=np.float32):
    return np.ascontiguousarray(a, dtype=dt)


def deconstruct_u8(samples):
    a = _c(samples, np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, ch = a.shape
    out = [np.empty((h, w), np.float32) for _ in range(4)]
    _lib.ko_deconstruct_u8(a.ctypes.data, w, h, ch, *[o.ctypes.data for o in out])
    return out


def to_u8(planes, srgb=False):
    planes = [_c(p) for p in planes]
    h, w = planes[0].shape
    out = np.empty((h, w, 4), np.uint8)
    ptrs = [p.ctypes.data for p in planes] + [None] * (4 - len(planes))
    _lib.ko_to_u8(ptrs[0], ptrs[1], ptrs[2], ptrs[3], h * w, int(srgb), out.ctypes.data)
    return out


def mix_plane(op, l, r):
    l, r = _c(l), _c(r)
    out = np.empty_like(l)
    _lib.ko_mix_plane(int(op), l.ctypes.data, r.ctypes.data, l.size, out.ctypes.data)
    return out


def rgb_to_gray(r, g, b):
    r, g, b = _c(r), _c(g), _c(b)
    out = np.empty_like(r)
    _lib.ko_rgb_to_gray(r.ctypes.data, g.ctypes.data, b.ctypes.data, r.size, out.ctypes.data)
    return out


def height_to_normal(hgt):
    hgt = _c(hgt)
    h, w = hgt.shape
    out = [np.empty((h, w), np.float32) for _ in range(3)]
    _lib.ko_height_to_normal(hgt.ctypes.data, w, h, *[o.ctypes.data for o in out])
    return out


def height_to_normal_strip(strip, h_full, halo_row):
    strip, halo_row = _c(strip), _c(halo_row)
    h, w = strip.shape
    out = [np.empty((h, w), np.float32) for _ in range(3)]
    _lib.ko_height_to_normal_strip(strip.ctypes.data, w, h, h_full, halo_row.ctypes.data, *[o.ctypes.data for o in out])
    return out


def resize_plane(src, w, h, filt):
    src = _c(src)
    sh, sw = src.shape
    out = np.empty((h, w), np.float32)
    _lib.ko_resize_plane(src.ctypes.data, sw, sh, out.ctypes.data, w, h, int(filt))
    return out


def resize_weights(src_len, dst_len, filt):
    mt = _lib.ko_resize_weights(src_len, dst_len, int(filt), None, None, None, 0)
    left = np.zeros(dst_len, np.uint32)
    count = np.zeros(dst_len, np.uint32)
    wts = np.zeros((dst_len, mt), np.float32)
    _lib.ko_resize_weights(src_len, dst_len, int(filt), left.ctypes.data, count.ctypes.data, wts.ctypes.data, mt)
    return left, count, wts


class Graph:
    """The oracle's graph evaluator, fed from plain node/edge descriptions."""

    def __init__(self):
        self._h = _lib.ko_graph_new()

    def __del__(self):
        if getattr(self, "_h", None):
            _lib.ko_graph_free(self._h)
            self._h = None

    def add_node(self, node_id, node_type, value=0.0, mix_type=0, name="", nested=None, embed_id=0,
                 policy=0, policy_slot=0, policy_w=0, policy_h=0, filt=1):
        rc = _lib.ko_graph_add_node(self._h, node_id, node_type, value, mix_type, name.encode(),
                                    nested._h if nested is not None else None, embed_id, policy, policy_slot,
                                    policy_w, policy_h, filt)
        if rc:
            raise OracleError(rc)

    def add_edge(self, output_id, input_id, output_slot, input_slot):
        _lib.ko_graph_add_edge(self._h, output_id, input_id, output_slot, input_slot)

    def set_image_u8(self, node_id, samples):
        a = _c(samples, np.uint8)
        if a.ndim == 2:
            a = a[:, :, None]
        _lib.ko_graph_set_image_u8(self._h, node_id, a.ctypes.data, a.shape[1], a.shape[0], a.shape[2])

    def _planes(self, planes):
        arrs = [_c(p) for p in planes]
        ptrs = (vp * len(arrs))(*[a.ctypes.data for a in arrs])
        return arrs, ptrs

    def add_input(self, node_id, planes):
        arrs, ptrs = self._planes(planes)
        _lib.ko_graph_add_input_f32(self._h, node_id, int(len(arrs) == 4), arrs[0].shape[1], arrs[0].shape[0], ptrs)

    def embed(self, embed_id, planes):
        arrs, ptrs = self._planes(planes)
        rc = _lib.ko_graph_embed_f32(self._h, embed_id, int(len(arrs) == 4), arrs[0].shape[1], arrs[0].shape[0], ptrs)
        if rc:
            raise OracleError(rc)

    def eval(self, max_threads=1):
        rc = _lib.ko_graph_eval(self._h, max_threads)
        if rc:
            raise OracleError(rc)

    def eval_batch_seconds(self, copies, max_threads):
        t = _lib.ko_graph_eval_batch(self._h, copies, max_threads)
        if t < 0:
            raise OracleError(int(-t))
        return t

    def slot(self, node_id, slot_id):
        """-> list of (h, w) float32 planes (1 Gray / 4 Rgba), copied."""
        rgba, w, h = i32(), u32(), u32()
        ptrs = (vp * 4)()
        rc = _lib.ko_graph_slot(self._h, node_id, slot_id, C.byref(rgba), C.byref(w), C.byref(h), ptrs)
        if rc:
            raise OracleError(rc)
        out = []
        for c in range(4 if rgba.value else 1):
            # planes of one image may differ in size only through aliasing of 1x1 defaults; size() is plane 0's
            buf = (C.c_float * (w.value * h.value)).from_address(ptrs[c])
            out.append(np.frombuffer(buf, dtype=np.float32).reshape(h.value, w.value).copy())
        return out

    def slot_ids(self, node_id):
        arr = (u32 * 64)()
        n = _lib.ko_graph_slot_ids(self._h, node_id, arr, 64)
        return [arr[i] for i in range(n)]

    def buffer_rgba(self, node_id, slot_id, srgb=False):
        return to_u8(self.slot(node_id, slot_id), srgb)


def from_node_graph(node_graph, images=None):
    """Build an oracle Graph from a kanter_core_b200 NodeGraph / LiveGraph description
    (host-side data only).  `images`: {node_id: decoded u8 array} for Image nodes."""
    g = Graph()
    for n in node_graph.nodes:
        t = n.node_type
        kw = dict(policy=n.resize_policy.kind, policy_slot=int(n.resize_policy.slot),
                  policy_w=n.resize_policy.size.width, policy_h=n.resize_policy.size.height,
                  filt=int(n.resize_filter))
        if t.kind == 8:
            kw["value"] = t.payload
        elif t.kind == 9:
            kw["mix_type"] = int(t.payload)
        elif t.kind == 6:
            kw["embed_id"] = int(t.payload)
        elif t.kind == 4:
            kw["nested"] = from_node_graph(t.payload)
        elif t.payload is not None:
            kw["name"] = t.payload
        g.add_node(int(n.node_id), t.kind, **kw)
    for e in node_graph.edges:
        g.add_edge(int(e.output_id), int(e.input_id), int(e.output_slot), int(e.input_slot))
    for nid, px in (images or {}).items():
        g.set_image_u8(int(nid), px)
    return g
