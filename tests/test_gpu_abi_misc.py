"""The plane / image level of the C ABI called directly (no Python mirror in between): plane life
cycle and reference counts, uploads and downloads, Separate / Combine as pure aliasing, the
asynchronous RGBA8 export, context accessors."""
import ctypes as C

import numpy as np
import pytest

import kanter_core_b200 as kc
import oracle
from kanter_core_b200._lib import call, kc_image, lib

pytestmark = pytest.mark.gpu


def rnd(seed, h, w):
    return np.random.default_rng(seed).random((h, w), dtype=np.float32)


def fptr(a):
    return a.ctypes.data_as(C.c_void_p)


def test_plane_life_cycle(tex_pro):
    ctx = tex_pro._ctx._h
    a, b = rnd(1, 20, 30), rnd(2, 20, 30)
    p = C.c_void_p()
    call("kc_plane_from_host", ctx, 30, 20, fptr(a), C.byref(p))
    w, h = C.c_uint32(), C.c_uint32()
    call("kc_plane_size", p, C.byref(w), C.byref(h))
    assert (w.value, h.value) == (30, 20)
    out = np.empty_like(a)
    call("kc_plane_download", p, fptr(out))
    assert np.array_equal(out, a)
    call("kc_plane_retain", p)
    call("kc_plane_release", p)                      # still one reference left
    q = C.c_void_p()
    call("kc_plane_create", ctx, 30, 20, C.byref(q))
    call("kc_plane_upload", q, fptr(b))
    call("kc_plane_download", q, fptr(out))
    assert np.array_equal(out, b)
    d = C.c_void_p()
    call("kc_plane_from_host_deferred", ctx, 30, 20, fptr(a), C.byref(d))
    inmem = C.c_int32(1)
    call("kc_plane_in_memory", d, C.byref(inmem))
    assert inmem.value == 0                          # still in the caller's memory
    call("kc_plane_download", d, fptr(out))
    assert np.array_equal(out, a)
    call("kc_plane_in_memory", d, C.byref(inmem))
    assert inmem.value == 1
    for x in (p, q, d):
        call("kc_plane_release", x)
    with pytest.raises(kc.TexProError):
        call("kc_plane_from_host", ctx, 30, 20, None, C.byref(p))
    assert lib.kc_last_error()                       # thread-local message of the failure above


def test_separate_combine_alias_planes_and_image_download(tex_pro):
    ctx = tex_pro._ctx._h
    P = [rnd(10 + c, 12, 9) for c in range(4)]
    img = kc.SlotImage.from_planes(tex_pro, P)
    parts = (kc_image * 4)()
    k0 = tex_pro.stats()["kernel_launches"]
    call("kc_separate_rgba", ctx, C.byref(img._im), parts)
    for c in range(4):
        assert parts[c].kind == 0 and parts[c].planes[0] == img._im.planes[c]        # the very same plane objects
    chans = (C.POINTER(kc_image) * 4)(C.pointer(parts[3]), None, C.pointer(parts[0]), None)
    comb = kc_image()
    call("kc_combine_rgba", ctx, chans, C.byref(comb))
    assert comb.kind == 1 and comb.planes[0] == img._im.planes[3] and comb.planes[2] == img._im.planes[0]
    assert tex_pro.stats()["kernel_launches"] == k0
    extra = kc_image()
    C.memmove(C.byref(extra), C.byref(comb), C.sizeof(kc_image))
    call("kc_image_retain", C.byref(extra))          # a second owner of the same planes ...
    call("kc_image_release", C.byref(extra))         # ... gives its references back (release also clears the struct)
    assert not any(extra.planes[c] for c in range(4))
    outs = [np.empty((12, 9), np.float32) for _ in range(4)]
    ptrs = (C.c_void_p * 4)(*[o.ctypes.data for o in outs])
    call("kc_image_download", ctx, C.byref(comb), ptrs)
    assert np.array_equal(outs[0], P[3]) and np.array_equal(outs[2], P[0])
    assert not outs[1].any() and np.array_equal(outs[3], np.ones((12, 9), np.float32))  # missing G -> 0, missing A -> 1
    call("kc_image_release", C.byref(comb))
    for c in range(4):
        call("kc_image_release", C.byref(parts[c]))


def test_async_export_and_context_accessors(tex_pro):
    ctx = tex_pro._ctx._h
    dev, stream = C.c_int32(-1), C.c_void_p()
    call("kc_context_device", ctx, C.byref(dev))
    call("kc_context_stream", ctx, C.byref(stream))
    assert dev.value == 0 and stream.value
    P = [rnd(30 + c, 64, 64) for c in range(4)]
    img = kc.SlotImage.from_planes(tex_pro, P)
    host = kc.pinned_empty((64, 64, 4), np.uint8)
    call("kc_image_to_u8_async", ctx, C.byref(img._im), 0, host.ctypes.data)
    call("kc_context_synchronize", ctx)
    assert np.array_equal(host, oracle.to_u8(P, False))
    kc.free_pinned(host)
    call("kc_context_trim", ctx)                     # recycled buffers go back to the driver; everything still works
    assert np.array_equal(img.planes()[1], P[1])


def test_clear_input_slot_data(tex_pro):
    from kanter_core_b200 import Node, NodeType, SlotData, SlotId
    H = rnd(40, 8, 8)
    lg = tex_pro.new_live_graph()
    i = lg.add_node(Node.new(NodeType.InputGray("in")))
    o = lg.add_node(Node.new(NodeType.OutputGray("out")))
    lg.connect(i, o, SlotId(0), SlotId(0))
    lg.add_input_slot_data(SlotData.new(i, 0, kc.SlotImage.from_planes(tex_pro, [H])))
    kc.LiveGraph.await_clean_read(lg, o)
    assert np.array_equal(lg.slot_data(o, SlotId(0)).image.planes()[0], H)
    call("kc_live_graph_clear_input_slot_data", lg._h)
    assert lg.node_state(i) == kc.NodeState.Dirty and lg.node_state(o) == kc.NodeState.Dirty
    # an Input node without data yields no buffers (src/node/input_gray.rs:7-16) -> InvalidBufferCount
    # (node_type.rs:124-137; the reference's engine shuts down on it, src/engine.rs:104-120)
    with pytest.raises(kc.TexProError) as e:
        lg.request(o)
    assert e.value.kind == "InvalidBufferCount"


def test_images_and_graphs_may_outlive_their_context():
    """The reference's SlotImages and LiveGraphs are Arc-owned and outlive the Engine;
    here planes and live graphs keep the context's bookkeeping alive, so releasing them
    after kc_context_destroy is legal, and using them for new work is an error."""
    ctx = C.c_void_p()
    call("kc_context_create", 0, None, C.byref(ctx))
    a = np.random.default_rng(0).random((64, 48), dtype=np.float32)
    im = kc_image()
    ptrs = (C.c_void_p * 4)(a.ctypes.data, 0, 0, 0)
    call("kc_image_from_host_planes", ctx, 0, 48, 64, ptrs, C.byref(im))
    lazy = kc_image()
    call("kc_mix", ctx, 2, C.byref(im), C.byref(im), C.byref(lazy))     # an unevaluated expression over `im`
    lg = C.c_void_p()
    call("kc_live_graph_create", ctx, C.byref(lg))
    call("kc_context_destroy", ctx)
    out = np.empty((64, 48), np.float32)
    optrs = (C.c_void_p * 4)(out.ctypes.data, 0, 0, 0)
    with pytest.raises(kc.TexProError):
        call("kc_image_download", ctx, C.byref(lazy), optrs)                   # needs a new plane: refused
    call("kc_image_release", C.byref(lazy))
    call("kc_image_release", C.byref(im))
    call("kc_live_graph_destroy", lg)                                          # the last handle: the context goes here


def test_python_objects_released_after_close():
    tp = kc.TextureProcessor.new()
    img = kc.SlotImage.from_planes(tp, [rnd(3, 16, 16)])
    lg = tp.new_live_graph()
    n = lg.add_node(kc.Node.new(kc.NodeType.Value(0.5)))
    kc.LiveGraph.await_clean_read(lg, n)
    held = lg.slot_data(n, kc.SlotId(0)).image
    tp.close()
    del img, held, lg            # releases run against a closed context: no leak, no crash
    import gc
    gc.collect()


def test_context_on_the_callers_stream():
    """kc_context_create_on_stream: the library's kernels go onto a stream the caller owns (here a
    torch stream), so they are ordered with the caller's own work on it and a torch event recorded
    after the call covers them."""
    import torch
    s = torch.cuda.Stream()
    tp = kc.TextureProcessor.new(cuda_stream=s.cuda_stream)
    st = C.c_void_p()
    call("kc_context_stream", tp._ctx._h, C.byref(st))
    assert (st.value or 0) == s.cuda_stream
    n = 1024
    a, b = rnd(5, n, n), rnd(6, n, n)
    # a torch tensor produced on the same stream, wrapped as a plane without a copy, consumed by the library
    with torch.cuda.stream(s):
        ta = torch.from_numpy(a).cuda(non_blocking=False)
        ta = ta * 2.0                                            # caller's own kernel on s
    A = kc.wrap_device_plane(tp, ta.data_ptr(), n, n)
    B = kc.SlotImage.from_planes(tp, [b])
    out = kc.mix(tp, kc.MixType.Add, A, B)
    call("kc_image_materialize", tp._ctx._h, C.byref(out._im), 0)   # enqueued on s, after the torch kernel
    done = torch.cuda.Event()
    done.record(s)
    done.synchronize()
    got = out.planes()[0]
    assert np.array_equal(got, oracle.mix_plane(0, (a * np.float32(2.0)).astype(np.float32), b))
    del A, B, out
    tp.close()
    torch.cuda.synchronize()
    with torch.cuda.stream(s):                                  # the stream is still the caller's, and usable
        assert float((ta + 1).sum()) > 0


def test_remaining_reference_surface(tex_pro):
    """SlotImage::from_buffers_rgb/rgba, from_self, SlotData::in_memory/from_self, TextureProcessor::buffer_rgba and
    await_slot_data_size (src/slot_image.rs:66-114, src/slot_data.rs:62-78, src/texture_processor.rs:75-105)."""
    r, g, b, a = (rnd(50 + k, 12, 10) for k in range(4))
    rgb = kc.SlotImage.from_buffers_rgb(tex_pro, [r, g, b])
    got = rgb.planes()
    assert rgb.is_rgba() and np.array_equal(got[0], r) and np.array_equal(got[2], b) and np.array_equal(got[3], np.ones_like(r))
    rgba = kc.SlotImage.from_buffers_rgba(tex_pro, [r, g, b, a])
    assert np.array_equal(rgba.planes()[3], a)
    for bad, n in ((kc.SlotImage.from_buffers_rgb, 4), (kc.SlotImage.from_buffers_rgba, 3)):
        with pytest.raises(kc.TexProError) as e:
            bad(tex_pro, [r] * n)
        assert e.value.kind == "InvalidBufferCount"
    twin = rgba.from_self()
    assert np.array_equal(twin.planes()[1], g) and twin.in_memory()
    sd = kc.SlotData.new(3, 1, rgba).from_self()
    assert int(sd.node_id) == 3 and int(sd.slot_id) == 1 and sd.in_memory() and sd.size() == kc.Size(10, 12)
    lg = tex_pro.new_live_graph()
    v = lg.add_node(kc.Node.new(kc.NodeType.Value(0.5)))
    o = lg.add_node(kc.Node.new(kc.NodeType.OutputGray("o")))
    lg.connect(v, o, kc.SlotId(0), kc.SlotId(0))
    assert kc.TextureProcessor.await_slot_data_size(lg, o, kc.SlotId(0)) == kc.Size(1, 1)
    px = np.asarray(kc.TextureProcessor.buffer_rgba(lg, o, kc.SlotId(0))).reshape(-1)
    assert list(px[:4]) == [127, 127, 127, 255]            # read_dirty_read's known answer for Value(0.5)
    assert tex_pro.processing_node_count() == 0
    tex_pro.set_max_processing_nodes(4)


def test_mix_rejects_operands_of_another_size_whatever_their_kind(tex_pro):
    """kc_mix through the ABI: the reference reads pixel (x, y) of both operands for every pixel of the LEFT image
    (src/node/mix.rs:136-192, get_pixel panics out of bounds).  A constant left of 64x64 with a 16x16 device or lazy
    right used to launch a 64x64 kernel over the 16x16 buffer (advisor finding, round 1)."""
    from kanter_core_b200 import MixType, Size
    from kanter_core_b200._lib import TexProError
    small = kc.SlotImage.from_planes(tex_pro, [np.full((16, 16), 0.5, np.float32)])
    big_const = kc.SlotImage.from_value(tex_pro, Size(64, 64), 0.25, False)
    lazy_small = kc.mix(tex_pro, MixType.Add, small, small)
    for l, r in ((big_const, small), (big_const, lazy_small), (small, big_const), (lazy_small, big_const)):
        with pytest.raises(TexProError):
            kc.mix(tex_pro, MixType.Multiply, l, r)
    same = kc.SlotImage.from_value(tex_pro, Size(16, 16), 0.25, False)
    got = kc.mix(tex_pro, MixType.Multiply, same, lazy_small).planes()[0]
    assert np.array_equal(got, np.full((16, 16), 0.25, np.float32))


def test_copy_rows_bounds_do_not_wrap(tex_pro):
    """dst_row + rows was formed in 32 bits: 0xffffffff + 2 passed the check (advisor finding, round 1)."""
    from kanter_core_b200._lib import TexProError
    a = kc.SlotImage.from_planes(tex_pro, [np.zeros((8, 8), np.float32)])
    b = kc.empty_gray(tex_pro, 8, 8)
    for args in ((0xffffffff, 0, 2), (0, 0xffffffff, 2), (7, 0, 2), (0, 7, 2), (9, 0, 0), (0, 0, 0xffffffff)):
        with pytest.raises(TexProError):
            kc.copy_rows(tex_pro, b, args[0], a, args[1], args[2])
    kc.copy_rows(tex_pro, b, 6, a, 0, 2)
    kc.copy_rows(tex_pro, b, 8, a, 8, 0)      # an empty range at the end is legal
