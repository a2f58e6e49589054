"""ctypes binding of libkanter_b200.so (the C ABI declared in include/kanter_b200.h).

The library is the product: if it is missing this module raises, there is no
fallback of any kind (and nothing here may import the CPU oracle).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# KANTER_B200_LIB: load another build of the same library (the AddressSanitizer build, scripts/README.md)
LIB_PATH = os.environ.get("KANTER_B200_LIB") or os.path.join(HERE, "libkanter_b200.so")

# ---- enums (include/kanter_b200.h) ------------------------------------------
KC_OK = 0
ERR_NAMES = {
    1: "Generic", 2: "Canceled", 3: "Image", 4: "InvalidBufferCount", 5: "InvalidNodeId",
    6: "InvalidNodeType", 7: "InvalidSlotId", 8: "InvalidSlotType", 9: "InvalidEdge",
    10: "NoSlotData", 11: "SlotOccupied", 12: "SlotNotOccupied", 13: "UnableToLock",
    14: "NodeProcessing", 15: "PoisonError", 16: "TryLockError", 17: "NodeDirty", 18: "Io",
    19: "InvalidName", 100: "Cuda", 101: "InvalidArgument",
}
(NODE_INPUT_GRAY, NODE_INPUT_RGBA, NODE_OUTPUT_GRAY, NODE_OUTPUT_RGBA, NODE_GRAPH, NODE_IMAGE,
 NODE_EMBED, NODE_WRITE, NODE_VALUE, NODE_MIX, NODE_HEIGHT_TO_NORMAL, NODE_SEPARATE_RGBA,
 NODE_COMBINE_RGBA) = range(13)
MATH_EXACT, MATH_FAST = 0, 1
IMAGE_GRAY, IMAGE_RGBA = 0, 1
SIDE_INPUT, SIDE_OUTPUT = 0, 1


class kc_options(C.Structure):
    _fields_ = [("math_mode", C.c_int32), ("fuse", C.c_int32), ("resize_unclamped", C.c_int32), ("reserved", C.c_int32 * 5)]


class kc_image(C.Structure):
    _fields_ = [("kind", C.c_int32), ("width", C.c_uint32), ("height", C.c_uint32),
                ("planes", C.c_void_p * 4)]


class kc_slot_data(C.Structure):
    _fields_ = [("node_id", C.c_uint32), ("slot_id", C.c_uint32), ("image", kc_image)]


class kc_embedded_slot_data(C.Structure):
    _fields_ = [("slot_data_id", C.c_uint32), ("slot_id", C.c_uint32), ("image", kc_image)]


class kc_edge(C.Structure):
    _fields_ = [("output_id", C.c_uint32), ("input_id", C.c_uint32), ("output_slot", C.c_uint32),
                ("input_slot", C.c_uint32)]


class kc_node_desc(C.Structure):
    _fields_ = [("node_id", C.c_uint32), ("node_type", C.c_int32), ("value", C.c_float),
                ("mix_type", C.c_int32), ("name", C.c_char_p), ("graph", C.c_void_p),
                ("embed_id", C.c_uint32), ("resize_policy", C.c_int32), ("policy_slot", C.c_uint32),
                ("policy_width", C.c_uint32), ("policy_height", C.c_uint32),
                ("resize_filter", C.c_int32)]


class kc_slot(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("slot_id", C.c_uint32), ("slot_type", C.c_int32)]


P = C.POINTER
vp, u32, i32, sz, f32 = C.c_void_p, C.c_uint32, C.c_int32, C.c_size_t, C.c_float
u64 = C.c_uint64

# name -> (restype, argtypes); restype i32 means "status code"
SIGNATURES = {
    "kc_abi_version": (i32, []),
    "kc_last_error": (C.c_char_p, []),
    "kc_error_string": (C.c_char_p, [i32]),
    "kc_free": (None, [vp]),
    "kc_host_alloc": (i32, [sz, P(vp)]),
    "kc_host_free": (i32, [vp]),
    "kc_host_alloc_near_device": (i32, [i32, sz, P(vp)]),
    "kc_numa_info": (i32, [i32, P(i32), P(i32), P(i32)]),
    "kc_bind_thread_near_device": (i32, [i32, P(i32)]),
    "kc_context_pcie_probe": (i32, [vp, vp, sz, i32, i32, P(C.c_double), P(C.c_double)]),
    "kc_options_default": (None, [P(kc_options)]),
    "kc_context_create": (i32, [i32, P(kc_options), P(vp)]),
    "kc_context_create_on_stream": (i32, [i32, P(kc_options), vp, P(vp)]),
    "kc_context_destroy": (i32, [vp]),
    "kc_context_synchronize": (i32, [vp]),
    "kc_context_device": (i32, [vp, P(i32)]),
    "kc_context_stream": (i32, [vp, P(vp)]),
    "kc_context_set_math_mode": (i32, [vp, i32]),
    "kc_context_set_fuse": (i32, [vp, i32]),
    "kc_context_set_resize_unclamped": (i32, [vp, i32]),
    "kc_context_stats": (i32, [vp, P(u64), P(u64)]),
    "kc_context_trim": (i32, [vp]),
    "kc_plane_from_host_deferred": (i32, [vp, u32, u32, vp, P(vp)]),
    "kc_image_from_host_planes_deferred": (i32, [vp, i32, u32, u32, P(vp), P(kc_image)]),
    "kc_context_transfer_stats": (i32, [vp, P(u64), P(u64)]),
    "kc_context_set_memory_threshold": (i32, [vp, u64]),
    "kc_context_spill_stats": (i32, [vp, P(u64), P(u64), P(u64)]),
    "kc_context_set_max_processing_nodes": (i32, [vp, sz]),
    "kc_context_concurrent_begin": (i32, [vp, i32]),
    "kc_context_concurrent_end": (i32, [vp]),
    "kc_context_max_processing_nodes": (i32, [vp, P(sz)]),
    "kc_graph_set_node_priority": (i32, [vp, u32, C.c_int8]),
    "kc_graph_node_priority": (i32, [vp, u32, P(C.c_int8), P(C.c_int8)]),
    "kc_live_graph_update_turn": (i32, [vp, P(u32), sz, P(sz)]),
    "kc_live_graph_set_replay": (i32, [vp, i32]),
    "kc_live_graph_replay_stats": (i32, [vp, P(u64), P(u64)]),
    "kc_live_graph_set_priority": (i32, [vp, u32, C.c_int8]),
    "kc_plane_in_memory": (i32, [vp, P(i32)]),
    "kc_live_graph_slot_in_memory": (i32, [vp, u32, u32, P(i32)]),
    "kc_png_decode": (i32, [vp, sz, P(vp), P(u32), P(u32), P(u32)]),
    "kc_png_decode_file": (i32, [C.c_char_p, P(vp), P(u32), P(u32), P(u32)]),
    "kc_png_encode": (i32, [vp, u32, u32, u32, P(vp), P(sz)]),
    "kc_png_encode_file": (i32, [C.c_char_p, vp, u32, u32, u32]),
    "kc_debug_set_tuning": (i32, [C.c_char_p, i32]),
    "kc_debug_last_tile_config": (i32, [P(i32), P(i32), P(i32)]),
    "kc_debug_jit_compile": (i32, [P(u32), u32, i32, i32, i32, P(sz)]),
    "kc_debug_jit_wait": (i32, [i32, P(i32)]),
    "kc_context_set_timing": (i32, [vp, i32]),
    "kc_context_timing_read": (i32, [vp, i32, P(C.c_double), P(u64)]),
    "kc_event_create": (i32, [P(vp)]),
    "kc_event_destroy": (i32, [vp]),
    "kc_event_record": (i32, [vp, vp]),
    "kc_event_record_download": (i32, [vp, vp]),
    "kc_event_synchronize": (i32, [vp]),
    "kc_event_elapsed_ms": (i32, [vp, vp, P(f32)]),
    "kc_plane_create": (i32, [vp, u32, u32, P(vp)]),
    "kc_plane_from_value": (i32, [vp, u32, u32, f32, P(vp)]),
    "kc_plane_from_host": (i32, [vp, u32, u32, vp, P(vp)]),
    "kc_plane_wrap_device": (i32, [vp, u32, u32, vp, P(vp)]),
    "kc_plane_retain": (i32, [vp]),
    "kc_plane_release": (i32, [vp]),
    "kc_plane_size": (i32, [vp, P(u32), P(u32)]),
    "kc_plane_is_constant": (i32, [vp, P(i32), P(f32)]),
    "kc_plane_device_ptr": (i32, [vp, P(vp)]),
    "kc_plane_upload": (i32, [vp, vp]),
    "kc_plane_download": (i32, [vp, vp]),
    "kc_image_from_u8": (i32, [vp, vp, u32, u32, u32, P(kc_image)]),
    "kc_image_from_host_planes": (i32, [vp, i32, u32, u32, P(vp), P(kc_image)]),
    "kc_image_from_value": (i32, [vp, u32, u32, f32, i32, P(kc_image)]),
    "kc_image_as_type": (i32, [vp, P(kc_image), i32, P(kc_image)]),
    "kc_image_to_u8": (i32, [vp, P(kc_image), i32, vp]),
    "kc_image_to_u8_async": (i32, [vp, P(kc_image), i32, vp]),
    "kc_image_to_u8_device": (i32, [vp, P(kc_image), i32, vp]),
    "kc_image_download": (i32, [vp, P(kc_image), P(vp)]),
    "kc_image_materialize": (i32, [vp, P(kc_image), i32]),
    "kc_image_retain": (i32, [P(kc_image)]),
    "kc_image_release": (i32, [P(kc_image)]),
    "kc_mix": (i32, [vp, i32, P(kc_image), P(kc_image), P(kc_image)]),
    "kc_height_to_normal": (i32, [vp, P(kc_image), P(kc_image)]),
    "kc_height_to_normal_strip": (i32, [vp, P(kc_image), vp, u32, P(kc_image)]),
    "kc_halo_outbox_create": (i32, [vp, u32, P(vp)]),
    "kc_halo_outbox_handle": (i32, [vp, vp]),
    "kc_halo_inbox_open": (i32, [vp, vp, u32, P(vp)]),
    "kc_halo_inbox_local": (i32, [vp, vp, P(vp)]),
    "kc_halo_publish": (i32, [vp, vp, u32, u64]),
    "kc_height_to_normal_strip_peer": (i32, [vp, P(kc_image), vp, u64, u32, P(kc_image)]),
    "kc_height_to_normal_strip_exchange": (i32, [vp, P(kc_image), vp, vp, u64, u32, P(kc_image)]),
    "kc_halo_timeouts": (i32, [vp, P(u32)]),
    "kc_halo_link_destroy": (i32, [vp]),
    "kc_plane_copy_rows": (i32, [vp, vp, u32, vp, u32, u32]),
    "kc_resize": (i32, [vp, P(kc_image), u32, u32, i32, P(kc_image)]),
    "kc_resize_rows": (i32, [vp, P(kc_image), u32, u32, i32, u32, u32, P(kc_image)]),
    "kc_separate_rgba": (i32, [vp, P(kc_image), P(kc_image)]),
    "kc_combine_rgba": (i32, [vp, P(P(kc_image)), P(kc_image)]),
    "kc_calculate_size": (i32, [P(kc_slot_data), sz, P(kc_edge), sz, i32, u32, u32, u32, P(u32), P(u32)]),
    "kc_process_node": (i32, [vp, P(kc_node_desc), P(kc_slot_data), sz, P(kc_embedded_slot_data), sz,
                              P(kc_slot_data), sz, P(kc_edge), sz, P(kc_slot_data), sz, P(sz)]),
    "kc_graph_create": (i32, [P(vp)]),
    "kc_graph_destroy": (i32, [vp]),
    "kc_graph_clone": (i32, [vp, P(vp)]),
    "kc_graph_from_json": (i32, [C.c_char_p, P(vp)]),
    "kc_graph_from_path": (i32, [C.c_char_p, P(vp)]),
    "kc_graph_export_json": (i32, [vp, P(vp)]),
    "kc_graph_export_json_path": (i32, [vp, C.c_char_p]),
    "kc_graph_add_node": (i32, [vp, P(kc_node_desc), P(u32)]),
    "kc_graph_add_node_with_id": (i32, [vp, P(kc_node_desc)]),
    "kc_graph_remove_node": (i32, [vp, u32]),
    "kc_graph_connect": (i32, [vp, u32, u32, u32, u32]),
    "kc_graph_try_connect": (i32, [vp, u32, u32, u32, u32]),
    "kc_graph_disconnect_slot": (i32, [vp, u32, i32, u32]),
    "kc_graph_remove_edge": (i32, [vp, P(kc_edge)]),
    "kc_graph_can_connect": (i32, [vp, u32, u32, u32, u32]),
    "kc_graph_connected_edges": (i32, [vp, u32, i32, u32, P(kc_edge), sz, P(sz)]),
    "kc_graph_new_id": (i32, [vp, P(u32)]),
    "kc_graph_rename_output_node": (i32, [vp, u32, C.c_char_p, P(vp)]),
    "kc_graph_node_count": (i32, [vp, P(sz)]),
    "kc_graph_node_at": (i32, [vp, sz, P(kc_node_desc)]),
    "kc_graph_node": (i32, [vp, u32, P(kc_node_desc)]),
    "kc_graph_set_node": (i32, [vp, P(kc_node_desc)]),
    "kc_graph_edge_count": (i32, [vp, P(sz)]),
    "kc_graph_edge_at": (i32, [vp, sz, P(kc_edge)]),
    "kc_graph_input_slot_id_with_name": (i32, [vp, C.c_char_p, P(u32)]),
    "kc_graph_output_slot_id_with_name": (i32, [vp, C.c_char_p, P(u32)]),
    "kc_graph_output_ids": (i32, [vp, P(u32), sz, P(sz)]),
    "kc_graph_input_ids": (i32, [vp, P(u32), sz, P(sz)]),
    "kc_node_input_slots": (i32, [P(kc_node_desc), P(kc_slot), sz, P(sz)]),
    "kc_node_output_slots": (i32, [P(kc_node_desc), P(kc_slot), sz, P(sz)]),
    "kc_live_graph_create": (i32, [vp, P(vp)]),
    "kc_live_graph_destroy": (i32, [vp]),
    "kc_live_graph_set_node_graph": (i32, [vp, vp]),
    "kc_live_graph_node_graph": (i32, [vp, P(vp)]),
    "kc_live_graph_set_use_cache": (i32, [vp, i32]),
    "kc_live_graph_set_auto_update": (i32, [vp, i32]),
    "kc_live_graph_add_node": (i32, [vp, P(kc_node_desc), P(u32)]),
    "kc_live_graph_add_node_with_id": (i32, [vp, P(kc_node_desc)]),
    "kc_live_graph_remove_node": (i32, [vp, u32]),
    "kc_live_graph_connect": (i32, [vp, u32, u32, u32, u32]),
    "kc_live_graph_disconnect_slot": (i32, [vp, u32, i32, u32]),
    "kc_live_graph_set_node": (i32, [vp, P(kc_node_desc)]),
    "kc_live_graph_add_input_slot_data": (i32, [vp, u32, u32, P(kc_image)]),
    "kc_live_graph_clear_input_slot_data": (i32, [vp]),
    "kc_live_graph_embed_slot_data_with_id": (i32, [vp, P(kc_image), u32, u32]),
    "kc_live_graph_replace_embedded": (i32, [vp, P(kc_image), u32]),
    "kc_live_graph_set_image_data_u8": (i32, [vp, u32, vp, u32, u32, u32]),
    "kc_live_graph_request": (i32, [vp, P(u32), sz]),
    "kc_live_graph_await_clean": (i32, [vp, u32]),
    "kc_live_graph_cancel": (i32, [vp]),
    "kc_live_graph_node_state": (i32, [vp, u32, P(i32)]),
    "kc_live_graph_slot_data": (i32, [vp, u32, u32, P(kc_image)]),
    "kc_live_graph_slot_data_size": (i32, [vp, u32, u32, P(u32), P(u32)]),
    "kc_live_graph_node_slot_ids": (i32, [vp, u32, P(u32), sz, P(sz)]),
    "kc_live_graph_buffer_rgba": (i32, [vp, u32, u32, vp, sz]),
    "kc_live_graph_buffer_srgba": (i32, [vp, u32, u32, vp, sz]),
    "kc_live_graph_changed_consume": (i32, [vp, P(u32), sz, P(sz)]),
    "kc_live_graph_node_ids_with_state": (i32, [vp, i32, i32, P(u32), sz, P(sz)]),
    "kc_live_graph_get_closest_processable": (i32, [vp, u32, P(u32), sz, P(sz)]),
    "kc_live_graph_mark": (i32, [vp, u32, i32]),
    "kc_live_graph_update": (i32, [vp, P(sz)]),
    "kc_live_graph_remove_edge": (i32, [vp, P(kc_edge)]),
    "kc_live_graph_rename_output_node": (i32, [vp, u32, C.c_char_p, P(vp)]),
    "kc_live_graph_new_id": (i32, [vp, P(u32)]),
    "kc_live_graph_read_rgba": (i32, [vp, u32, u32, i32, vp, sz]),
    "kc_live_graph_read_rgba_async": (i32, [vp, u32, u32, i32, vp, sz]),
    "kc_live_graph_last_run_stats": (i32, [vp, P(u64), P(u64), P(u64)]),
}

_NO_STATUS = {"kc_abi_version", "kc_free"}


class TexProError(Exception):
    """`TexProError` (src/error.rs:5-27); `.kind` is the variant name."""

    def __init__(self, code, message=""):
        self.code = code
        self.kind = ERR_NAMES.get(code, "Unknown(%d)" % code)
        super().__init__("%s: %s" % (self.kind, message))


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "kanter_core_b200: %s is missing. Build it with `python kanter_core_b200/build.py` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == the library does not export the ABI
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(code):
    if code != KC_OK:
        msg = lib.kc_last_error()
        raise TexProError(code, msg.decode("utf-8", "replace") if msg else "")


def call(name, *args):
    """Call a status-returning entry point, raising TexProError on failure."""
    rc = getattr(lib, name)(*args)
    if name not in _NO_STATUS:
        check(rc)
    return rc
