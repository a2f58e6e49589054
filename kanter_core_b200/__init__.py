"""kanter_core_b200 — B200-native evaluation backend for kanter_core / vismut_core.

The package is the host-side mirror of the crate's interface (api.py) over the
C ABI of libkanter_b200.so (csrc/, include/kanter_b200.h).  Importing it loads
the shared library and fails loudly if it has not been built.
"""
from ._lib import MATH_EXACT, MATH_FAST, TexProError, lib  # noqa: F401
from .api import (  # noqa: F401
    Edge, EmbeddedSlotDataId, LiveGraph, MixType, Node, NodeGraph, NodeId, NodeState, NodeType, Priority,
    ResizeFilter, ResizePolicy, Side, Size, Slot, SlotData, SlotId, SlotImage, SlotType,
    HaloLink, TextureProcessor, copy_rows, empty_gray, free_pinned, graph_to_dict, halo_timeouts, height_to_normal,
    height_to_normal_strip, height_to_normal_strip_exchange, height_to_normal_strip_peer, jit_wait, mix, pinned_empty, process_node, resize, wrap_device_plane,
)
