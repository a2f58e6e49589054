"""BASELINE.json configs[4] -- the 32-node batch graph (Separate/Mix/HeightToNormal/Resize/
Combine + nested Graph) -- at sizes the CPU oracle finishes in seconds: oracle on the CPU,
product on the GPU, bit-exact in EXACT mode, 1e-5/1e-6 in FAST mode."""
import numpy as np
import pytest

import kanter_core_b200 as kc
from kanter_core_b200 import SlotId
from tests import graphs


def test_config5_graph_shape_and_oracle_runs():
    g, out = graphs.config5_graph(64)
    kinds = [n.node_type.kind for n in g.nodes]
    assert len(kinds) == 32
    planes = graphs.config5_oracle(g, out, graphs.config5_inputs(100, 64))
    assert len(planes) == 4 and planes[0].shape == (64, 64)
    assert np.array_equal(planes[3], np.ones((64, 64), np.float32))   # Mix-rgba: alpha := 1
    assert all(np.isfinite(p).all() for p in planes)


@pytest.mark.gpu
@pytest.mark.parametrize("size,seed", [(64, 100), (256, 101), (500, 102)])
def test_config5_bit_exact(tex_pro, size, seed):
    g, out = graphs.config5_graph(size)
    inputs = graphs.config5_inputs(seed, size)
    want = graphs.config5_oracle(g, out, inputs)
    lg = graphs.config5_product(tex_pro, g, out, inputs)
    got = lg.slot_data(out, SlotId(0)).image.planes()
    for c in range(4):
        assert np.array_equal(got[c].view(np.uint32), want[c].view(np.uint32)), "plane %d" % c
    st = lg.last_run_stats()
    assert 0 < st["kernels"] < 32          # fusion: far fewer launches than nodes
    # and the fused-export path gives the oracle's bytes
    import oracle
    lg2 = graphs.config5_product(tex_pro, g, out, inputs, read=False)
    assert np.array_equal(lg2.read_rgba(out, SlotId(0), kc.Size(size, size)), oracle.to_u8(want, False))


@pytest.mark.gpu
def test_config5_fast_within_tolerance(tex_pro_fast):
    size = 256
    g, out = graphs.config5_graph(size)
    inputs = graphs.config5_inputs(103, size)
    want = graphs.config5_oracle(g, out, inputs)
    got = graphs.config5_product(tex_pro_fast, g, out, inputs).slot_data(out, SlotId(0)).image.planes()
    for c in range(4):
        w = want[c].astype(np.float64)
        assert (np.abs(got[c].astype(np.float64) - w) <= 1e-6 + 1e-5 * np.abs(w)).all(), "plane %d" % c
