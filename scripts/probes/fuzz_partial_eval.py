"""Memory-safety probe: evaluate only the graph outputs (use_cache off, so parents are
dropped), then ask for every node's slot data, expecting NoSlotData for the dropped
ones; free everything and repeat.  Run under faulthandler."""
import gc
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import kanter_core_b200 as kc
import oracle
from kanter_core_b200 import SlotId
from tests.test_gpu_fuzz import random_graph

tp = kc.TextureProcessor.new()
missing = found = 0
for seed in list(range(1000, 1040)) + list(range(2000, 2012)) + list(range(4000, 4024)):
    graph, embeds = random_graph(seed, 6 + seed % 17, typed=seed < 4000)
    og = oracle.from_node_graph(graph)
    for eid, planes in embeds.items():
        og.embed(eid, planes)
    try:
        og.eval()
    except oracle.OracleError:
        pass
    lg = tp.new_live_graph()
    lg.set_node_graph(graph)
    for eid, planes in embeds.items():
        lg.embed_slot_data_with_id(kc.SlotData.new(0, 0, kc.SlotImage.from_planes(tp, planes)), eid)
    try:
        for nid in lg.output_ids():
            kc.LiveGraph.await_clean_read(lg, nid)
    except kc.TexProError:
        pass
    for n in graph.nodes:
        for s in range(4):
            try:
                lg.slot_data(n.node_id, SlotId(s)).image.planes()
                found += 1
            except kc.TexProError:
                missing += 1
    del lg, og
    gc.collect()
tp.close()
print("ok found=%d missing=%d" % (found, missing))
