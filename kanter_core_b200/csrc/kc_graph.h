// kc_graph.h — host-side graph model (NodeGraph/Node/Edge) shared by the
// graph ABI and the evaluator.  Mirrors src/node_graph.rs, src/node/mod.rs,
// src/node/node_type.rs and src/edge.rs of the reference.
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "kc_internal.h"

struct KcNode {
    uint32_t node_id = 0;
    int type = KC_NODE_VALUE;
    float value = 0.0f;
    int mix_type = KC_MIX_ADD;
    std::string name;                  // Input*/Output* name, Image/Write path
    std::shared_ptr<kc_graph> graph;   // Graph payload
    uint32_t embed_id = 0;
    int policy = KC_POLICY_MOST_PIXELS;
    uint32_t policy_slot = 0, policy_w = 0, policy_h = 0;
    int filter = KC_FILTER_TRIANGLE;
    int8_t priority = 0;               // Node.priority (src/node/mod.rs:120, src/priority.rs:11-15): engine state, not serialised
};

struct kc_graph {
    std::vector<KcNode> nodes;
    std::vector<kc_edge> edges;
    uint32_t node_id_counter = 0;
};

struct KcSlotInfo {
    std::string name;
    uint32_t slot_id;
    int slot_type;
};

bool kcg_is_input(int type);
bool kcg_is_output(int type);
std::vector<KcSlotInfo> kcg_input_slots(const KcNode& n);
std::vector<KcSlotInfo> kcg_output_slots(const KcNode& n);
const KcNode* kcg_find(const kc_graph& g, uint32_t node_id);
KcNode* kcg_find(kc_graph& g, uint32_t node_id);
void kcg_from_desc(const kc_node_desc& d, KcNode& n);
void kcg_to_desc(const KcNode& n, kc_node_desc& d);

int32_t kcg_add_node(kc_graph& g, KcNode node, uint32_t* out_id);
int32_t kcg_add_node_with_id(kc_graph& g, KcNode node);
int32_t kcg_connect(kc_graph& g, uint32_t out_id, uint32_t in_id, uint32_t out_slot, uint32_t in_slot);
int32_t kcg_disconnect_slot(kc_graph& g, uint32_t node_id, int side, uint32_t slot_id, std::vector<kc_edge>* removed);
int32_t kcg_remove_node(kc_graph& g, uint32_t node_id, std::vector<kc_edge>* removed);

// PriorityPropagator::update, src/priority.rs:101-127, as its fixed point: a node's propagated priority is the
// largest of its own and its children's propagated priorities ("a node that has a high priority needs all its
// parents to have the same high priority").  out[i] belongs to g.nodes[i].
void kcg_propagated_priorities(const kc_graph& g, std::vector<int8_t>& out);

int32_t kcg_parse_json(const std::string& text, kc_graph& out);
std::string kcg_to_json(const kc_graph& g);
