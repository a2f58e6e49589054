// kc_graph.cu — NodeGraph: nodes, edges, slot tables, connect rules, JSON.
// Host-only code.  Mirrors src/node_graph.rs, src/node/mod.rs:197-238,
// src/node/node_type.rs:56-96,141-211 and the serde schema of
// data/invert_graph.json.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "kc_graph.h"

// ---------------------------------------------------------------------------
// slot tables: Node::input_slots / output_slots, src/node/node_type.rs:141-211
// ---------------------------------------------------------------------------
bool kcg_is_input(int t) { return t == KC_NODE_INPUT_GRAY || t == KC_NODE_INPUT_RGBA; }
bool kcg_is_output(int t) { return t == KC_NODE_OUTPUT_GRAY || t == KC_NODE_OUTPUT_RGBA; }

static int slot_type_of(int node_type) {  // NodeType::to_slot_type, :89-95
    switch (node_type) {
        case KC_NODE_INPUT_GRAY: case KC_NODE_OUTPUT_GRAY: return KC_SLOT_GRAY;
        default: return KC_SLOT_RGBA;
    }
}

std::vector<KcSlotInfo> kcg_input_slots(const KcNode& n) {
    std::vector<KcSlotInfo> s;
    switch (n.type) {
        case KC_NODE_OUTPUT_GRAY: s.push_back({"input", 0, KC_SLOT_GRAY}); break;
        case KC_NODE_OUTPUT_RGBA: s.push_back({"input", 0, KC_SLOT_RGBA}); break;
        case KC_NODE_GRAPH:  // NodeGraph::input_slots, src/node_graph.rs:299-313
            if (n.graph)
                for (const KcNode& m : n.graph->nodes)
                    if (kcg_is_input(m.type)) s.push_back({m.name, m.node_id, slot_type_of(m.type)});
            break;
        case KC_NODE_MIX:
            s.push_back({"left", 0, KC_SLOT_GRAY_OR_RGBA});
            s.push_back({"right", 1, KC_SLOT_GRAY_OR_RGBA});
            break;
        case KC_NODE_HEIGHT_TO_NORMAL: s.push_back({"input", 0, KC_SLOT_GRAY}); break;
        case KC_NODE_SEPARATE_RGBA: s.push_back({"input", 0, KC_SLOT_RGBA}); break;
        case KC_NODE_COMBINE_RGBA:
            s.push_back({"red", 0, KC_SLOT_GRAY});
            s.push_back({"green", 1, KC_SLOT_GRAY});
            s.push_back({"blue", 2, KC_SLOT_GRAY});
            s.push_back({"alpha", 3, KC_SLOT_GRAY});
            break;
        case KC_NODE_WRITE:  // `unimplemented!()` in the reference; a Write node takes one image
            s.push_back({"input", 0, KC_SLOT_GRAY_OR_RGBA});
            break;
        default: break;  // InputGray, InputRgba, Image, Embed, Value: no inputs
    }
    return s;
}

std::vector<KcSlotInfo> kcg_output_slots(const KcNode& n) {
    std::vector<KcSlotInfo> s;
    switch (n.type) {
        case KC_NODE_INPUT_GRAY: s.push_back({"output", 0, KC_SLOT_GRAY}); break;
        case KC_NODE_INPUT_RGBA: s.push_back({"output", 0, KC_SLOT_RGBA}); break;
        case KC_NODE_GRAPH:  // NodeGraph::output_slots, src/node_graph.rs:315-330
            if (n.graph)
                for (const KcNode& m : n.graph->nodes)
                    if (kcg_is_output(m.type)) s.push_back({m.name, m.node_id, slot_type_of(m.type)});
            break;
        case KC_NODE_IMAGE: case KC_NODE_EMBED: s.push_back({"output", 0, KC_SLOT_RGBA}); break;
        case KC_NODE_VALUE: s.push_back({"output", 0, KC_SLOT_GRAY}); break;
        case KC_NODE_MIX: s.push_back({"output", 0, KC_SLOT_GRAY_OR_RGBA}); break;
        case KC_NODE_HEIGHT_TO_NORMAL: s.push_back({"output", 0, KC_SLOT_RGBA}); break;
        case KC_NODE_SEPARATE_RGBA:
            s.push_back({"red", 0, KC_SLOT_GRAY});
            s.push_back({"green", 1, KC_SLOT_GRAY});
            s.push_back({"blue", 2, KC_SLOT_GRAY});
            s.push_back({"alpha", 3, KC_SLOT_GRAY});
            break;
        case KC_NODE_COMBINE_RGBA: s.push_back({"output", 0, KC_SLOT_RGBA}); break;
        default: break;  // OutputGray, OutputRgba, Write: no outputs
    }
    return s;
}

// SlotType::fits, src/node/mod.rs:210-220
static bool slot_fits(int self, int other) {
    switch (self) {
        case KC_SLOT_GRAY: return other == KC_SLOT_GRAY || other == KC_SLOT_GRAY_OR_RGBA;
        case KC_SLOT_RGBA: return other == KC_SLOT_RGBA || other == KC_SLOT_GRAY_OR_RGBA;
        default: return true;
    }
}

const KcNode* kcg_find(const kc_graph& g, uint32_t id) {
    for (const KcNode& n : g.nodes)
        if (n.node_id == id) return &n;
    return nullptr;
}
KcNode* kcg_find(kc_graph& g, uint32_t id) {
    for (KcNode& n : g.nodes)
        if (n.node_id == id) return &n;
    return nullptr;
}

void kcg_from_desc(const kc_node_desc& d, KcNode& n) {
    n.node_id = d.node_id;
    n.type = d.node_type;
    n.value = d.value;
    n.mix_type = d.mix_type;
    n.name = d.name ? d.name : "";
    n.graph = d.graph ? std::make_shared<kc_graph>(*d.graph) : nullptr;
    n.embed_id = d.embed_id;
    n.policy = d.resize_policy;
    n.policy_slot = d.policy_slot;
    n.policy_w = d.policy_width;
    n.policy_h = d.policy_height;
    n.filter = d.resize_filter;
}

void kcg_to_desc(const KcNode& n, kc_node_desc& d) {
    d.node_id = n.node_id;
    d.node_type = n.type;
    d.value = n.value;
    d.mix_type = n.mix_type;
    d.name = n.name.c_str();
    d.graph = n.graph.get();
    d.embed_id = n.embed_id;
    d.resize_policy = n.policy;
    d.policy_slot = n.policy_slot;
    d.policy_width = n.policy_w;
    d.policy_height = n.policy_h;
    d.resize_filter = n.filter;
}

void kcg_propagated_priorities(const kc_graph& g, std::vector<int8_t>& out) {
    const size_t N = g.nodes.size();
    out.resize(N);
    for (size_t i = 0; i < N; ++i) out[i] = g.nodes[i].priority;
    bool any = false;
    for (size_t i = 0; i < N && !any; ++i) any = out[i] != 0;
    if (!any || g.edges.empty()) return;
    std::map<uint32_t, size_t> pos;
    for (size_t i = 0; i < N; ++i) pos.emplace(g.nodes[i].node_id, i);
    // relax parent >= child until nothing moves: at most N sweeps on a DAG, and a cycle cannot raise anything forever
    for (size_t sweep = 0; sweep <= N; ++sweep) {
        bool moved = false;
        for (const kc_edge& e : g.edges) {
            auto p = pos.find(e.output_id), c = pos.find(e.input_id);
            if (p == pos.end() || c == pos.end()) continue;
            if (out[c->second] > out[p->second]) { out[p->second] = out[c->second]; moved = true; }
        }
        if (!moved) break;
    }
}

// ---------------------------------------------------------------------------
// node management: src/node_graph.rs:81-96,141-197,332-348
// ---------------------------------------------------------------------------
static uint32_t new_id(kc_graph& g) {  // NodeGraph::new_id, :84-96
    uint32_t out = g.node_id_counter++;
    while (kcg_find(g, out)) out = g.node_id_counter++;
    return out;
}

// NodeGraph::avoid_name_collision, :141-164
static std::string avoid_name_collision(const std::vector<std::string>& names, const std::string& name) {
    std::string edit = name;
    auto contains = [&](const std::string& s) { return std::find(names.begin(), names.end(), s) != names.end(); };
    while (contains(edit)) {
        size_t us = edit.rfind('_');
        if (us != std::string::npos) {
            std::string base = edit.substr(0, us), num = edit.substr(us + 1);
            bool numeric = std::all_of(num.begin(), num.end(), [](unsigned char ch) { return std::isdigit(ch); });
            if (numeric) {  // note: an empty suffix is "all numeric" and fails to parse -> 0
                uint32_t v = 0;
                bool ok = !num.empty();
                unsigned long long acc = 0;
                for (char ch : num) {
                    acc = acc * 10 + (unsigned)(ch - '0');
                    if (acc > 0xffffffffull) { ok = false; break; }
                }
                v = ok ? (uint32_t)acc + 1u : 0u;  // wrapping_add(1)
                edit = base + "_" + std::to_string(v);
            } else {
                edit = base + "_0";
            }
        } else {
            edit = edit + "_0";
        }
    }
    return edit;
}

static int32_t add_node_internal(kc_graph& g, KcNode node, uint32_t id) {  // :166-190
    if (kcg_is_input(node.type) || kcg_is_output(node.type)) {
        if (node.name.empty()) node.name = "untitled";
        std::vector<std::string> names;
        const bool in = kcg_is_input(node.type);
        for (const KcNode& n : g.nodes)
            if (in ? kcg_is_input(n.type) : kcg_is_output(n.type)) names.push_back(n.name);
        node.name = avoid_name_collision(names, node.name);
    }
    if (node.type == KC_NODE_GRAPH && !node.graph) KC_FAIL(KC_ERR_INVALID_NODE_TYPE, "Graph node without a NodeGraph payload");
    if (node.type < KC_NODE_INPUT_GRAY || node.type > KC_NODE_COMBINE_RGBA) KC_FAIL(KC_ERR_INVALID_NODE_TYPE, "unknown node type %d", node.type);
    node.node_id = id;
    g.nodes.push_back(std::move(node));
    return KC_OK;
}

int32_t kcg_add_node(kc_graph& g, KcNode node, uint32_t* out_id) {
    uint32_t id = new_id(g);
    KC_TRY(add_node_internal(g, std::move(node), id));
    if (out_id) *out_id = id;
    return KC_OK;
}

int32_t kcg_add_node_with_id(kc_graph& g, KcNode node) {
    if (kcg_find(g, node.node_id)) KC_FAIL(KC_ERR_INVALID_NODE_ID, "node id %u already in use", node.node_id);
    uint32_t id = node.node_id;
    return add_node_internal(g, std::move(node), id);
}

static int32_t slot_type_lookup(const std::vector<KcSlotInfo>& slots, uint32_t slot_id, int* type) {
    for (const KcSlotInfo& s : slots)
        if (s.slot_id == slot_id) { *type = s.slot_type; return KC_OK; }
    KC_FAIL(KC_ERR_INVALID_SLOT_ID, "no slot with id %u", slot_id);
}

int32_t kcg_disconnect_slot(kc_graph& g, uint32_t node_id, int side, uint32_t slot_id, std::vector<kc_edge>* removed) {
    // NodeGraph::disconnect_slot, :500-520
    if (!kcg_find(g, node_id)) KC_FAIL(KC_ERR_INVALID_NODE_ID, "no node %u", node_id);
    size_t before = g.edges.size();
    std::vector<kc_edge> keep;
    for (const kc_edge& e : g.edges) {
        bool hit = side == KC_SIDE_INPUT ? (e.input_id == node_id && e.input_slot == slot_id)
                                         : (e.output_id == node_id && e.output_slot == slot_id);
        if (hit) { if (removed) removed->push_back(e); }
        else keep.push_back(e);
    }
    g.edges.swap(keep);
    if (g.edges.size() == before) KC_FAIL(KC_ERR_SLOT_NOT_OCCUPIED, "slot %u of node %u is not in use", slot_id, node_id);
    return KC_OK;
}

int32_t kcg_connect(kc_graph& g, uint32_t out_id, uint32_t in_id, uint32_t out_slot, uint32_t in_slot) {
    // NodeGraph::connect, :416-446
    const KcNode* on = kcg_find(g, out_id);
    const KcNode* in = kcg_find(g, in_id);
    if (!on) KC_FAIL(KC_ERR_INVALID_NODE_ID, "no node %u", out_id);
    if (!in) KC_FAIL(KC_ERR_INVALID_NODE_ID, "no node %u", in_id);
    int ot = 0, it = 0;
    KC_TRY(slot_type_lookup(kcg_output_slots(*on), out_slot, &ot));
    KC_TRY(slot_type_lookup(kcg_input_slots(*in), in_slot, &it));
    if (!slot_fits(ot, it)) KC_FAIL(KC_ERR_INVALID_SLOT_TYPE, "slot types do not fit");
    kcg_disconnect_slot(g, in_id, KC_SIDE_INPUT, in_slot, nullptr);  // result deliberately ignored, :435
    for (const kc_edge& e : g.edges)
        if (e.output_id == out_id && e.input_id == in_id && e.output_slot == out_slot && e.input_slot == in_slot)
            KC_FAIL(KC_ERR_INVALID_EDGE, "edge already exists");
    g.edges.push_back(kc_edge{out_id, in_id, out_slot, in_slot});
    return KC_OK;
}

int32_t kcg_remove_node(kc_graph& g, uint32_t node_id, std::vector<kc_edge>* removed) {
    // NodeGraph::remove_node + disconnect_node, :476-498
    if (!kcg_find(g, node_id)) KC_FAIL(KC_ERR_INVALID_NODE_ID, "no node %u", node_id);
    std::vector<kc_edge> keep;
    for (const kc_edge& e : g.edges) {
        if (e.output_id == node_id || e.input_id == node_id) { if (removed) removed->push_back(e); }
        else keep.push_back(e);
    }
    g.edges.swap(keep);
    g.nodes.erase(std::remove_if(g.nodes.begin(), g.nodes.end(), [&](const KcNode& n) { return n.node_id == node_id; }), g.nodes.end());
    return KC_OK;
}

// ---------------------------------------------------------------------------
// JSON (serde_json externally-tagged enums, as in data/invert_graph.json)
// ---------------------------------------------------------------------------
namespace {

struct JVal {
    enum T { NUL, BOOL, NUM, STR, ARR, OBJ } t = NUL;
    double num = 0;
    bool is_int = false;   // the token had no fraction and no exponent (serde takes u32 fields only from those)
    bool b = false;
    std::string str;
    std::vector<JVal> arr;
    std::vector<std::pair<std::string, JVal>> obj;
    const JVal* get(const char* k) const {
        for (auto& kv : obj)
            if (kv.first == k) return &kv.second;
        return nullptr;
    }
};

struct JParser {
    const char* p;
    const char* end;
    bool ok = true;
    int depth = 0;         // serde_json refuses documents nested deeper than 128
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    bool lit(const char* s) {
        size_t n = strlen(s);
        if ((size_t)(end - p) >= n && memcmp(p, s, n) == 0) { p += n; return true; }
        return false;
    }
    JVal parse() {
        JVal v;
        if (++depth > 128) ok = false;
        if (ok) parse_into(v);
        --depth;
        return v;
    }
    // -?(0|[1-9][0-9]*)(.[0-9]+)?([eE][+-]?[0-9]+)? and nothing strtod would take beyond that (inf, nan, hex, +1, .5)
    bool number(JVal& v) {
        const char* q = p;
        if (q < end && *q == '-') ++q;
        if (q >= end || *q < '0' || *q > '9') return false;
        if (*q == '0') ++q;
        else while (q < end && *q >= '0' && *q <= '9') ++q;
        bool integral = true;
        if (q < end && *q == '.') {
            integral = false;
            ++q;
            if (q >= end || *q < '0' || *q > '9') return false;
            while (q < end && *q >= '0' && *q <= '9') ++q;
        }
        if (q < end && (*q == 'e' || *q == 'E')) {
            integral = false;
            ++q;
            if (q < end && (*q == '+' || *q == '-')) ++q;
            if (q >= end || *q < '0' || *q > '9') return false;
            while (q < end && *q >= '0' && *q <= '9') ++q;
        }
        const std::string tok(p, q);
        v.t = JVal::NUM;
        v.num = strtod(tok.c_str(), nullptr);
        v.is_int = integral;
        if (!std::isfinite(v.num)) return false;   // serde_json: "number out of range"
        p = q;
        return true;
    }
    void parse_into(JVal& v) {
        ws();
        if (p >= end) { ok = false; return; }
        if (*p == '{') {
            v.t = JVal::OBJ;
            ++p; ws();
            if (p < end && *p == '}') { ++p; return; }
            while (ok) {
                ws();
                JVal k = parse_string();
                ws();
                if (!ok || p >= end || *p != ':') { ok = false; break; }
                ++p;
                JVal val = parse();
                v.obj.emplace_back(k.str, std::move(val));
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; break; }
                ok = false;
            }
        } else if (*p == '[') {
            v.t = JVal::ARR;
            ++p; ws();
            if (p < end && *p == ']') { ++p; return; }
            while (ok) {
                v.arr.push_back(parse());
                ws();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; break; }
                ok = false;
            }
        } else if (*p == '"') {
            v = parse_string();
        } else if (lit("true")) { v.t = JVal::BOOL; v.b = true; }
        else if (lit("false")) { v.t = JVal::BOOL; v.b = false; }
        else if (lit("null")) { v.t = JVal::NUL; }
        else if (!number(v)) ok = false;
    }
    JVal parse_string() {
        JVal v;
        v.t = JVal::STR;
        if (p >= end || *p != '"') { ok = false; return v; }
        ++p;
        while (p < end && *p != '"') {
            if (*p == '\\' && p + 1 < end) {
                ++p;
                switch (*p) {
                    case 'n': v.str += '\n'; break;
                    case 't': v.str += '\t'; break;
                    case 'r': v.str += '\r'; break;
                    case 'b': v.str += '\b'; break;
                    case 'f': v.str += '\f'; break;
                    case 'u': {
                        if (end - p < 5) { ok = false; return v; }
                        unsigned cp = (unsigned)strtoul(std::string(p + 1, p + 5).c_str(), nullptr, 16);
                        p += 4;
                        if (cp < 0x80) v.str += (char)cp;
                        else if (cp < 0x800) { v.str += (char)(0xc0 | (cp >> 6)); v.str += (char)(0x80 | (cp & 0x3f)); }
                        else { v.str += (char)(0xe0 | (cp >> 12)); v.str += (char)(0x80 | ((cp >> 6) & 0x3f)); v.str += (char)(0x80 | (cp & 0x3f)); }
                        break;
                    }
                    default: v.str += *p; break;  // \" \\ \/
                }
                ++p;
            } else {
                v.str += *p++;
            }
        }
        if (p >= end) { ok = false; return v; }
        ++p;
        return v;
    }
};

const char* kNodeTypeNames[] = {"InputGray", "InputRgba", "OutputGray", "OutputRgba", "Graph", "Image", "Embed",
                                "Write", "Value", "Mix", "HeightToNormal", "SeparateRgba", "CombineRgba"};
const char* kMixNames[] = {"Add", "Subtract", "Multiply", "Divide", "Pow"};
const char* kPolicyNames[] = {"MostPixels", "LeastPixels", "LargestAxes", "SmallestAxes", "SpecificSlot", "SpecificSize"};
const char* kFilterNames[] = {"Nearest", "Triangle", "CatmullRom", "Gaussian", "Lanczos3"};

int index_of(const char* const* names, int n, const std::string& s) {
    for (int i = 0; i < n; ++i)
        if (s == names[i]) return i;
    return -1;
}

int32_t graph_from_jval(const JVal& root, kc_graph& g);

// a u32 field: serde takes it only from a non-negative integer token that fits
bool to_u32(const JVal* j, uint32_t& out) {
    if (!j || j->t != JVal::NUM || !j->is_int || j->num < 0.0 || j->num > 4294967295.0) return false;
    out = (uint32_t)j->num;
    return true;
}

int32_t node_from_jval(const JVal& j, KcNode& n) {
    if (j.t != JVal::OBJ) KC_FAIL(KC_ERR_IO, "json: node is not an object");
    const JVal* id = j.get("node_id");
    const JVal* ty = j.get("node_type");
    const JVal* pol = j.get("resize_policy");
    const JVal* fil = j.get("resize_filter");
    if (!id || !ty || !pol || !fil) KC_FAIL(KC_ERR_IO, "json: node is missing a field");
    if (!to_u32(id, n.node_id)) KC_FAIL(KC_ERR_IO, "json: node_id is not a u32");
    // node_type: "Variant" for unit variants, {"Variant": payload} otherwise
    std::string variant;
    const JVal* payload = nullptr;
    if (ty->t == JVal::STR) variant = ty->str;
    else if (ty->t == JVal::OBJ && ty->obj.size() == 1) { variant = ty->obj[0].first; payload = &ty->obj[0].second; }
    else KC_FAIL(KC_ERR_IO, "json: bad node_type");
    n.type = index_of(kNodeTypeNames, 13, variant);
    if (n.type < 0) KC_FAIL(KC_ERR_IO, "json: unknown node type '%s'", variant.c_str());
    switch (n.type) {
        case KC_NODE_INPUT_GRAY: case KC_NODE_INPUT_RGBA: case KC_NODE_OUTPUT_GRAY: case KC_NODE_OUTPUT_RGBA:
        case KC_NODE_IMAGE: case KC_NODE_WRITE:
            if (!payload || payload->t != JVal::STR) KC_FAIL(KC_ERR_IO, "json: %s needs a string", variant.c_str());
            n.name = payload->str;
            break;
        case KC_NODE_GRAPH: {
            if (!payload) KC_FAIL(KC_ERR_IO, "json: Graph needs a NodeGraph");
            n.graph = std::make_shared<kc_graph>();
            KC_TRY(graph_from_jval(*payload, *n.graph));
            break;
        }
        case KC_NODE_EMBED:
            if (!to_u32(payload, n.embed_id)) KC_FAIL(KC_ERR_IO, "json: Embed needs an id (u32)");
            break;
        case KC_NODE_VALUE:
            if (!payload || payload->t != JVal::NUM) KC_FAIL(KC_ERR_IO, "json: Value needs a number");
            n.value = (float)payload->num;
            break;
        case KC_NODE_MIX:
            if (!payload || payload->t != JVal::STR) KC_FAIL(KC_ERR_IO, "json: Mix needs a MixType");
            n.mix_type = index_of(kMixNames, 5, payload->str);
            if (n.mix_type < 0) KC_FAIL(KC_ERR_IO, "json: unknown MixType '%s'", payload->str.c_str());
            break;
        default: break;
    }
    if (pol->t == JVal::STR) {
        n.policy = index_of(kPolicyNames, 4, pol->str);
        if (n.policy < 0) KC_FAIL(KC_ERR_IO, "json: unknown resize_policy '%s'", pol->str.c_str());
    } else if (pol->t == JVal::OBJ && pol->obj.size() == 1) {
        const std::string& k = pol->obj[0].first;
        const JVal& v = pol->obj[0].second;
        if (k == "SpecificSlot" && to_u32(&v, n.policy_slot)) {
            n.policy = KC_POLICY_SPECIFIC_SLOT;
        } else if (k == "SpecificSize" && v.t == JVal::OBJ && to_u32(v.get("width"), n.policy_w) && to_u32(v.get("height"), n.policy_h)) {
            n.policy = KC_POLICY_SPECIFIC_SIZE;
        } else KC_FAIL(KC_ERR_IO, "json: bad resize_policy");
    } else KC_FAIL(KC_ERR_IO, "json: bad resize_policy");
    if (fil->t != JVal::STR) KC_FAIL(KC_ERR_IO, "json: bad resize_filter");
    n.filter = index_of(kFilterNames, 5, fil->str);
    if (n.filter < 0) KC_FAIL(KC_ERR_IO, "json: unknown resize_filter '%s'", fil->str.c_str());
    return KC_OK;
}

int32_t graph_from_jval(const JVal& root, kc_graph& g) {
    if (root.t != JVal::OBJ) KC_FAIL(KC_ERR_IO, "json: graph is not an object");
    const JVal* nodes = root.get("nodes");
    const JVal* edges = root.get("edges");
    if (!nodes || nodes->t != JVal::ARR || !edges || edges->t != JVal::ARR) KC_FAIL(KC_ERR_IO, "json: graph needs nodes[] and edges[]");
    g.nodes.clear();
    g.edges.clear();
    for (const JVal& j : nodes->arr) {
        KcNode n;
        KC_TRY(node_from_jval(j, n));
        g.nodes.push_back(std::move(n));
    }
    for (const JVal& j : edges->arr) {
        kc_edge e{};
        if (j.t != JVal::OBJ || !to_u32(j.get("output_id"), e.output_id) || !to_u32(j.get("input_id"), e.input_id) ||
            !to_u32(j.get("output_slot"), e.output_slot) || !to_u32(j.get("input_slot"), e.input_slot))
            KC_FAIL(KC_ERR_IO, "json: an edge needs output_id, input_id, output_slot, input_slot (u32)");
        g.edges.push_back(e);
    }
    // NodeGraph::from_path, :36-43: the id counter restarts after the largest id
    uint32_t mx = 0;
    bool any = false;
    for (const KcNode& n : g.nodes) { mx = std::max(mx, n.node_id); any = true; }
    g.node_id_counter = any ? mx + 1 : 0;
    return KC_OK;
}

std::string json_escape(const std::string& s) {
    std::string o = "\"";
    for (unsigned char ch : s) {
        switch (ch) {
            case '"': o += "\\\""; break;
            case '\\': o += "\\\\"; break;
            case '\n': o += "\\n"; break;
            case '\t': o += "\\t"; break;
            case '\r': o += "\\r"; break;
            default:
                if (ch < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", ch); o += b; }
                else o += (char)ch;
        }
    }
    return o + "\"";
}

std::string f32_to_json(float v) {  // shortest text that round-trips, like serde_json/ryu
    if (!std::isfinite(v)) return "null";
    char buf[64];
    for (int prec = 1; prec <= 9; ++prec) {
        snprintf(buf, sizeof buf, "%.*g", prec, (double)v);
        if (strtof(buf, nullptr) == v) break;
    }
    std::string s = buf;
    if (s.find('e') != std::string::npos) {
        // ryu prints plain decimals for moderate exponents; fall back to fixed notation
        snprintf(buf, sizeof buf, "%.9g", (double)v);
        std::ostringstream os;
        os.precision(9);
        os << std::fixed << (double)v;
        if (strtof(os.str().c_str(), nullptr) == v && std::fabs(v) < 1e16f && std::fabs(v) >= 1e-5f) {
            s = os.str();
            while (s.size() > 1 && s.back() == '0' && s[s.size() - 2] != '.') s.pop_back();
        }
    }
    if (s.find('.') == std::string::npos && s.find('e') == std::string::npos) s += ".0";
    return s;
}

void indent(std::string& o, int n) { o.append((size_t)n * 2, ' '); }

void graph_to_json(const kc_graph& g, std::string& o, int lvl);

void node_to_json(const KcNode& n, std::string& o, int lvl) {
    indent(o, lvl); o += "{\n";
    indent(o, lvl + 1); o += "\"node_id\": " + std::to_string(n.node_id) + ",\n";
    indent(o, lvl + 1); o += "\"node_type\": ";
    const std::string variant = kNodeTypeNames[n.type];
    switch (n.type) {
        case KC_NODE_HEIGHT_TO_NORMAL: case KC_NODE_SEPARATE_RGBA: case KC_NODE_COMBINE_RGBA:
            o += "\"" + variant + "\"";
            break;
        default: {
            o += "{\n";
            indent(o, lvl + 2); o += "\"" + variant + "\": ";
            if (n.type == KC_NODE_GRAPH) graph_to_json(*n.graph, o, lvl + 2);
            else if (n.type == KC_NODE_EMBED) o += std::to_string(n.embed_id);
            else if (n.type == KC_NODE_VALUE) o += f32_to_json(n.value);
            else if (n.type == KC_NODE_MIX) o += std::string("\"") + kMixNames[n.mix_type] + "\"";
            else o += json_escape(n.name);
            o += "\n";
            indent(o, lvl + 1); o += "}";
        }
    }
    o += ",\n";
    indent(o, lvl + 1); o += "\"resize_policy\": ";
    if (n.policy == KC_POLICY_SPECIFIC_SLOT) {
        o += "{\n"; indent(o, lvl + 2); o += "\"SpecificSlot\": " + std::to_string(n.policy_slot) + "\n"; indent(o, lvl + 1); o += "}";
    } else if (n.policy == KC_POLICY_SPECIFIC_SIZE) {
        o += "{\n"; indent(o, lvl + 2); o += "\"SpecificSize\": {\n";
        indent(o, lvl + 3); o += "\"width\": " + std::to_string(n.policy_w) + ",\n";
        indent(o, lvl + 3); o += "\"height\": " + std::to_string(n.policy_h) + "\n";
        indent(o, lvl + 2); o += "}\n"; indent(o, lvl + 1); o += "}";
    } else {
        o += std::string("\"") + kPolicyNames[n.policy] + "\"";
    }
    o += ",\n";
    indent(o, lvl + 1); o += std::string("\"resize_filter\": \"") + kFilterNames[n.filter] + "\"\n";
    indent(o, lvl); o += "}";
}

void graph_to_json(const kc_graph& g, std::string& o, int lvl) {
    o += "{\n";
    indent(o, lvl + 1); o += "\"nodes\": [";
    for (size_t i = 0; i < g.nodes.size(); ++i) {
        o += i ? ",\n" : "\n";
        node_to_json(g.nodes[i], o, lvl + 2);
    }
    if (!g.nodes.empty()) { o += "\n"; indent(o, lvl + 1); }
    o += "],\n";
    indent(o, lvl + 1); o += "\"edges\": [";
    for (size_t i = 0; i < g.edges.size(); ++i) {
        const kc_edge& e = g.edges[i];
        o += i ? ",\n" : "\n";
        indent(o, lvl + 2); o += "{\n";
        indent(o, lvl + 3); o += "\"output_id\": " + std::to_string(e.output_id) + ",\n";
        indent(o, lvl + 3); o += "\"input_id\": " + std::to_string(e.input_id) + ",\n";
        indent(o, lvl + 3); o += "\"output_slot\": " + std::to_string(e.output_slot) + ",\n";
        indent(o, lvl + 3); o += "\"input_slot\": " + std::to_string(e.input_slot) + "\n";
        indent(o, lvl + 2); o += "}";
    }
    if (!g.edges.empty()) { o += "\n"; indent(o, lvl + 1); }
    o += "]\n";
    indent(o, lvl); o += "}";
}

}  // namespace

int32_t kcg_parse_json(const std::string& text, kc_graph& out) {
    JParser ps{text.data(), text.data() + text.size()};
    JVal root = ps.parse();
    ps.ws();
    if (!ps.ok || ps.p != ps.end) KC_FAIL(KC_ERR_IO, "json: syntax error at byte %zu", (size_t)(ps.p - text.data()));
    return graph_from_jval(root, out);
}

std::string kcg_to_json(const kc_graph& g) {
    std::string o;
    graph_to_json(g, o, 0);
    return o;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int32_t kc_graph_create(kc_graph** out) try {
    if (!out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "out is NULL");
    *out = new kc_graph();
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_destroy(kc_graph* g) try {
    delete g;
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_clone(const kc_graph* g, kc_graph** out) try {
    if (!g || !out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = new kc_graph(*g);
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_from_json(const char* text, kc_graph** out) try {
    if (!text || !out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    auto g = std::make_unique<kc_graph>();
    KC_TRY(kcg_parse_json(text, *g));
    *out = g.release();
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_from_path(const char* path, kc_graph** out) try {
    if (!path || !out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    std::ifstream f(path, std::ios::binary);
    if (!f) KC_FAIL(KC_ERR_IO, "cannot open '%s'", path);
    std::stringstream ss;
    ss << f.rdbuf();
    return kc_graph_from_json(ss.str().c_str(), out);
} KC_ABI_CATCH
int32_t kc_graph_export_json(const kc_graph* g, char** out_text) try {
    if (!g || !out_text) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    std::string s = kcg_to_json(*g);
    *out_text = (char*)malloc(s.size() + 1);
    if (!*out_text) KC_FAIL(KC_ERR_GENERIC, "out of memory");
    memcpy(*out_text, s.c_str(), s.size() + 1);
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_export_json_path(const kc_graph* g, const char* path) try {
    if (!g || !path) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    std::ofstream f(path, std::ios::binary);
    if (!f) KC_FAIL(KC_ERR_IO, "cannot create '%s'", path);
    f << kcg_to_json(*g);
    return f.good() ? KC_OK : KC_ERR_IO;
} KC_ABI_CATCH
int32_t kc_graph_add_node(kc_graph* g, const kc_node_desc* node, uint32_t* out_node_id) try {
    if (!g || !node) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcNode n;
    kcg_from_desc(*node, n);
    return kcg_add_node(*g, std::move(n), out_node_id);
} KC_ABI_CATCH
int32_t kc_graph_add_node_with_id(kc_graph* g, const kc_node_desc* node) try {
    if (!g || !node) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcNode n;
    kcg_from_desc(*node, n);
    return kcg_add_node_with_id(*g, std::move(n));
} KC_ABI_CATCH
int32_t kc_graph_remove_node(kc_graph* g, uint32_t node_id) try {
    if (!g) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    return kcg_remove_node(*g, node_id, nullptr);
} KC_ABI_CATCH
int32_t kc_graph_connect(kc_graph* g, uint32_t o, uint32_t i, uint32_t os, uint32_t is) try {
    if (!g) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    return kcg_connect(*g, o, i, os, is);
} KC_ABI_CATCH
int32_t kc_graph_try_connect(kc_graph* g, uint32_t o, uint32_t i, uint32_t os, uint32_t is) try {
    // NodeGraph::try_connect + can_connect, :376-413 (no slot-type check there)
    if (!g) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    const KcNode* on = kcg_find(*g, o);
    const KcNode* in = kcg_find(*g, i);
    if (!on || !in) KC_FAIL(KC_ERR_INVALID_NODE_ID, "no such node");
    int t = 0;
    KC_TRY(slot_type_lookup(kcg_output_slots(*on), os, &t));
    KC_TRY(slot_type_lookup(kcg_input_slots(*in), is, &t));
    for (const kc_edge& e : g->edges)
        if (e.input_id == i && e.input_slot == is) KC_FAIL(KC_ERR_SLOT_OCCUPIED, "slot %u of node %u is occupied", is, i);
    g->edges.push_back(kc_edge{o, i, os, is});
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_can_connect(const kc_graph* g, uint32_t o, uint32_t i, uint32_t os, uint32_t is) try {
    // NodeGraph::can_connect, src/node_graph.rs:376-393
    if (!g) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    const KcNode* on = kcg_find(*g, o);
    const KcNode* in = kcg_find(*g, i);
    if (!on || !in) KC_FAIL(KC_ERR_INVALID_NODE_ID, "no such node");
    int t = 0;
    KC_TRY(slot_type_lookup(kcg_output_slots(*on), os, &t));
    KC_TRY(slot_type_lookup(kcg_input_slots(*in), is, &t));
    for (const kc_edge& e : g->edges)
        if (e.input_id == i && e.input_slot == is) KC_FAIL(KC_ERR_SLOT_OCCUPIED, "slot %u of node %u is occupied", is, i);
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_connected_edges(const kc_graph* g, uint32_t node_id, int32_t side, uint32_t slot_id, kc_edge* edges, size_t cap, size_t* n) try {
    // NodeGraph::connected_edges, src/node_graph.rs:518-537 (side 0 = Input, 1 = Output)
    if (!g || !n) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    if (!kcg_find(*g, node_id)) KC_FAIL(KC_ERR_INVALID_NODE_ID, "no node %u", node_id);
    size_t k = 0;
    for (const kc_edge& e : g->edges) {
        const bool hit = side == 0 ? (e.input_id == node_id && e.input_slot == slot_id) : (e.output_id == node_id && e.output_slot == slot_id);
        if (!hit) continue;
        if (edges && k < cap) edges[k] = e;
        ++k;
    }
    *n = k;
    if (k == 0) KC_FAIL(KC_ERR_SLOT_NOT_OCCUPIED, "slot %u of node %u has no edges", slot_id, node_id);
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_new_id(kc_graph* g, uint32_t* out) try {
    // NodeGraph::new_id, src/node_graph.rs:86-96
    if (!g || !out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = new_id(*g);
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_rename_output_node(kc_graph* g, uint32_t node_id, const char* new_name, char** old_name) try {
    // NodeGraph::rename_output_node, src/node_graph.rs:232-270: the new name is de-collided
    // against the OTHER output names; returns the old name (kc_free it)
    if (!g || !new_name) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcNode* node = kcg_find(*g, node_id);
    if (!node) KC_FAIL(KC_ERR_INVALID_NODE_ID, "no node %u", node_id);
    if (!kcg_is_output(node->type)) KC_FAIL(KC_ERR_INVALID_NODE_TYPE, "node %u is not an output node", node_id);
    std::vector<std::string> names;
    bool skipped = false;
    for (const KcNode& m : g->nodes) {
        if (!kcg_is_output(m.type)) continue;
        if (!skipped && m.name == node->name) { skipped = true; continue; }   // remove the first occurrence of the old name
        names.push_back(m.name);
    }
    const std::string old = node->name;
    node->name = avoid_name_collision(names, new_name);
    if (old_name) {
        *old_name = (char*)malloc(old.size() + 1);
        if (!*old_name) KC_FAIL(KC_ERR_GENERIC, "out of memory");
        memcpy(*old_name, old.c_str(), old.size() + 1);
    }
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_disconnect_slot(kc_graph* g, uint32_t node_id, int32_t side, uint32_t slot_id) try {
    if (!g) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    return kcg_disconnect_slot(*g, node_id, side, slot_id, nullptr);
} KC_ABI_CATCH
int32_t kc_graph_remove_edge(kc_graph* g, const kc_edge* e) try {
    if (!g || !e) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    for (size_t i = 0; i < g->edges.size(); ++i) {
        const kc_edge& c = g->edges[i];
        if (c.output_id == e->output_id && c.input_id == e->input_id && c.output_slot == e->output_slot && c.input_slot == e->input_slot) {
            g->edges.erase(g->edges.begin() + (long)i);
            return KC_OK;
        }
    }
    KC_FAIL(KC_ERR_INVALID_EDGE, "no such edge");
} KC_ABI_CATCH
int32_t kc_graph_node_count(const kc_graph* g, size_t* n) try {
    if (!g || !n) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    *n = g->nodes.size();
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_node_at(const kc_graph* g, size_t index, kc_node_desc* out) try {
    if (!g || !out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    if (index >= g->nodes.size()) KC_FAIL(KC_ERR_INVALID_NODE_ID, "node index out of range");
    kcg_to_desc(g->nodes[index], *out);
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_node(const kc_graph* g, uint32_t node_id, kc_node_desc* out) try {
    if (!g || !out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    const KcNode* n = kcg_find(*g, node_id);
    if (!n) KC_FAIL(KC_ERR_INVALID_NODE_ID, "no node %u", node_id);
    kcg_to_desc(*n, *out);
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_set_node(kc_graph* g, const kc_node_desc* node) try {
    if (!g || !node) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcNode* n = kcg_find(*g, node->node_id);
    if (!n) KC_FAIL(KC_ERR_INVALID_NODE_ID, "no node %u", node->node_id);
    KcNode tmp;
    kcg_from_desc(*node, tmp);
    tmp.priority = n->priority;        // engine state stays with the node (the reference keeps the Arc<Priority>)
    *n = std::move(tmp);
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_set_node_priority(kc_graph* g, uint32_t node_id, int8_t priority) try {
    // Priority::set_priority, src/priority.rs:33-37
    if (!g) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcNode* n = kcg_find(*g, node_id);
    if (!n) KC_FAIL(KC_ERR_INVALID_NODE_ID, "no node %u", node_id);
    n->priority = priority;
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_node_priority(const kc_graph* g, uint32_t node_id, int8_t* priority, int8_t* propagated) try {
    // Priority::priority / propagated_priority after PriorityPropagator::update, src/priority.rs:39-45,101-127
    if (!g) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    const KcNode* n = kcg_find(*g, node_id);
    if (!n) KC_FAIL(KC_ERR_INVALID_NODE_ID, "no node %u", node_id);
    if (priority) *priority = n->priority;
    if (propagated) {
        std::vector<int8_t> prop;
        kcg_propagated_priorities(*g, prop);
        *propagated = prop[(size_t)(n - g->nodes.data())];
    }
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_edge_count(const kc_graph* g, size_t* n) try {
    if (!g || !n) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    *n = g->edges.size();
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_graph_edge_at(const kc_graph* g, size_t index, kc_edge* out) try {
    if (!g || !out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    if (index >= g->edges.size()) KC_FAIL(KC_ERR_INVALID_EDGE, "edge index out of range");
    *out = g->edges[index];
    return KC_OK;
} KC_ABI_CATCH
static int32_t slot_id_with_name(const kc_graph* g, const char* name, bool input, uint32_t* slot_id) {
    if (!g || !name || !slot_id) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    for (const KcNode& n : g->nodes)
        if ((input ? kcg_is_input(n.type) : kcg_is_output(n.type)) && n.name == name) {
            *slot_id = n.node_id;
            return KC_OK;
        }
    KC_FAIL(KC_ERR_INVALID_NAME, "no %s node named '%s'", input ? "input" : "output", name);
}
int32_t kc_graph_input_slot_id_with_name(const kc_graph* g, const char* name, uint32_t* slot_id) try {
    return slot_id_with_name(g, name, true, slot_id);
} KC_ABI_CATCH
int32_t kc_graph_output_slot_id_with_name(const kc_graph* g, const char* name, uint32_t* slot_id) try {
    return slot_id_with_name(g, name, false, slot_id);
} KC_ABI_CATCH
static int32_t ids_of(const kc_graph* g, bool input, uint32_t* ids, size_t cap, size_t* n) {
    if (!g || !n) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    size_t c = 0;
    for (const KcNode& nd : g->nodes)
        if (input ? kcg_is_input(nd.type) : kcg_is_output(nd.type)) {
            if (ids && c < cap) ids[c] = nd.node_id;
            ++c;
        }
    *n = c;
    return KC_OK;
}
int32_t kc_graph_output_ids(const kc_graph* g, uint32_t* ids, size_t cap, size_t* n) { return ids_of(g, false, ids, cap, n); }
int32_t kc_graph_input_ids(const kc_graph* g, uint32_t* ids, size_t cap, size_t* n) { return ids_of(g, true, ids, cap, n); }

static int32_t slots_out(const std::vector<KcSlotInfo>& v, kc_slot* slots, size_t cap, size_t* n) try {
    if (!n) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    *n = v.size();
    for (size_t i = 0; i < v.size() && i < cap && slots; ++i) {
        memset(&slots[i], 0, sizeof(kc_slot));
        strncpy(slots[i].name, v[i].name.c_str(), sizeof(slots[i].name) - 1);
        slots[i].slot_id = v[i].slot_id;
        slots[i].slot_type = v[i].slot_type;
    }
    return KC_OK;
} KC_ABI_CATCH
int32_t kc_node_input_slots(const kc_node_desc* node, kc_slot* slots, size_t cap, size_t* n) try {
    if (!node) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcNode k;
    kcg_from_desc(*node, k);
    return slots_out(kcg_input_slots(k), slots, cap, n);
} KC_ABI_CATCH
int32_t kc_node_output_slots(const kc_node_desc* node, kc_slot* slots, size_t cap, size_t* n) try {
    if (!node) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcNode k;
    kcg_from_desc(*node, k);
    return slots_out(kcg_output_slots(k), slots, cap, n);
} KC_ABI_CATCH

}  // extern "C"
