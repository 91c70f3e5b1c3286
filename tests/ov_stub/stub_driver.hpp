// stub_driver.hpp — a tiny C interface for driving ov::Op classes built against the stand-in OpenVINO API
// (tests/ov_stub/openvino/stub_core.hpp).  TEST INFRASTRUCTURE ONLY.  It plays the part of the OpenVINO runtime:
// "load a layer" (OpExtension::create with the producers and the layer's attributes), then call evaluate() with host
// tensors.  Included by oracle/ref_driver.cpp (the reference's own op classes -> oracle/_ref/libovtok_ref.so) and by
// tests/ov_stub/shim_driver.cpp (this repo's ov::Op shim).
#pragma once
#include <openvino/stub_core.hpp>

#include <chrono>

extern "C" {
struct ovs_tensor {
    int32_t dtype;      // 0 = absent, 1 = i32, 2 = i64, 3 = u8, 4 = boolean, 5 = f32
    int32_t ndim;
    int64_t shape[4];
    void* data;         // node_create: non-null => the input is a Constant holding a copy of these bytes
};
}

namespace ovs {

inline ov::element::Type dtype_of(int32_t d) {
    switch (d) {
    case 1: return ov::element::i32;
    case 2: return ov::element::i64;
    case 3: return ov::element::u8;
    case 4: return ov::element::boolean;
    case 5: return ov::element::f32;
    default: OPENVINO_THROW("stub driver: unknown dtype code ", d);
    }
}
inline int32_t code_of(const ov::element::Type& t) {
    if (t == ov::element::i32) return 1;
    if (t == ov::element::i64) return 2;
    if (t == ov::element::u8) return 3;
    if (t == ov::element::boolean) return 4;
    if (t == ov::element::f32) return 5;
    return 0;
}
inline ov::Shape shape_of(const ovs_tensor& t) {
    ov::Shape s;
    for (int i = 0; i < t.ndim; ++i) s.push_back((size_t)t.shape[i]);
    return s;
}

struct Graph {                                  // a chain of layers "loaded" one after the other, like an IR
    std::vector<std::shared_ptr<ov::Node>> keep;        // producers (Parameters / Constants) and every created node
    std::vector<ov::OutputVector> layer_outputs;        // what OpExtension::create returned per layer
};

struct NodeHandle {
    std::shared_ptr<ov::Node> node;
    std::vector<std::shared_ptr<ov::Node>> producers;
    ov::TensorVector outputs;
    double last_ms = 0;
};

inline thread_local std::string g_error;
using Registry = std::map<std::string, std::shared_ptr<ov::BaseOpExtension>>;

inline std::map<std::string, std::string> parse_attrs(const char* attrs) {     // "key=value\n" lines; values are raw text
    std::map<std::string, std::string> m;
    if (!attrs) return m;
    std::string s(attrs);
    size_t pos = 0;
    while (pos < s.size()) {
        size_t nl = s.find('\x1e', pos);        // record separator: values (regex patterns) may contain newlines
        if (nl == std::string::npos) nl = s.size();
        const std::string line = s.substr(pos, nl - pos);
        const size_t eq = line.find('=');
        if (eq != std::string::npos) m[line.substr(0, eq)] = line.substr(eq + 1);
        pos = nl + 1;
    }
    return m;
}

inline ov::OutputVector make_producers(int n_inputs, const ovs_tensor* protos, std::vector<std::shared_ptr<ov::Node>>& keep) {
    ov::OutputVector in;
    for (int i = 0; i < n_inputs; ++i) {
        const auto t = dtype_of(protos[i].dtype);
        std::shared_ptr<ov::Node> p;
        if (protos[i].data) p = std::make_shared<ov::op::v0::Constant>(t, shape_of(protos[i]), (const void*)protos[i].data);
        else {
            std::vector<ov::Dimension> dims((size_t)protos[i].ndim);   // dynamic dims of the given rank
            p = std::make_shared<ov::op::v0::Parameter>(t, ov::PartialShape(dims));
        }
        keep.push_back(p);
        in.emplace_back(p, 0);
    }
    return in;
}

inline NodeHandle* node_create(const Registry& reg, const char* op_name, int n_inputs, const ovs_tensor* protos, const char* attrs) {
    try {
        auto it = reg.find(op_name);
        OPENVINO_ASSERT(it != reg.end(), "stub driver: no extension registered for op type ", op_name);
        auto h = std::make_unique<NodeHandle>();
        ov::OutputVector in = make_producers(n_inputs, protos, h->producers);
        ov::AttributeVisitor visitor(parse_attrs(attrs));
        ov::OutputVector out = it->second->create(in, visitor);
        OPENVINO_ASSERT(!out.empty(), "stub driver: extension created no outputs");
        h->node = out.at(0).get_node_shared_ptr();
        return h.release();
    } catch (const std::exception& e) {
        g_error = e.what();
        return nullptr;
    }
}

inline int node_evaluate(NodeHandle* h, int n_in, const ovs_tensor* in) {
    try {
        ov::TensorVector inputs;
        for (int i = 0; i < n_in; ++i) {
            if (in[i].data || in[i].ndim == 0) inputs.emplace_back(dtype_of(in[i].dtype), shape_of(in[i]), in[i].data ? in[i].data : (void*)&in[i].shape[3]);
            else inputs.emplace_back(dtype_of(in[i].dtype), shape_of(in[i]));     // empty tensor
        }
        h->outputs.clear();
        for (size_t i = 0; i < h->node->get_output_size(); ++i) h->outputs.emplace_back(h->node->get_output_element_type(i), ov::Shape{0});
        const auto t0 = std::chrono::steady_clock::now();
        const bool ok = h->node->evaluate(h->outputs, inputs);
        h->last_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        OPENVINO_ASSERT(ok, "evaluate() returned false");
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return -1;
    }
}

inline int node_output(NodeHandle* h, int i, ovs_tensor* d) {
    if (i < 0 || (size_t)i >= h->outputs.size()) return -1;
    const ov::Tensor& t = h->outputs[(size_t)i];
    d->dtype = code_of(t.get_element_type());
    const ov::Shape& s = t.get_shape();
    d->ndim = (int32_t)s.size();
    for (size_t k = 0; k < 4; ++k) d->shape[k] = k < s.size() ? (int64_t)s[k] : 0;
    d->data = t.data();
    return 0;
}


// ---- a whole chain of layers, "read" one after the other like an IR, and a minimal executor for it ---------------------------------
struct GraphHandle {
    std::vector<ov::Output<ov::Node>> values;              // value id -> producing output
    std::vector<std::shared_ptr<ov::Node>> params;         // Parameters in creation order
    std::vector<std::shared_ptr<ov::Node>> keep;
    ov::TensorVector results;                              // of the last graph_evaluate
    int evaluated_nodes = 0;
};
inline int graph_add_input(GraphHandle* g, const ovs_tensor* proto) {
    try {
        std::vector<std::shared_ptr<ov::Node>> made;
        ov::OutputVector o = make_producers(1, proto, made);
        g->keep.push_back(made[0]);
        if (!proto->data) g->params.push_back(made[0]);
        g->values.push_back(o[0]);
        return (int)g->values.size() - 1;
    } catch (const std::exception& e) { g_error = e.what(); return -1; }
}
inline int graph_add_layer(const Registry& reg, GraphHandle* g, const char* op, int n_in, const int* ids, const char* attrs, int* out_ids, int max_out) {
    try {
        auto it = reg.find(op);
        OPENVINO_ASSERT(it != reg.end(), "stub driver: no extension registered for op type ", op);
        ov::OutputVector in;
        for (int i = 0; i < n_in; ++i) in.push_back(g->values.at((size_t)ids[i]));
        ov::AttributeVisitor visitor(parse_attrs(attrs));
        ov::OutputVector out = it->second->create(in, visitor);
        OPENVINO_ASSERT(!out.empty() && (int)out.size() <= max_out, "stub driver: unexpected number of outputs");
        g->keep.push_back(out[0].get_node_shared_ptr());
        for (size_t i = 0; i < out.size(); ++i) { g->values.push_back(out[i]); out_ids[i] = (int)g->values.size() - 1; }
        return (int)out.size();
    } catch (const std::exception& e) { g_error = e.what(); return -1; }
}
inline void graph_eval_node(GraphHandle* g, ov::Node* n, std::map<ov::Node*, ov::TensorVector>& done, const ov::TensorVector& param_values) {
    if (done.count(n)) return;
    for (size_t p = 0; p < g->params.size(); ++p)
        if (g->params[p].get() == n) { done[n] = {param_values.at(p)}; return; }
    ov::TensorVector in;
    for (size_t i = 0; i < n->get_input_size(); ++i) {
        ov::Node* src = n->input_value(i).get_node();
        graph_eval_node(g, src, done, param_values);
        in.push_back(done[src].at(n->input_value(i).get_index()));
    }
    ov::TensorVector out;
    for (size_t i = 0; i < n->get_output_size(); ++i) out.emplace_back(n->get_output_element_type(i), ov::Shape{0});
    OPENVINO_ASSERT(n->evaluate(out, in), n->get_type_name(), ": evaluate() returned false");
    if (std::strcmp(n->get_type_name(), "Constant") != 0) ++g->evaluated_nodes;          // layers with an evaluate() of their own
    done[n] = out;
}
inline int graph_evaluate(GraphHandle* g, int n_want, const int* want, int n_params, const ovs_tensor* params) {
    try {
        OPENVINO_ASSERT((size_t)n_params == g->params.size(), "stub driver: expected ", g->params.size(), " parameter tensors");
        ov::TensorVector pv;
        for (int i = 0; i < n_params; ++i) {
            if (params[i].data || params[i].ndim == 0) pv.emplace_back(dtype_of(params[i].dtype), shape_of(params[i]), params[i].data ? params[i].data : (void*)&params[i].shape[3]);
            else pv.emplace_back(dtype_of(params[i].dtype), shape_of(params[i]));
        }
        std::map<ov::Node*, ov::TensorVector> done;
        g->results.clear();
        g->evaluated_nodes = 0;
        for (int i = 0; i < n_want; ++i) {
            const auto& v = g->values.at((size_t)want[i]);
            graph_eval_node(g, v.get_node(), done, pv);
            g->results.push_back(done[v.get_node()].at(v.get_index()));
        }
        return 0;
    } catch (const std::exception& e) { g_error = e.what(); return -1; }
}

}  // namespace ovs

// The extern "C" surface; REGISTRY is an expression yielding `const ovs::Registry&`.
#define OVS_DEFINE_C_API(PREFIX, REGISTRY)                                                                                          \
    extern "C" __attribute__((visibility("default"))) void* PREFIX##_node_create(const char* op, int n, const ovs_tensor* protos, const char* attrs) { \
        return ovs::node_create(REGISTRY, op, n, protos, attrs);                                                                     \
    }                                                                                                                                \
    extern "C" __attribute__((visibility("default"))) int PREFIX##_node_evaluate(void* h, int n, const ovs_tensor* in) {             \
        return ovs::node_evaluate(static_cast<ovs::NodeHandle*>(h), n, in);                                                          \
    }                                                                                                                                \
    extern "C" __attribute__((visibility("default"))) int PREFIX##_node_n_outputs(void* h) {                                         \
        return (int)static_cast<ovs::NodeHandle*>(h)->outputs.size();                                                                \
    }                                                                                                                                \
    extern "C" __attribute__((visibility("default"))) int PREFIX##_node_output(void* h, int i, ovs_tensor* d) {                      \
        return ovs::node_output(static_cast<ovs::NodeHandle*>(h), i, d);                                                             \
    }                                                                                                                                \
    /* a second handle on the SAME op instance: evaluate() is const and may run concurrently from several infer requests */        \
    extern "C" __attribute__((visibility("default"))) void* PREFIX##_node_share(void* h) {                                           \
        auto* n = new ovs::NodeHandle();                                                                                             \
        n->node = static_cast<ovs::NodeHandle*>(h)->node;                                                                            \
        n->producers = static_cast<ovs::NodeHandle*>(h)->producers;                                                                  \
        return n;                                                                                                                    \
    }                                                                                                                                \
    extern "C" __attribute__((visibility("default"))) double PREFIX##_node_last_ms(void* h) { return static_cast<ovs::NodeHandle*>(h)->last_ms; } \
    extern "C" __attribute__((visibility("default"))) const char* PREFIX##_node_type(void* h) { return static_cast<ovs::NodeHandle*>(h)->node->get_type_name(); } \
    extern "C" __attribute__((visibility("default"))) void PREFIX##_node_destroy(void* h) { delete static_cast<ovs::NodeHandle*>(h); } \
    extern "C" __attribute__((visibility("default"))) void* PREFIX##_graph_create(void) { return new ovs::GraphHandle(); }                \
    extern "C" __attribute__((visibility("default"))) void PREFIX##_graph_destroy(void* g) { delete static_cast<ovs::GraphHandle*>(g); }    \
    extern "C" __attribute__((visibility("default"))) int PREFIX##_graph_add_input(void* g, const ovs_tensor* proto) {                       \
        return ovs::graph_add_input(static_cast<ovs::GraphHandle*>(g), proto);                                                              \
    }                                                                                                                                        \
    extern "C" __attribute__((visibility("default"))) int PREFIX##_graph_add_layer(void* g, const char* op, int n, const int* ids, const char* attrs, int* out_ids, int max_out) { \
        return ovs::graph_add_layer(REGISTRY, static_cast<ovs::GraphHandle*>(g), op, n, ids, attrs, out_ids, max_out);                       \
    }                                                                                                                                        \
    extern "C" __attribute__((visibility("default"))) const char* PREFIX##_graph_value_type(void* g, int id) {                               \
        return static_cast<ovs::GraphHandle*>(g)->values.at((size_t)id).get_node()->get_type_name();                                         \
    }                                                                                                                                        \
    extern "C" __attribute__((visibility("default"))) int PREFIX##_graph_evaluate(void* g, int n_want, const int* want, int n_params, const ovs_tensor* params) { \
        return ovs::graph_evaluate(static_cast<ovs::GraphHandle*>(g), n_want, want, n_params, params);                                       \
    }                                                                                                                                        \
    extern "C" __attribute__((visibility("default"))) int PREFIX##_graph_evaluated_nodes(void* g) { return static_cast<ovs::GraphHandle*>(g)->evaluated_nodes; } \
    extern "C" __attribute__((visibility("default"))) int PREFIX##_graph_result(void* g, int i, ovs_tensor* d) {                             \
        auto* G = static_cast<ovs::GraphHandle*>(g);                                                                                         \
        if (i < 0 || (size_t)i >= G->results.size()) return -1;                                                                              \
        ovs::NodeHandle tmp; tmp.outputs = {G->results[(size_t)i]};                                                                          \
        return ovs::node_output(&tmp, 0, d);                                                                                                 \
    }                                                                                                                                        \
    extern "C" __attribute__((visibility("default"))) const char* PREFIX##_last_error(void) { return ovs::g_error.c_str(); }
