// stub_core.hpp — a MINIMAL stand-in for the OpenVINO C++ API, TEST INFRASTRUCTURE ONLY.
//
// OpenVINO is not installed in the build container.  This header tree implements just enough of the
// public API — ov::Tensor, ov::Shape / PartialShape, ov::element::Type, ov::Node / ov::op::Op / ov::Output,
// ov::AttributeVisitor, ov::op::v0::Constant / Parameter, ov::OpExtension, OPENVINO_ASSERT … — with OpenVINO's documented
// semantics so that
//   (1) the reference's own op sources compile UNMODIFIED, from where they lie under /root/reference/src, into
//       oracle/_ref/libovtok_ref.so (recipe: oracle/Makefile target `_ref`) and their evaluate() can be called, and
//   (2) this repo's ov::Op shim (openvino_tokenizers_b200/csrc/ov_shim/) compiles in the CPU tier and its evaluate()
//       / IR-load hooks can be driven by tests.
// It is original code written against the public API's documented behaviour; nothing here comes from OpenVINO's sources.
// It is never linked into the product library (libb200tok.so) and never shipped as an OpenVINO replacement.
#pragma once
#include <algorithm>
#include <any>
#include <cstdint>
#include <cstring>
#include <deque>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <typeinfo>
#include <unordered_map>
#include <utility>
#include <vector>

#define OPENVINO_STUB 1

namespace ov {

// ------------------------------------------------------------------ errors
class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string& m) : std::runtime_error(m) {}
};
class AssertFailure : public Exception {
public:
    using Exception::Exception;
};
namespace stub {
template <class... A>
std::string cat(A&&... a) {
    std::ostringstream s;
    (void)std::initializer_list<int>{((s << a), 0)...};
    return s.str();
}
}  // namespace stub
#define OPENVINO_THROW(...) throw ::ov::Exception(::ov::stub::cat(__VA_ARGS__))
#define OPENVINO_ASSERT(cond, ...)                                                                                 \
    do {                                                                                                           \
        if (!(cond)) throw ::ov::AssertFailure(::ov::stub::cat("Check '" #cond "' failed at ", __FILE__, ":", __LINE__, " ", ##__VA_ARGS__)); \
    } while (0)
#define FRONT_END_GENERAL_CHECK(cond, ...) OPENVINO_ASSERT(cond, ##__VA_ARGS__)
#define OPENVINO_NOT_IMPLEMENTED OPENVINO_THROW("not implemented")

// ------------------------------------------------------------------ element types
namespace element {
enum class Type_t { dynamic, boolean, bf16, f16, f32, f64, i4, i8, i16, i32, i64, u1, u4, u8, u16, u32, u64, string };
class Type {
public:
    constexpr Type() = default;
    constexpr Type(Type_t t) : m_t(t) {}
    constexpr operator Type_t() const { return m_t; }
    size_t size() const {
        switch (m_t) {
        case Type_t::boolean: case Type_t::i8: case Type_t::u8: case Type_t::i4: case Type_t::u4: case Type_t::u1: return 1;
        case Type_t::bf16: case Type_t::f16: case Type_t::i16: case Type_t::u16: return 2;
        case Type_t::f32: case Type_t::i32: case Type_t::u32: return 4;
        case Type_t::f64: case Type_t::i64: case Type_t::u64: return 8;
        case Type_t::string: return sizeof(std::string);
        default: return 0;
        }
    }
    bool is_static() const { return m_t != Type_t::dynamic; }
    bool is_dynamic() const { return m_t == Type_t::dynamic; }
    bool is_real() const { return m_t == Type_t::bf16 || m_t == Type_t::f16 || m_t == Type_t::f32 || m_t == Type_t::f64; }
    bool is_integral() const { return !is_real() && m_t != Type_t::dynamic && m_t != Type_t::string; }
    bool is_integral_number() const { return is_integral() && m_t != Type_t::boolean; }
    bool compatible(const Type& o) const { return is_dynamic() || o.is_dynamic() || m_t == o.m_t; }
    static bool merge(Type& dst, const Type& a, const Type& b) {
        if (a.is_dynamic()) { dst = b; return true; }
        if (b.is_dynamic() || a.m_t == b.m_t) { dst = a; return true; }
        return false;
    }
    std::string get_type_name() const {
        static const char* n[] = {"dynamic", "boolean", "bf16", "f16", "f32", "f64", "i4", "i8", "i16", "i32", "i64", "u1", "u4", "u8", "u16", "u32", "u64", "string"};
        return n[(int)m_t];
    }
    std::string to_string() const { return get_type_name(); }
    bool operator==(const Type& o) const { return m_t == o.m_t; }
    bool operator!=(const Type& o) const { return m_t != o.m_t; }
    bool operator==(Type_t o) const { return m_t == o; }
    bool operator!=(Type_t o) const { return m_t != o; }
private:
    Type_t m_t = Type_t::dynamic;
};
inline std::ostream& operator<<(std::ostream& s, const Type& t) { return s << t.get_type_name(); }
constexpr Type dynamic(Type_t::dynamic), boolean(Type_t::boolean), bf16(Type_t::bf16), f16(Type_t::f16), f32(Type_t::f32), f64(Type_t::f64),
    i4(Type_t::i4), i8(Type_t::i8), i16(Type_t::i16), i32(Type_t::i32), i64(Type_t::i64), u1(Type_t::u1), u4(Type_t::u4), u8(Type_t::u8),
    u16(Type_t::u16), u32(Type_t::u32), u64(Type_t::u64), string(Type_t::string);
template <class T> Type from();
template <> inline Type from<bool>() { return boolean; }
template <> inline Type from<char>() { return i8; }
template <> inline Type from<int8_t>() { return i8; }
template <> inline Type from<uint8_t>() { return u8; }
template <> inline Type from<int16_t>() { return i16; }
template <> inline Type from<uint16_t>() { return u16; }
template <> inline Type from<int32_t>() { return i32; }
template <> inline Type from<uint32_t>() { return u32; }
template <> inline Type from<int64_t>() { return i64; }
template <> inline Type from<uint64_t>() { return u64; }
template <> inline Type from<float>() { return f32; }
template <> inline Type from<double>() { return f64; }
template <> inline Type from<std::string>() { return string; }
}  // namespace element

// ------------------------------------------------------------------ shapes
class Shape : public std::vector<size_t> {
public:
    using std::vector<size_t>::vector;
    Shape() = default;
    Shape(const std::vector<size_t>& v) : std::vector<size_t>(v) {}
    std::string to_string() const {
        std::ostringstream s;
        s << "[";
        for (size_t i = 0; i < size(); ++i) s << (i ? "," : "") << (*this)[i];
        s << "]";
        return s.str();
    }
};
inline std::ostream& operator<<(std::ostream& s, const Shape& sh) { return s << sh.to_string(); }
inline size_t shape_size(const Shape& s) {
    size_t n = 1;
    for (auto d : s) n *= d;
    return n;
}

class Dimension {
public:
    using value_type = int64_t;
    Dimension() = default;                      // dynamic
    Dimension(value_type v) : m_v(v) {}
    bool is_static() const { return m_v >= 0; }
    bool is_dynamic() const { return m_v < 0; }
    value_type get_length() const {
        OPENVINO_ASSERT(is_static(), "Cannot get length of dynamic dimension");
        return m_v;
    }
    bool compatible(const Dimension& o) const { return is_dynamic() || o.is_dynamic() || m_v == o.m_v; }
    static bool merge(Dimension& dst, const Dimension& a, const Dimension& b) {
        if (a.is_dynamic()) { dst = b; return true; }
        if (b.is_dynamic() || a.m_v == b.m_v) { dst = a; return true; }
        return false;
    }
    static Dimension dynamic() { return Dimension(); }
    bool operator==(const Dimension& o) const { return m_v == o.m_v; }
    bool operator!=(const Dimension& o) const { return m_v != o.m_v; }
    std::string to_string() const { return is_dynamic() ? "?" : std::to_string(m_v); }
private:
    value_type m_v = -1;
};
inline std::ostream& operator<<(std::ostream& s, const Dimension& d) { return s << d.to_string(); }
using Rank = Dimension;

class PartialShape {
public:
    PartialShape() : m_rank_static(true) {}   // rank-0 static (scalar), like OpenVINO's default
    PartialShape(std::initializer_list<Dimension> d) : m_rank_static(true), m_dims(d) {}
    PartialShape(std::vector<Dimension> d) : m_rank_static(true), m_dims(std::move(d)) {}
    PartialShape(const Shape& s) : m_rank_static(true) { for (auto v : s) m_dims.emplace_back((int64_t)v); }
    static PartialShape dynamic(Rank r = Rank()) {
        PartialShape p;
        if (r.is_dynamic()) { p.m_rank_static = false; }
        else p.m_dims.assign((size_t)r.get_length(), Dimension());
        return p;
    }
    Rank rank() const { return m_rank_static ? Rank((int64_t)m_dims.size()) : Rank(); }
    bool is_static() const { return m_rank_static && std::all_of(m_dims.begin(), m_dims.end(), [](const Dimension& d) { return d.is_static(); }); }
    bool is_dynamic() const { return !is_static(); }
    size_t size() const { return m_dims.size(); }
    Dimension& operator[](size_t i) { OPENVINO_ASSERT(m_rank_static && i < m_dims.size(), "PartialShape index out of range"); return m_dims[i]; }
    const Dimension& operator[](size_t i) const { OPENVINO_ASSERT(m_rank_static && i < m_dims.size(), "PartialShape index out of range"); return m_dims[i]; }
    void push_back(const Dimension& d) { m_rank_static = true; m_dims.push_back(d); }
    std::vector<Dimension>::const_iterator begin() const { return m_dims.begin(); }
    std::vector<Dimension>::const_iterator end() const { return m_dims.end(); }
    Shape to_shape() const { OPENVINO_ASSERT(is_static(), "to_shape on dynamic shape"); Shape s; for (auto& d : m_dims) s.push_back((size_t)d.get_length()); return s; }
    Shape get_shape() const { return to_shape(); }
    bool compatible(const PartialShape& o) const {
        if (!m_rank_static || !o.m_rank_static) return true;
        if (m_dims.size() != o.m_dims.size()) return false;
        for (size_t i = 0; i < m_dims.size(); ++i) if (!m_dims[i].compatible(o.m_dims[i])) return false;
        return true;
    }
    bool merge_into(PartialShape& dst, const PartialShape& src) const { return merge_into_static(dst, src); }
    static bool merge_into_static(PartialShape& dst, const PartialShape& src) {
        if (!dst.m_rank_static) { dst = src; return true; }
        if (!src.m_rank_static) return true;
        if (dst.m_dims.size() != src.m_dims.size()) return false;
        for (size_t i = 0; i < dst.m_dims.size(); ++i) if (!Dimension::merge(dst.m_dims[i], dst.m_dims[i], src.m_dims[i])) return false;
        return true;
    }
    bool operator==(const PartialShape& o) const { return m_rank_static == o.m_rank_static && m_dims == o.m_dims; }
    bool operator!=(const PartialShape& o) const { return !(*this == o); }
    std::string to_string() const {
        if (!m_rank_static) return "[...]";
        std::ostringstream s;
        s << "[";
        for (size_t i = 0; i < m_dims.size(); ++i) s << (i ? "," : "") << m_dims[i];
        s << "]";
        return s.str();
    }
private:
    bool m_rank_static = true;
    std::vector<Dimension> m_dims;
};
inline std::ostream& operator<<(std::ostream& s, const PartialShape& p) { return s << p.to_string(); }

// ------------------------------------------------------------------ Tensor (host memory; shared handle like ov::Tensor)
class Tensor {
    struct Impl {
        element::Type type;
        Shape shape;
        std::vector<uint64_t> own;      // owned storage (8-byte aligned), grow-only like the runtime's allocator
        void* ext = nullptr;            // user memory (view); set_shape may only shrink / keep the byte size within ext_bytes
        size_t ext_bytes = 0;
        std::vector<std::string> strs;  // element::string storage
    };
    std::shared_ptr<Impl> m;
public:
    Tensor() = default;
    Tensor(const element::Type& t, const Shape& s) : m(std::make_shared<Impl>()) { m->type = t; set_shape(s); }
    Tensor(const element::Type& t, const Shape& s, void* host_ptr) : m(std::make_shared<Impl>()) {
        m->type = t; m->shape = s; m->ext = host_ptr; m->ext_bytes = shape_size(s) * t.size();
    }
    explicit operator bool() const { return (bool)m; }
    bool operator!() const { return !m; }
    const element::Type& get_element_type() const { return impl().type; }
    const Shape& get_shape() const { return impl().shape; }
    size_t get_size() const { return shape_size(impl().shape); }
    size_t get_byte_size() const { return get_size() * impl().type.size(); }
    void set_shape(const Shape& s) {
        Impl& i = impl();
        const size_t bytes = shape_size(s) * i.type.size();
        if (i.type == element::string) { i.strs.resize(shape_size(s)); i.shape = s; return; }
        if (i.ext) OPENVINO_ASSERT(bytes <= i.ext_bytes, "Could set new shape: ", s.to_string(), " (view over user memory cannot grow)");
        else if (bytes > i.own.size() * 8) i.own.resize((bytes + 7) / 8 + 1);
        i.shape = s;
    }
    void* data(const element::Type& = element::dynamic) const {
        const Impl& i = impl();
        if (i.type == element::string) return (void*)i.strs.data();
        return i.ext ? i.ext : (void*)(i.own.empty() ? nullptr : i.own.data());
    }
    template <class T>
    T* data() const {
        using U = std::remove_const_t<T>;
        const Impl& i = impl();
        if constexpr (!std::is_same_v<U, char> && !std::is_same_v<U, void>)
            OPENVINO_ASSERT(element::from<U>().size() == i.type.size() || i.type == element::dynamic,
                            "Tensor data with element type ", i.type, " is not representable as pointer to a ", sizeof(U), "-byte type");
        return static_cast<T*>(data());
    }
    void copy_to(Tensor dst) const {
        dst.impl().type = impl().type;
        dst.set_shape(get_shape());
        if (impl().type == element::string) { dst.impl().strs = impl().strs; return; }
        if (get_byte_size()) std::memcpy(dst.data(), data(), get_byte_size());
    }
    bool is_same(const Tensor& o) const { return m == o.m; }   // (stub extension: used by tests to check aliasing)
private:
    Impl& impl() const { OPENVINO_ASSERT(m, "Tensor was not initialized."); return *m; }
};
using TensorVector = std::vector<Tensor>;

// ------------------------------------------------------------------ Any / AnyMap
class Any {
    std::any m;
public:
    Any() = default;
    template <class T, class = std::enable_if_t<!std::is_same_v<std::decay_t<T>, Any>>>
    Any(T&& v) : m(std::forward<T>(v)) {}
    Any(const char* s) : m(std::string(s)) {}
    bool empty() const { return !m.has_value(); }
    template <class T> bool is() const { return m.type() == typeid(T); }
    template <class T>
    T as() const {
        if (m.type() == typeid(T)) return std::any_cast<T>(m);
        if constexpr (std::is_arithmetic_v<T>) {
            if (m.type() == typeid(int)) return (T)std::any_cast<int>(m);
            if (m.type() == typeid(int64_t)) return (T)std::any_cast<int64_t>(m);
            if (m.type() == typeid(size_t)) return (T)std::any_cast<size_t>(m);
            if (m.type() == typeid(bool)) return (T)std::any_cast<bool>(m);
            if (m.type() == typeid(float)) return (T)std::any_cast<float>(m);
            if (m.type() == typeid(double)) return (T)std::any_cast<double>(m);
        }
        OPENVINO_THROW("Bad cast from: ", m.type().name(), " to: ", typeid(T).name());
    }
};
using AnyMap = std::map<std::string, Any>;

// ------------------------------------------------------------------ type info / graph
struct DiscreteTypeInfo {
    const char* name;
    const char* version_id;
    const DiscreteTypeInfo* parent;
    bool is_castable(const DiscreteTypeInfo& t) const { return (std::strcmp(name, t.name) == 0 && std::strcmp(version_id ? version_id : "", t.version_id ? t.version_id : "") == 0) || (parent && parent->is_castable(t)); }
    bool operator==(const DiscreteTypeInfo& o) const { return std::strcmp(name, o.name) == 0 && std::strcmp(version_id ? version_id : "", o.version_id ? o.version_id : "") == 0; }
};

class Node;
template <class T> class Output;
template <class T> class Input;
using NodeVector = std::vector<std::shared_ptr<Node>>;
using OutputVector = std::vector<Output<Node>>;

namespace descriptor {
class Tensor {
public:
    element::Type type;
    PartialShape pshape = PartialShape::dynamic();
    std::set<std::string> names;
    const element::Type& get_element_type() const { return type; }
    const PartialShape& get_partial_shape() const { return pshape; }
    void add_names(const std::set<std::string>& n) { names.insert(n.begin(), n.end()); }
    void set_names(const std::set<std::string>& n) { names = n; }
    const std::set<std::string>& get_names() const { return names; }
};
}  // namespace descriptor

template <>
class Output<Node> {
public:
    Output() = default;
    Output(const std::shared_ptr<Node>& n, size_t i = 0) : m_node(n), m_index(i) {}
    template <class T, class = std::enable_if_t<std::is_base_of_v<Node, T>>>
    Output(const std::shared_ptr<T>& n) : m_node(n), m_index(0) {}
    Node* get_node() const { return m_node.get(); }
    std::shared_ptr<Node> get_node_shared_ptr() const { return m_node; }
    size_t get_index() const { return m_index; }
    inline const element::Type& get_element_type() const;
    inline const PartialShape& get_partial_shape() const;
    inline Shape get_shape() const;
    inline descriptor::Tensor& get_tensor() const;
    inline std::set<Input<Node>> get_target_inputs() const;
    inline void replace(const Output<Node>& replacement) const;
    bool operator==(const Output& o) const { return m_node == o.m_node && m_index == o.m_index; }
    bool operator!=(const Output& o) const { return !(*this == o); }
    bool operator<(const Output& o) const { return m_node < o.m_node || (m_node == o.m_node && m_index < o.m_index); }
private:
    std::shared_ptr<Node> m_node;
    size_t m_index = 0;
};

template <>
class Input<Node> {
public:
    Input(Node* n, size_t i) : m_node(n), m_index(i) {}
    Node* get_node() const { return m_node; }
    size_t get_index() const { return m_index; }
    inline Output<Node> get_source_output() const;
    inline void replace_source_output(const Output<Node>& o) const;
    bool operator<(const Input& o) const { return m_node < o.m_node || (m_node == o.m_node && m_index < o.m_index); }
    bool operator==(const Input& o) const { return m_node == o.m_node && m_index == o.m_index; }
private:
    Node* m_node;
    size_t m_index;
};

// ------------------------------------------------------------------ AttributeVisitor
// Concrete (de)serialiser over a string map: `on_attribute(name, value)` reads the value from the map when the visitor
// was built from attributes (IR load), or records it when it was built empty (serialisation / inspection).
class AttributeVisitor {
public:
    AttributeVisitor() : m_load(false) {}
    explicit AttributeVisitor(std::map<std::string, std::string> attrs) : m_load(true), m_attrs(std::move(attrs)) {}
    virtual ~AttributeVisitor() = default;
    const std::map<std::string, std::string>& attributes() const { return m_attrs; }
    void on_attribute(const std::string& name, std::string& v) { if (m_load) { auto it = m_attrs.find(name); if (it != m_attrs.end()) v = it->second; } else m_attrs[name] = v; }
    void on_attribute(const std::string& name, bool& v) {
        if (m_load) { auto it = m_attrs.find(name); if (it != m_attrs.end()) { std::string s = it->second; std::transform(s.begin(), s.end(), s.begin(), ::tolower); v = (s == "true" || s == "1"); } }
        else m_attrs[name] = v ? "true" : "false";
    }
    template <class T, class = std::enable_if_t<std::is_arithmetic_v<T> && !std::is_same_v<T, bool>>>
    void on_attribute(const std::string& name, T& v) {
        if (m_load) { auto it = m_attrs.find(name); if (it != m_attrs.end()) { if constexpr (std::is_floating_point_v<T>) v = (T)std::stod(it->second); else v = (T)std::stoll(it->second); } }
        else m_attrs[name] = std::to_string(v);
    }
    template <class T, class = std::enable_if_t<std::is_arithmetic_v<T>>>
    void on_attribute(const std::string& name, std::vector<T>& v) {
        if (m_load) {
            auto it = m_attrs.find(name);
            if (it == m_attrs.end()) return;
            v.clear();
            std::stringstream ss(it->second);
            std::string tok;
            while (std::getline(ss, tok, ',')) { size_t a = tok.find_first_not_of(" "); if (a == std::string::npos) continue; v.push_back((T)std::stod(tok.substr(a))); }
        } else {
            std::ostringstream s;
            for (size_t i = 0; i < v.size(); ++i) s << (i ? ", " : "") << v[i];
            m_attrs[name] = s.str();
        }
    }
private:
    bool m_load;
    std::map<std::string, std::string> m_attrs;
};

// ------------------------------------------------------------------ Node
class Node : public std::enable_shared_from_this<Node> {
public:
    using type_info_t = DiscreteTypeInfo;
    Node() = default;
    explicit Node(const OutputVector& args) { set_arguments(args); }
    virtual ~Node() { for (size_t i = 0; i < m_inputs.size(); ++i) unlink(i); }
    virtual const type_info_t& get_type_info() const = 0;
    const char* get_type_name() const { return get_type_info().name; }
    virtual void validate_and_infer_types() {}
    virtual std::shared_ptr<Node> clone_with_new_inputs(const OutputVector& inputs) const = 0;
    virtual bool visit_attributes(AttributeVisitor&) { return true; }
    virtual bool evaluate(TensorVector&, const TensorVector&) const { return false; }
    virtual bool has_evaluate() const { return false; }
    void constructor_validate_and_infer_types() { validate_and_infer_types(); }
    void set_input_is_relevant_to_shape(size_t, bool = true) {}
    void set_input_is_relevant_to_value(size_t, bool = true) {}

    void set_arguments(const OutputVector& args) {
        for (size_t i = 0; i < m_inputs.size(); ++i) unlink(i);
        m_inputs = args;
        for (size_t i = 0; i < m_inputs.size(); ++i) link(i);
    }
    void set_argument(size_t i, const Output<Node>& o) {
        if (i >= m_inputs.size()) m_inputs.resize(i + 1);
        else unlink(i);
        m_inputs[i] = o;
        link(i);
    }
    size_t get_input_size() const { return m_inputs.size(); }
    size_t get_output_size() const { return m_outputs.size(); }
    void set_output_size(size_t n) { if (m_outputs.size() < n) m_outputs.resize(n); }
    const element::Type& get_input_element_type(size_t i) const { return input_value(i).get_element_type(); }
    const PartialShape& get_input_partial_shape(size_t i) const { return input_value(i).get_partial_shape(); }
    Shape get_input_shape(size_t i) const { return get_input_partial_shape(i).to_shape(); }
    const element::Type& get_output_element_type(size_t i) const { return out_desc(i).type; }
    const PartialShape& get_output_partial_shape(size_t i) const { return out_desc(i).pshape; }
    Shape get_output_shape(size_t i) const { return out_desc(i).pshape.to_shape(); }
    void set_output_type(size_t i, const element::Type& t, const PartialShape& p) {
        set_output_size(i + 1);
        m_outputs[i].desc.type = t;
        m_outputs[i].desc.pshape = p;
    }
    const Output<Node>& input_value(size_t i) const { OPENVINO_ASSERT(i < m_inputs.size(), "input index ", i, " out of range (", m_inputs.size(), " inputs)"); return m_inputs[i]; }
    OutputVector input_values() const { return m_inputs; }
    Node* get_input_node_ptr(size_t i) const { return input_value(i).get_node(); }
    std::shared_ptr<Node> get_input_node_shared_ptr(size_t i) const { return input_value(i).get_node_shared_ptr(); }
    Output<Node> input_source(size_t i) const { return input_value(i); }
    Input<Node> input(size_t i) { return Input<Node>(this, i); }
    Output<Node> output(size_t i) { return Output<Node>(shared_from_this(), i); }
    OutputVector outputs() {
        OutputVector v;
        for (size_t i = 0; i < m_outputs.size(); ++i) v.emplace_back(shared_from_this(), i);
        return v;
    }
    std::set<Input<Node>> get_output_target_inputs(size_t i) const { return out_slot(i).targets; }
    descriptor::Tensor& get_output_tensor(size_t i) { set_output_size(i + 1); return m_outputs[i].desc; }
    void set_friendly_name(const std::string& n) { m_name = n; }
    const std::string& get_friendly_name() const { return m_name; }
    const std::string& get_name() const { return m_name; }
    std::map<std::string, Any>& get_rt_info() { return m_rt; }
    std::shared_ptr<Node> copy_with_new_inputs(const OutputVector& in) const { auto c = clone_with_new_inputs(in); c->m_name = m_name; return c; }

private:
    friend class Output<Node>;
    friend class Input<Node>;
    struct OutSlot { descriptor::Tensor desc; std::set<Input<Node>> targets; };
    const descriptor::Tensor& out_desc(size_t i) const { OPENVINO_ASSERT(i < m_outputs.size(), "output index ", i, " out of range"); return m_outputs[i].desc; }
    const OutSlot& out_slot(size_t i) const { OPENVINO_ASSERT(i < m_outputs.size(), "output index ", i, " out of range"); return m_outputs[i]; }
    void link(size_t i) { if (auto* src = m_inputs[i].get_node()) { src->set_output_size(m_inputs[i].get_index() + 1); src->m_outputs[m_inputs[i].get_index()].targets.insert(Input<Node>(this, i)); } }
    void unlink(size_t i) { if (auto* src = m_inputs[i].get_node()) { auto& t = src->m_outputs[m_inputs[i].get_index()].targets; t.erase(Input<Node>(this, i)); } }
    OutputVector m_inputs;
    std::vector<OutSlot> m_outputs;
    std::string m_name;
    std::map<std::string, Any> m_rt;
};

inline const element::Type& Output<Node>::get_element_type() const { return m_node->out_desc(m_index).type; }
inline const PartialShape& Output<Node>::get_partial_shape() const { return m_node->out_desc(m_index).pshape; }
inline Shape Output<Node>::get_shape() const { return get_partial_shape().to_shape(); }
inline descriptor::Tensor& Output<Node>::get_tensor() const { return m_node->get_output_tensor(m_index); }
inline std::set<Input<Node>> Output<Node>::get_target_inputs() const { return m_node->get_output_target_inputs(m_index); }
inline void Output<Node>::replace(const Output<Node>& r) const { for (auto& in : get_target_inputs()) in.replace_source_output(r); }
inline Output<Node> Input<Node>::get_source_output() const { return m_node->input_value(m_index); }
inline void Input<Node>::replace_source_output(const Output<Node>& o) const { m_node->set_argument(m_index, o); }

#define OPENVINO_OP(NAME, ...)                                                                         \
    static const ::ov::DiscreteTypeInfo& get_type_info_static() {                                      \
        static const ::ov::DiscreteTypeInfo info{NAME, "extension", nullptr};                          \
        return info;                                                                                   \
    }                                                                                                  \
    const ::ov::DiscreteTypeInfo& get_type_info() const override { return get_type_info_static(); }
#define OPENVINO_RTTI(NAME, ...) OPENVINO_OP(NAME)

template <class T, class U>
std::shared_ptr<T> as_type_ptr(const std::shared_ptr<U>& p) { return std::dynamic_pointer_cast<T>(p); }
template <class T, class U>
T* as_type(U* p) { return dynamic_cast<T*>(p); }
template <class T, class U>
bool is_type(const std::shared_ptr<U>& p) { return (bool)std::dynamic_pointer_cast<T>(p); }

namespace op {
class Op : public Node {
public:
    Op() = default;
    explicit Op(const OutputVector& args) : Node(args) {}
};
namespace v0 {
class Constant : public Op {
public:
    OPENVINO_OP("Constant");
    Constant() = default;
    Constant(const element::Type& t, const Shape& s, const void* data) : m_t(t, s) {
        if (t == element::string) { auto* src = static_cast<const std::string*>(data); auto* dst = m_t.data<std::string>(); for (size_t i = 0; i < shape_size(s); ++i) dst[i] = src[i]; }
        else if (m_t.get_byte_size()) std::memcpy(m_t.data(), data, m_t.get_byte_size());
        set_output_type(0, t, PartialShape(s));
    }
    template <class T>
    Constant(const element::Type& t, const Shape& s, const std::vector<T>& v) : m_t(t, s) {
        OPENVINO_ASSERT(v.size() == shape_size(s) || v.size() == 1, "Constant: value count does not match the shape");
        fill(v);
        set_output_type(0, t, PartialShape(s));
    }
    explicit Constant(const Tensor& t) : m_t(t) { set_output_type(0, t.get_element_type(), PartialShape(t.get_shape())); }
    template <class T>
    static std::shared_ptr<Constant> create(const element::Type& t, const Shape& s, const std::vector<T>& v) { return std::make_shared<Constant>(t, s, v); }
    const void* get_data_ptr() const { return m_t.data(); }
    template <class T> const T* get_data_ptr() const { return static_cast<const T*>(m_t.data()); }
    const Shape& get_shape() const { return m_t.get_shape(); }
    const element::Type& get_element_type() const { return m_t.get_element_type(); }
    size_t get_byte_size() const { return m_t.get_byte_size(); }
    const Tensor& get_tensor_view() const { return m_t; }
    template <class T>
    std::vector<T> cast_vector() const {
        std::vector<T> r(m_t.get_size());
        for (size_t i = 0; i < r.size(); ++i) r[i] = (T)get_as_double(i);
        return r;
    }
    template <class T> std::vector<T> get_vector() const { return cast_vector<T>(); }
    std::shared_ptr<Node> clone_with_new_inputs(const OutputVector&) const override { return std::make_shared<Constant>(m_t); }
    bool has_evaluate() const override { return true; }
    bool evaluate(TensorVector& out, const TensorVector&) const override { out[0] = m_t; return true; }
private:
    template <class T>
    void fill(const std::vector<T>& v) {
        const size_t n = m_t.get_size();
        for (size_t i = 0; i < n; ++i) {
            const double x = (double)v[v.size() == 1 ? 0 : i];
            switch ((element::Type_t)m_t.get_element_type()) {
            case element::Type_t::boolean: case element::Type_t::u8: static_cast<uint8_t*>(m_t.data())[i] = (uint8_t)x; break;
            case element::Type_t::i8: static_cast<int8_t*>(m_t.data())[i] = (int8_t)x; break;
            case element::Type_t::i32: static_cast<int32_t*>(m_t.data())[i] = (int32_t)x; break;
            case element::Type_t::i64: static_cast<int64_t*>(m_t.data())[i] = (int64_t)x; break;
            case element::Type_t::f32: static_cast<float*>(m_t.data())[i] = (float)x; break;
            default: OPENVINO_THROW("stub Constant: unsupported element type ", m_t.get_element_type());
            }
        }
    }
    double get_as_double(size_t i) const {
        switch ((element::Type_t)m_t.get_element_type()) {
        case element::Type_t::boolean: case element::Type_t::u8: return static_cast<const uint8_t*>(m_t.data())[i];
        case element::Type_t::i8: return static_cast<const int8_t*>(m_t.data())[i];
        case element::Type_t::i32: return static_cast<const int32_t*>(m_t.data())[i];
        case element::Type_t::i64: return (double)static_cast<const int64_t*>(m_t.data())[i];
        case element::Type_t::f32: return static_cast<const float*>(m_t.data())[i];
        default: OPENVINO_THROW("stub Constant: unsupported element type ", m_t.get_element_type());
        }
    }
    Tensor m_t;
};
class Parameter : public Op {
public:
    OPENVINO_OP("Parameter");
    Parameter() = default;
    Parameter(const element::Type& t, const PartialShape& p) { set_output_type(0, t, p); }
    void set_partial_shape(const PartialShape& p) { set_output_type(0, get_output_element_type(0), p); }
    void set_element_type(const element::Type& t) { set_output_type(0, t, get_output_partial_shape(0)); }
    std::shared_ptr<Node> clone_with_new_inputs(const OutputVector&) const override { return std::make_shared<Parameter>(get_output_element_type(0), get_output_partial_shape(0)); }
};
class Result : public Op {
public:
    OPENVINO_OP("Result");
    Result() = default;
    explicit Result(const Output<Node>& a) : Op({a}) { validate_and_infer_types(); }
    void validate_and_infer_types() override { set_output_type(0, get_input_element_type(0), get_input_partial_shape(0)); }
    std::shared_ptr<Node> clone_with_new_inputs(const OutputVector& in) const override { return std::make_shared<Result>(in.at(0)); }
};
}  // namespace v0
// Only what the reference's translation helpers mention (src/utils.cpp:128-152); never evaluated here.
namespace v15 {
class StringTensorPack : public Op {
public:
    OPENVINO_OP("StringTensorPack");
    StringTensorPack() = default;
    StringTensorPack(const Output<Node>& b, const Output<Node>& e, const Output<Node>& c) : Op({b, e, c}) { set_output_type(0, element::string, b.get_partial_shape()); }
    std::shared_ptr<Node> clone_with_new_inputs(const OutputVector& in) const override { return std::make_shared<StringTensorPack>(in.at(0), in.at(1), in.at(2)); }
};
class StringTensorUnpack : public Op {
public:
    OPENVINO_OP("StringTensorUnpack");
    StringTensorUnpack() = default;
    explicit StringTensorUnpack(const Output<Node>& s) : Op({s}) {
        set_output_type(0, element::i32, s.get_partial_shape());
        set_output_type(1, element::i32, s.get_partial_shape());
        set_output_type(2, element::u8, PartialShape{Dimension()});
    }
    std::shared_ptr<Node> clone_with_new_inputs(const OutputVector& in) const override { return std::make_shared<StringTensorUnpack>(in.at(0)); }
};
}  // namespace v15
namespace util {
class FrameworkNode : public Op {
public:
    OPENVINO_OP("FrameworkNode");
    using Op::Op;
};
}  // namespace util
}  // namespace op
namespace opset13 { using namespace ::ov::op::v0; }
namespace opset15 { using namespace ::ov::op::v0; using ::ov::op::v15::StringTensorPack; using ::ov::op::v15::StringTensorUnpack; }

// ------------------------------------------------------------------ parallel helpers (openvino/core/parallel.hpp)
namespace stub {
inline int parallel_threads() {
    static const int n = [] { const char* e = std::getenv("OV_STUB_THREADS"); return e ? std::max(1, atoi(e)) : 1; }();
    return n;
}
}  // namespace stub
template <class T0, class F>
void parallel_for(const T0& n, const F& f) {
    const size_t N = (size_t)n;
    const int T = (int)std::min<size_t>((size_t)stub::parallel_threads(), N ? N : 1);
    if (T <= 1) { for (size_t i = 0; i < N; ++i) f(i); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < T; ++t) th.emplace_back([&, t] { for (size_t i = N * t / T; i < N * (t + 1) / T; ++i) f(i); });
    for (auto& x : th) x.join();
}
template <class T0, class R, class F>
R parallel_sum(const T0& n, const R& init, const F& f) {
    R s = init;
    for (size_t i = 0; i < (size_t)n; ++i) s += f(i);
    return s;
}

// ------------------------------------------------------------------ extensions
class Extension {
public:
    using Ptr = std::shared_ptr<Extension>;
    virtual ~Extension() = default;
};
class BaseOpExtension : public Extension {
public:
    using Ptr = std::shared_ptr<BaseOpExtension>;
    virtual const DiscreteTypeInfo& get_type_info() const = 0;
    // What the IR frontend calls for a layer of this type: `inputs` are the already-built producers, `visitor` reads the
    // layer's <data> attributes.  The returned outputs replace the layer.
    virtual OutputVector create(const OutputVector& inputs, AttributeVisitor& visitor) const = 0;
    virtual std::vector<Extension::Ptr> get_attached_extensions() const { return {}; }
};
template <class T>
class OpExtension : public BaseOpExtension {
public:
    const DiscreteTypeInfo& get_type_info() const override { return T::get_type_info_static(); }
    OutputVector create(const OutputVector& inputs, AttributeVisitor& visitor) const override {
        auto node = std::make_shared<T>();
        node->set_arguments(inputs);
        if (node->visit_attributes(visitor)) node->constructor_validate_and_infer_types();
        return node->outputs();
    }
};
#define OPENVINO_EXTENSION_C_API extern "C" __attribute__((visibility("default")))
#define OPENVINO_EXTENSION_API __attribute__((visibility("default")))
#define OPENVINO_API_C(...) extern "C" __attribute__((visibility("default"))) __VA_ARGS__
#define OPENVINO_CREATE_EXTENSIONS(extensions)                                                     \
    OPENVINO_EXTENSION_C_API void create_extensions(std::vector<::ov::Extension::Ptr>& ext);       \
    OPENVINO_EXTENSION_C_API void create_extensions(std::vector<::ov::Extension::Ptr>& ext) { ext = extensions; }

// ------------------------------------------------------------------ frontend bits the reference's helpers mention
namespace frontend {
class NodeContext {
public:
    virtual ~NodeContext() = default;
    template <class T>
    T get_attribute(const std::string& name) const { auto it = attrs.find(name); OPENVINO_ASSERT(it != attrs.end(), "no attribute ", name); return it->second.as<T>(); }
    size_t get_input_size() const { return inputs.size(); }
    Output<Node> get_input(int i) const { return inputs.at((size_t)i); }
    AnyMap attrs;
    OutputVector inputs;
};
}  // namespace frontend

}  // namespace ov
