// ov_extension_b200.cpp — the host side that stays C++ inside OpenVINO: registers the SAME op names, attributes and
// input/output signatures as the reference extension (src/ov_extension.cpp:72-109) so a converted tokenizer IR loads
// unchanged, and forwards each evaluate() to the C ABI of libb200tok.so (include/b200tok.h).
//
// Build:  g++ -shared -fPIC -DIMPLEMENT_OPENVINO_EXTENSION_API ov_extension_b200.cpp -I<openvino>/include -I../../../include
//   [-I<reference>/src for precompiled_charsmap.hpp] -L.. -lb200tok -lopenvino   (INTEGRATION.md).
// OpenVINO is not installed in the build container; there the file is compiled against the stand-in API of
// tests/ov_stub by tests/shimlib.py (CPU tier: it must compile and export the entry points) and its evaluate() bodies and the
// load-time fusion below are driven by tests/test_ov_shim.py on the GPU box, against the reference's own op classes.
//
// Load-time fusion: a converted IR has RegexSplit -> BPETokenizer (and RegexSplit -> RegexSplit -> WordpieceTokenizer) as separate
// layers (python/openvino_tokenizers/tokenizer_pipeline.py:1600-1646); evaluated one by one, the piece offsets between them
// (8 bytes per ~1.8-byte piece) would cross PCIe twice.  The IR frontend creates every layer through its OpExtension::create
// (inputs = the producers already built, visitor = the layer's attributes) and wires the consumers to whatever outputs it
// returns — so the extensions registered for "BPETokenizer" / "WordpieceTokenizer" look at their producers and, when those are
// this library's RegexSplit layers in a supported configuration whose only consumer is the tokenizer, return the outputs of ONE
// fused op (B200SplitBPE / B200SplitWordpiece) fed by the splitters' own inputs.  The by-passed RegexSplit layers have no
// consumer left and drop out of the model; the IR file itself is unchanged.
#if __has_include(<openvino/op/op.hpp>)
#include <openvino/core/extension.hpp>
#include <openvino/core/op_extension.hpp>
#include <openvino/op/constant.hpp>
#include <openvino/op/op.hpp>

#include <algorithm>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "b200tok.h"
// CharsMapNormalization's attribute form needs the charsmaps the reference generates at build time (src/precompiled_charsmap.hpp:
// get_precompiled_charsmap(form, case_fold)); present when the shim is built inside the reference tree.
#if __has_include("precompiled_charsmap.hpp")
#include "precompiled_charsmap.hpp"
#else
inline std::string get_precompiled_charsmap(const std::string&, bool) { return std::string(); }   // attribute form unavailable: pass the charsmap as an input
#endif

namespace b200 {

inline void check(int rc) { OPENVINO_ASSERT(rc == B200TOK_OK, "b200tok: ", b200tok_last_error()); }

// CUDA device the ops of this process run on: B200TOK_DEVICE (default 0) — one process per GPU, like the rest of the stack.
inline int device() {
    static const int d = [] { const char* e = std::getenv("B200TOK_DEVICE"); return e ? std::atoi(e) : 0; }();
    return d;
}

inline b200tok_strings strings_of(const ov::TensorVector& in, size_t i) {
    return b200tok_strings{in[i].data<const int32_t>(), in[i + 1].data<const int32_t>(), in[i + 2].data<const uint8_t>(),
                           (int64_t)in[i].get_size(), (int64_t)in[i + 2].get_size()};
}
inline b200tok_ragged_strings ragged_of(const ov::TensorVector& in, const uint8_t* skips = nullptr) {
    return b200tok_ragged_strings{in[0].data<const int32_t>(), in[1].data<const int32_t>(), (int64_t)in[0].get_size(),
                                  in[2].data<const int32_t>(), in[3].data<const int32_t>(), (int64_t)in[2].get_size(),
                                  in[4].data<const uint8_t>(), (int64_t)in[4].get_size(), skips, B200TOK_MEM_HOST};
}
inline void ragged_ids_out(ov::TensorVector& out, const ov::TensorVector& in, int64_t capacity,
                           const std::function<int(b200tok_ragged_ids*)>& run) {
    out[0].set_shape(in[0].get_shape());
    out[1].set_shape(in[1].get_shape());
    out[2].set_shape({(size_t)capacity});                       // worst case first (src/bpe_tokenizer.cpp:135)
    b200tok_ragged_ids r{out[0].data<int32_t>(), out[1].data<int32_t>(), out[2].data<int32_t>(), capacity, 0, nullptr, B200TOK_MEM_HOST};
    check(run(&r));
    out[2].set_shape({(size_t)r.n_ids});                        // then the real size (src/bpe_tokenizer.cpp:162)
}

struct Handle {   // shared by clones, like the reference's shared_ptr<BPETokenizerImpl> (src/bpe_tokenizer.hpp:215-218)
    b200tok_handle h = nullptr;
    std::once_flag once, once2;
    ~Handle() { b200tok_destroy(h); }
};

// ---- RegexSplit (src/regex_split.hpp:15-73) ---------------------------------------------------------------------
class RegexSplit : public ov::op::Op {
public:
    OPENVINO_OP("RegexSplit");
    RegexSplit() = default;
    RegexSplit(const ov::OutputVector& args, const std::string& behaviour = "remove", bool invert = false, int max_splits = -1)
        : ov::op::Op(args), m_behaviour(behaviour), m_invert(invert), m_max_splits(max_splits) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        const auto n = get_input_size();
        OPENVINO_ASSERT(n == 6 || n == 7 || n == 9, "Incorrect number of inputs passed to RegexSplit: ", n,
                        "; try to reconvert tokenizer with newer version of OpenVINO Tokenizers");      // src/regex_split.cpp:102
        for (size_t i = 0; i < 4; ++i) set_output_type(i, ov::element::i32, i < 2 ? get_input_partial_shape(0) : ov::PartialShape{ov::Dimension()});
        set_output_type(4, ov::element::u8, ov::PartialShape{ov::Dimension()});
        if (n == 7) set_output_type(5, get_input_element_type(5), get_input_partial_shape(5));
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override {
        auto c = std::make_shared<RegexSplit>(in, m_behaviour, m_invert, m_max_splits);
        c->m_state = m_state;
        return c;
    }
    bool visit_attributes(ov::AttributeVisitor& v) override { return visit_prefixed(v, ""); }
    bool visit_prefixed(ov::AttributeVisitor& v, const std::string& prefix) {      // (the fused layers store two splitters' attributes)
        v.on_attribute(prefix + "behaviour", m_behaviour);
        v.on_attribute(prefix + "invert", m_invert);
        v.on_attribute(prefix + "max_splits", m_max_splits);
        return true;
    }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        const bool has_skips = in.size() == 7;
        const auto& pt = in[5 + has_skips];
        ensure(pt.data<const char>(), pt.get_size());
        if (in.size() == 9 && in[6].get_size() > 0)                       // legacy skip tokens, set once (src/regex_split.cpp:164-178)
            std::call_once(m_state->once2, [&] {
                const b200tok_strings toks = strings_of(in, 6);
                check(b200tok_regexsplit_set_skip_tokens(m_state->h, &toks));
            });
        const size_t cap = in[4].get_size() + in[2].get_size();          // src/regex_split.cpp:182
        out[0].set_shape(in[0].get_shape());
        out[1].set_shape(in[1].get_shape());
        out[2].set_shape({cap});
        out[3].set_shape({cap});
        out[4] = in[4];                                                   // chars are aliased (src/regex_split.cpp:203)
        if (has_skips) out[5].set_shape({cap});
        auto rin = ragged_of(in, has_skips ? reinterpret_cast<const uint8_t*>(in[5].data<bool>()) : nullptr);
        b200tok_ragged_strings_out r{out[0].data<int32_t>(), out[1].data<int32_t>(), out[2].data<int32_t>(), out[3].data<int32_t>(),
                                     has_skips ? reinterpret_cast<uint8_t*>(out[5].data<bool>()) : nullptr, (int64_t)cap, 0, 0, B200TOK_MEM_HOST};
        check(b200tok_regexsplit_run(m_state->h, &rin, &r, nullptr));
        if ((size_t)r.n_rows != in[0].get_size()) { out[0].set_shape({(size_t)r.n_rows}); out[1].set_shape({(size_t)r.n_rows}); }  // :129-143
        out[2].set_shape({(size_t)r.n_elems});
        out[3].set_shape({(size_t)r.n_elems});
        if (has_skips) out[5].set_shape({(size_t)r.n_elems});
        return true;
    }
    // lazily compiled splitter (src/regex_split.cpp:147-151), shared with clones and with the fused layers
    void ensure(const char* pattern, size_t len) const {
        std::call_once(m_state->once, [&] {
            b200tok_regexsplit_desc d{pattern, (int64_t)len, m_behaviour.c_str(), m_invert, m_max_splits, device()};
            check(b200tok_regexsplit_create(&d, &m_state->h));
        });
    }
    b200tok_handle handle() const { return m_state->h; }
    const std::string& behaviour() const { return m_behaviour; }
    int max_splits() const { return m_max_splits; }
private:
    std::string m_behaviour = "remove";
    bool m_invert = false;
    int m_max_splits = -1;
    mutable std::shared_ptr<Handle> m_state = std::make_shared<Handle>();
};

// ---- BPETokenizer (src/bpe_tokenizer.hpp:168-246) ---------------------------------------------------------------
class BPETokenizer : public ov::op::Op {
public:
    OPENVINO_OP("BPETokenizer");
    BPETokenizer() = default;
    BPETokenizer(const ov::OutputVector& args, const std::string& unk_token = "", bool fuse_unk = false,
                 const std::string& suffix_indicator = "", const std::string& end_suffix = "", bool byte_fallback = false,
                 size_t cache_capacity = 20000)
        : ov::op::Op(args), m_unk_token(unk_token), m_fuse_unk(fuse_unk), m_suffix_indicator(suffix_indicator),
          m_end_suffix(end_suffix), m_byte_fallback(byte_fallback), m_cache_capacity(cache_capacity) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        const auto n = get_input_size();
        OPENVINO_ASSERT(n == 11 || n == 14 || n == 15 || n == 18,
                        "Incorrect number of inputs passed to BPETokenizer, try to reconvert tokenizer with newer version of OpenVINO Tokenizers");
        set_output_type(0, ov::element::i32, get_input_partial_shape(0));
        set_output_type(1, ov::element::i32, get_input_partial_shape(0));
        set_output_type(2, ov::element::i32, ov::PartialShape{ov::Dimension()});
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override {
        auto c = std::make_shared<BPETokenizer>(in, m_unk_token, m_fuse_unk, m_suffix_indicator, m_end_suffix, m_byte_fallback, m_cache_capacity);
        c->m_state = m_state;
        return c;
    }
    bool visit_attributes(ov::AttributeVisitor& v) override {
        v.on_attribute("unk_token", m_unk_token);
        v.on_attribute("fuse_unk", m_fuse_unk);
        v.on_attribute("suffix_indicator", m_suffix_indicator);
        v.on_attribute("end_suffix", m_end_suffix);
        v.on_attribute("byte_fallback", m_byte_fallback);
        v.on_attribute("cache_capacity", m_cache_capacity);
        return true;
    }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        ensure(in);
        auto rin = ragged_of(in);
        ragged_ids_out(out, in, capacity(in),
                       [&](b200tok_ragged_ids* r) { return b200tok_bpe_run(m_state->h, &rin, r, nullptr); });
        return true;
    }
    // device tables built once from the Constant inputs [5..] of the 11 / 14 / 15 / 18-input forms (src/bpe_tokenizer.cpp:50-120)
    void ensure(const ov::TensorVector& in) const {
        const auto n = in.size();
        std::call_once(m_state->once, [&] {
            b200tok_bpe_desc d{};
            d.vocab = strings_of(in, 5);
            d.merges_left = strings_of(in, 8);
            if (n == 14 || n == 18) d.merges_right = strings_of(in, 11);
            if (n == 15 || n == 18) { d.added_tokens = strings_of(in, n - 4); d.added_ids = in[n - 1].data<const int32_t>(); }
            d.unk_token = m_unk_token.data(); d.unk_token_len = (int64_t)m_unk_token.size();
            d.suffix_indicator = m_suffix_indicator.data(); d.suffix_indicator_len = (int64_t)m_suffix_indicator.size();
            d.end_suffix = m_end_suffix.data(); d.end_suffix_len = (int64_t)m_end_suffix.size();
            d.fuse_unk = m_fuse_unk; d.byte_fallback = m_byte_fallback; d.cache_capacity = (int64_t)m_cache_capacity; d.device = device();
            check(b200tok_bpe_create(&d, &m_state->h));
        });
    }
    int64_t capacity(const ov::TensorVector& in) const { return (int64_t)(in[4].get_size() + in[2].get_size() * m_end_suffix.size()); }
    b200tok_handle handle() const { return m_state->h; }
private:
    std::string m_unk_token;
    bool m_fuse_unk = false;
    std::string m_suffix_indicator, m_end_suffix;
    bool m_byte_fallback = false;
    size_t m_cache_capacity = 20000;
    mutable std::shared_ptr<Handle> m_state = std::make_shared<Handle>();
};

// ---- WordpieceTokenizer (src/wordpiece_tokenizer.hpp:15-59) -------------------------------------------------------
class WordpieceTokenizer : public ov::op::Op {
public:
    OPENVINO_OP("WordpieceTokenizer");
    WordpieceTokenizer() = default;
    WordpieceTokenizer(const ov::OutputVector& args, const std::string& suffix_indicator = "##", int max_bytes_per_word = 100)
        : ov::op::Op(args), m_suffix_indicator(suffix_indicator), m_max_bytes_per_word(max_bytes_per_word) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        set_output_type(0, ov::element::i32, get_input_partial_shape(0));
        set_output_type(1, ov::element::i32, get_input_partial_shape(0));
        set_output_type(2, ov::element::i32, ov::PartialShape{ov::Dimension()});
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override {
        auto c = std::make_shared<WordpieceTokenizer>(in, m_suffix_indicator, m_max_bytes_per_word);
        c->m_state = m_state;
        return c;
    }
    bool visit_attributes(ov::AttributeVisitor& v) override {
        v.on_attribute("suffix_indicator", m_suffix_indicator);
        v.on_attribute("max_bytes_per_word", m_max_bytes_per_word);
        return true;
    }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        ensure(in);
        const int32_t unk = *in[8].data<const int32_t>();
        auto rin = ragged_of(in);
        ragged_ids_out(out, in, (int64_t)(in[4].get_size() + in[2].get_size()),
                       [&](b200tok_ragged_ids* r) { return b200tok_wordpiece_run(m_state->h, &rin, unk, r, nullptr); });
        return true;
    }
    void ensure(const ov::TensorVector& in) const {
        std::call_once(m_state->once, [&] {
            b200tok_wordpiece_desc d{strings_of(in, 5), m_suffix_indicator.data(), (int64_t)m_suffix_indicator.size(), m_max_bytes_per_word, device()};
            check(b200tok_wordpiece_create(&d, &m_state->h));
        });
    }
    b200tok_handle handle() const { return m_state->h; }
private:
    std::string m_suffix_indicator = "##";
    int m_max_bytes_per_word = 100;
    mutable std::shared_ptr<Handle> m_state = std::make_shared<Handle>();
};

// ---- VocabEncoder (src/vocab_encoder.hpp) / VocabDecoder (src/vocab_decoder.hpp) / ByteFallback (src/byte_fallback.hpp)
class VocabEncoder : public ov::op::Op {
public:
    OPENVINO_OP("VocabEncoder");
    VocabEncoder() = default;
    explicit VocabEncoder(const ov::OutputVector& args) : ov::op::Op(args) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override { set_output_type(0, get_input_element_type(6), get_input_partial_shape(0)); }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override {
        auto c = std::make_shared<VocabEncoder>(in);
        c->m_state = m_state;
        return c;
    }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        const bool i64 = in[6].get_element_type() == ov::element::i64;
        OPENVINO_ASSERT(i64 || in[6].get_element_type() == ov::element::i32, "VocabEncoder: unsupported element type: ", in[6].get_element_type());
        std::call_once(m_state->once, [&] {
            b200tok_vocabenc_desc d{strings_of(in, 3), in[6].data(), i64, device()};
            check(b200tok_vocabenc_create(&d, &m_state->h));
        });
        out[0].set_shape({in[0].get_size()});
        const int64_t def = i64 ? *in[7].data<const int64_t>() : (int64_t)*in[7].data<const int32_t>();
        check(b200tok_vocabenc_run(m_state->h, in[0].data<const int32_t>(), in[1].data<const int32_t>(), (int64_t)in[0].get_size(),
                                   in[2].data<const uint8_t>(), (int64_t)in[2].get_size(), def, out[0].data(), B200TOK_MEM_HOST, nullptr));
        return true;
    }
private:
    mutable std::shared_ptr<Handle> m_state = std::make_shared<Handle>();
};

class VocabDecoder : public ov::op::Op {
public:
    OPENVINO_OP("VocabDecoder");
    VocabDecoder() = default;
    VocabDecoder(const ov::OutputVector& args, std::vector<int> skip_tokens = {}) : ov::op::Op(args), m_skip_tokens(std::move(skip_tokens)) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        const auto shape = get_input_partial_shape(0);
        set_output_type(0, ov::element::i32, {shape[0]});
        set_output_type(1, ov::element::i32, {shape[0]});
        set_output_type(2, ov::element::i32, ov::PartialShape{ov::Dimension()});
        set_output_type(3, ov::element::i32, ov::PartialShape{ov::Dimension()});
        set_output_type(4, ov::element::u8, ov::PartialShape{ov::Dimension()});
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override {
        auto c = std::make_shared<VocabDecoder>(in, m_skip_tokens);
        c->m_state = m_state;
        return c;
    }
    bool visit_attributes(ov::AttributeVisitor& v) override { v.on_attribute("skip_tokens", m_skip_tokens); return true; }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        OPENVINO_ASSERT(in.size() == 4 || in.size() == 5, "Too few inputs passed to VocabDecoder, it means it is not converted properly or it is not used in the supported pattern");
        std::call_once(m_state->once, [&] {
            b200tok_vocabdec_desc d{strings_of(in, 1), device()};
            check(b200tok_vocabdec_create(&d, &m_state->h));
        });
        const int64_t B = (int64_t)in[0].get_shape()[0], S = (int64_t)in[0].get_shape()[1], W = S > 0 ? S : 1;
        const int32_t* skip = in.size() == 5 ? in[4].data<const int32_t>() : m_skip_tokens.data();
        const int64_t n_skip = in.size() == 5 ? (int64_t)in[4].get_shape()[0] : (int64_t)m_skip_tokens.size();
        const int64_t cap = std::max<int64_t>(b200tok_vocabdec_max_chars(m_state->h, B, S), 1);
        out[0].set_shape({(size_t)B}); out[1].set_shape({(size_t)B});
        out[2].set_shape({(size_t)(B * W)}); out[3].set_shape({(size_t)(B * W)});
        out[4].set_shape({(size_t)cap});
        b200tok_decoded r{out[0].data<int32_t>(), out[1].data<int32_t>(), out[2].data<int32_t>(), out[3].data<int32_t>(),
                          out[4].data<uint8_t>(), cap, 0, B200TOK_MEM_HOST};
        check(b200tok_vocabdec_run(m_state->h, in[0].data<const int32_t>(), B, S, skip, n_skip, /*byte_fallback=*/0, &r, B200TOK_MEM_HOST, nullptr));
        out[4].set_shape({(size_t)r.n_chars});
        return true;
    }
private:
    std::vector<int> m_skip_tokens;
    mutable std::shared_ptr<Handle> m_state = std::make_shared<Handle>();
};

class ByteFallback : public ov::op::Op {
public:
    OPENVINO_OP("ByteFallback");
    ByteFallback() = default;
    explicit ByteFallback(const ov::OutputVector& args) : ov::op::Op(args) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        set_output_type(0, ov::element::i32, get_input_partial_shape(0));
        set_output_type(1, ov::element::i32, get_input_partial_shape(0));
        set_output_type(2, ov::element::u8, ov::PartialShape{ov::Dimension()});
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override { return std::make_shared<ByteFallback>(in); }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        out[0].set_shape(in[0].get_shape()); out[1].set_shape(in[1].get_shape()); out[2].set_shape({in[2].get_size()});
        int64_t n_chars = 0;
        check(b200tok_bytefallback_run(device(), in[0].data<const int32_t>(), in[1].data<const int32_t>(), (int64_t)in[0].get_size(),
                                       in[2].data<const uint8_t>(), (int64_t)in[2].get_size(), out[0].data<int32_t>(), out[1].data<int32_t>(),
                                       out[2].data<uint8_t>(), &n_chars, B200TOK_MEM_HOST, nullptr));
        out[2].set_shape({(size_t)n_chars});
        return true;
    }
};

// ---- SpecialTokensSplit (src/special_tokens_split.hpp; evaluate src/special_tokens_split.cpp:61-162) --------------------
class SpecialTokensSplit : public ov::op::Op {
public:
    OPENVINO_OP("SpecialTokensSplit");
    SpecialTokensSplit() = default;
    explicit SpecialTokensSplit(const ov::OutputVector& args) : ov::op::Op(args) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        const auto n = get_input_size();
        OPENVINO_ASSERT(n == 6 || n == 7, "Incorrect number of inputs passed to SpecialTokensSplit: ", n,
                        "; try to reconvert tokenizer with newer version of OpenVINO Tokenizers");
        for (size_t i = 0; i < 4; ++i) set_output_type(i, ov::element::i32, i < 2 ? get_input_partial_shape(0) : ov::PartialShape{ov::Dimension()});
        set_output_type(4, ov::element::u8, ov::PartialShape{ov::Dimension()});
        set_output_type(5, ov::element::boolean, ov::PartialShape{ov::Dimension()});
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override {
        auto c = std::make_shared<SpecialTokensSplit>(in);
        c->m_state = m_state;
        return c;
    }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        const bool has_skips = in.size() == 7;
        std::call_once(m_state->once, [&] {                                // src/special_tokens_split.cpp:65-69
            const auto& p = in[5 + has_skips];
            check(b200tok_specialsplit_create(p.data<const char>(), (int64_t)p.get_size(), device(), &m_state->h));
        });
        const size_t cap = in[4].get_size() + in[2].get_size();
        out[0].set_shape(in[0].get_shape());
        out[1].set_shape(in[1].get_shape());
        out[2].set_shape({cap}); out[3].set_shape({cap}); out[5].set_shape({cap});
        out[4] = in[4];                                                     // :93
        auto rin = ragged_of(in, has_skips ? reinterpret_cast<const uint8_t*>(in[5].data<bool>()) : nullptr);
        b200tok_ragged_strings_out r{out[0].data<int32_t>(), out[1].data<int32_t>(), out[2].data<int32_t>(), out[3].data<int32_t>(),
                                     reinterpret_cast<uint8_t*>(out[5].data<bool>()), (int64_t)cap, 0, 0, B200TOK_MEM_HOST};
        check(b200tok_specialsplit_run(m_state->h, &rin, &r, nullptr));
        out[2].set_shape({(size_t)r.n_elems}); out[3].set_shape({(size_t)r.n_elems}); out[5].set_shape({(size_t)r.n_elems});   // :155-158
        return true;
    }
private:
    mutable std::shared_ptr<Handle> m_state = std::make_shared<Handle>();
};

// ---- Truncate (src/truncate.cpp:37-147): outputs alias the inputs and are edited in place ------------------------------------
class Truncate : public ov::op::Op {
public:
    OPENVINO_OP("Truncate");
    Truncate() = default;
    Truncate(const ov::OutputVector& args, int num_inputs = 1) : ov::op::Op(args), m_num_inputs(num_inputs) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        for (int i = 0; i < m_num_inputs; ++i)
            for (int k = 0; k < 3; ++k) set_output_type(3 * i + k, get_input_element_type(3 * i + k), get_input_partial_shape(3 * i + k));
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override { return std::make_shared<Truncate>(in, m_num_inputs); }
    bool visit_attributes(ov::AttributeVisitor& v) override { v.on_attribute("m_num_inputs", m_num_inputs); return true; }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        const size_t n = in.size();
        const int32_t max_length = in[n - 3].data<const int32_t>()[0];
        const std::string side(in[n - 2].data<const char>(), in[n - 2].get_size()), mode(in[n - 1].data<const char>(), in[n - 1].get_size());
        for (int i = 0; i < 3 * m_num_inputs; ++i) out[i] = in[i];                                   // :52-56
        int32_t* b1 = m_num_inputs == 2 ? out[3].data<int32_t>() : nullptr;
        int32_t* e1 = m_num_inputs == 2 ? out[4].data<int32_t>() : nullptr;
        check(b200tok_truncate_run(device(), m_num_inputs, out[0].data<int32_t>(), out[1].data<int32_t>(), b1, e1, (int64_t)out[0].get_size(), max_length,
                                   side.c_str(), mode.c_str(), B200TOK_MEM_HOST, nullptr));
        return true;
    }
private:
    int m_num_inputs = 1;
};

// ---- CombineSegments (src/combine_segments.cpp:36-134), i32 elements ---------------------------------------------------------
class CombineSegments : public ov::op::Op {
public:
    OPENVINO_OP("CombineSegments");
    CombineSegments() = default;
    explicit CombineSegments(const ov::OutputVector& args) : ov::op::Op(args) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        OPENVINO_ASSERT((get_input_size() - 1) % 3 == 0);
        for (size_t k = 0; k < 6; ++k) set_output_type(k, ov::element::i32, ov::PartialShape{ov::Dimension()});
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override { return std::make_shared<CombineSegments>(in); }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        const size_t num = (in.size() - 1) / 3;
        OPENVINO_ASSERT(num == in.back().get_size());
        std::vector<b200tok_ragged_i32> segs(num);
        size_t rows = 0, flat = 0;
        ov::Shape ps;
        for (size_t j = 0; j < num; ++j) {
            OPENVINO_ASSERT(in[3 * j + 2].get_element_type() == ov::element::i32, "the B200 path combines i32 segments");
            segs[j] = b200tok_ragged_i32{in[3 * j].data<const int32_t>(), in[3 * j + 1].data<const int32_t>(), (int64_t)in[3 * j].get_size(),
                                         in[3 * j + 2].data<const int32_t>(), (int64_t)in[3 * j + 2].get_size()};
            if (in[3 * j].get_size() >= rows) { rows = in[3 * j].get_size(); ps = in[3 * j].get_shape(); }
        }
        for (size_t j = 0; j < num; ++j) flat += (segs[j].n == 1 ? rows : 1) * (size_t)segs[j].n_elems;      // :65-71
        for (int t = 0; t < 2; ++t) { out[3 * t].set_shape(ps); out[3 * t + 1].set_shape(ps); out[3 * t + 2].set_shape({flat}); }
        int64_t n_out = 0;
        check(b200tok_combine_segments_run(device(), segs.data(), (int)num, in.back().data<const int32_t>(), out[0].data<int32_t>(), out[1].data<int32_t>(),
                                           out[2].data<int32_t>(), out[5].data<int32_t>(), (int64_t)flat, &n_out, B200TOK_MEM_HOST, nullptr));
        std::copy_n(out[0].data<int32_t>(), rows, out[3].data<int32_t>());                                   // both ragged outputs share the offsets (:30-32)
        std::copy_n(out[1].data<int32_t>(), rows, out[4].data<int32_t>());
        out[2].set_shape({(size_t)n_out}); out[5].set_shape({(size_t)n_out});                                // :127-128
        return true;
    }
};

// ---- RaggedToDense (src/ragged_to_dense.cpp:70-174), i32 elements, no trailing dense dimensions -------------------------------
class RaggedToDense : public ov::op::Op {
public:
    OPENVINO_OP("RaggedToDense");
    RaggedToDense() = default;
    RaggedToDense(const ov::OutputVector& args, bool pad_right = true, bool pad_max_length = false)
        : ov::op::Op(args), m_pad_right(pad_right), m_pad_max_length(pad_max_length) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        auto shape = get_input_partial_shape(0);
        shape.push_back(ov::Dimension());
        set_output_type(0, get_input_element_type(2), shape);
        set_output_type(1, ov::element::boolean, shape);
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override { return std::make_shared<RaggedToDense>(in, m_pad_right, m_pad_max_length); }
    bool visit_attributes(ov::AttributeVisitor& v) override { v.on_attribute("pad_right", m_pad_right); v.on_attribute("m_pad_max_length", m_pad_max_length); return true; }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        OPENVINO_ASSERT(in[2].get_element_type() == ov::element::i32 && in[2].get_shape().size() == 1, "the B200 path densifies 1-D i32 data");
        const int32_t target = in[3].data<const int32_t>()[0];
        ov::Shape shape = in[0].get_shape();
        shape.push_back((size_t)target);
        out[0].set_shape(shape); out[1].set_shape(shape);
        const bool pad_right = in.size() == 6 ? in[5].data<bool>()[0] : m_pad_right;                          // :113-116
        check(b200tok_ragged_to_dense_run(device(), in[0].data<const int32_t>(), in[1].data<const int32_t>(), (int64_t)in[0].get_size(), in[2].data<const int32_t>(),
                                          (int64_t)in[2].get_size(), target, in[4].data<const int32_t>()[0], pad_right, m_pad_max_length,
                                          out[0].data<int32_t>(), out[1] ? reinterpret_cast<uint8_t*>(out[1].data<bool>()) : nullptr, B200TOK_MEM_HOST, nullptr));
        return true;
    }
private:
    bool m_pad_right = true, m_pad_max_length = false;
};

// ---- RegexNormalization (src/regex_normalization.hpp; evaluate src/regex_normalization.cpp:127-153) and
// ---- CharsMapNormalization (src/charsmap_normalization.hpp; evaluate src/charsmap_normalization.cpp:34-69) ----------------
// Both run evaluate_normalization_helper (src/utils.cpp:178-234) through b200tok_normalize_run.
inline bool run_normalizer(b200tok_handle h, ov::TensorVector& out, const ov::TensorVector& in, bool has_skips, size_t expand) {
    const size_t n = in[0].get_size(), n_chars = in[2].get_size();
    out[0].set_shape(in[0].get_shape());
    out[1].set_shape(in[1].get_shape());
    if (has_skips) out[3] = in[3];                                          // src/utils.cpp:189-191
    int64_t produced = 0;
    size_t cap = expand * n_chars + 64;
    for (int attempt = 0; attempt < 2; ++attempt) {                         // a too small guess reports the size needed
        out[2].set_shape({cap});
        const int rc = b200tok_normalize_run(h, in[0].data<const int32_t>(), in[1].data<const int32_t>(), (int64_t)n, in[2].data<const uint8_t>(),
                                             (int64_t)n_chars, has_skips ? reinterpret_cast<const uint8_t*>(in[3].data<bool>()) : nullptr,
                                             out[0].data<int32_t>(), out[1].data<int32_t>(), out[2].data<uint8_t>(), (int64_t)cap, &produced,
                                             B200TOK_MEM_HOST, nullptr);
        if (rc == B200TOK_E_CAPACITY && (size_t)produced > cap) { cap = (size_t)produced; continue; }
        check(rc);
        break;
    }
    out[2].set_shape({(size_t)produced});                                   // src/utils.cpp:223
    return true;
}

class RegexNormalization : public ov::op::Op {
public:
    OPENVINO_OP("RegexNormalization");
    RegexNormalization() = default;
    RegexNormalization(const ov::OutputVector& args, bool global_replace = true) : ov::op::Op(args), m_global_replace(global_replace) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        const auto n = get_input_size();
        OPENVINO_ASSERT(n == 5 || n == 6, "supported input sizes are 5 or 6, got", n);
        set_output_type(0, ov::element::i32, get_input_partial_shape(0));
        set_output_type(1, ov::element::i32, get_input_partial_shape(0));
        set_output_type(2, ov::element::u8, ov::PartialShape{ov::Dimension()});
        if (n == 6) set_output_type(3, get_input_element_type(3), get_input_partial_shape(3));
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override {
        auto c = std::make_shared<RegexNormalization>(in, m_global_replace);
        c->m_state = m_state;
        return c;
    }
    bool visit_attributes(ov::AttributeVisitor& v) override { v.on_attribute("global_replace", m_global_replace); return true; }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        const bool has_skips = in.size() == 6;
        const auto& sp = in[3 + has_skips];
        const auto& rp = in[4 + has_skips];
        std::call_once(m_state->once, [&] {                                // src/regex_normalization.cpp:133-142 (lazy PCRE2 compile)
            check(b200tok_regexnorm_create(sp.data<const char>(), (int64_t)sp.get_size(), rp.data<const char>(), (int64_t)rp.get_size(),
                                           m_global_replace, device(), &m_state->h));   // patterns outside the single-character set throw here
        });
        return run_normalizer(m_state->h, out, in, has_skips, 2 + rp.get_size());
    }
private:
    bool m_global_replace = true;
    mutable std::shared_ptr<Handle> m_state = std::make_shared<Handle>();
};

class CharsMapNormalization : public ov::op::Op {
public:
    OPENVINO_OP("CharsMapNormalization");
    CharsMapNormalization() = default;
    CharsMapNormalization(const ov::OutputVector& args, bool add_dummy_prefix = false, bool remove_extra_whitespaces = true, bool escape_whitespaces = false,
                          bool case_fold = false, const std::string& normalization_form = "", bool nmt = false)
        : ov::op::Op(args), m_add_dummy_prefix(add_dummy_prefix), m_remove_extra_whitespaces(remove_extra_whitespaces),
          m_escape_whitespaces(escape_whitespaces), m_case_fold(case_fold), m_normalization_form(normalization_form), m_nmt(nmt) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        const auto n = get_input_size();
        OPENVINO_ASSERT(n == 3 || n == 4 || n == 5, "CharsMapNormalization supports input sizes 3, 4 or 5.");
        set_output_type(0, ov::element::i32, get_input_partial_shape(0));
        set_output_type(1, ov::element::i32, get_input_partial_shape(0));
        set_output_type(2, ov::element::u8, ov::PartialShape{ov::Dimension()});
        const bool has_skips = n == 5 || (n == 4 && get_input_element_type(3) == ov::element::boolean);
        if (has_skips) set_output_type(3, get_input_element_type(3), get_input_partial_shape(3));
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override {
        auto c = std::make_shared<CharsMapNormalization>(in, m_add_dummy_prefix, m_remove_extra_whitespaces, m_escape_whitespaces, m_case_fold, m_normalization_form, m_nmt);
        c->m_state = m_state;
        return c;
    }
    bool visit_attributes(ov::AttributeVisitor& v) override {
        v.on_attribute("add_dummy_prefix", m_add_dummy_prefix); v.on_attribute("remove_extra_whitespaces", m_remove_extra_whitespaces);
        v.on_attribute("escape_whitespaces", m_escape_whitespaces); v.on_attribute("normalization_form", m_normalization_form);
        v.on_attribute("case_fold", m_case_fold); v.on_attribute("nmt", m_nmt);
        return true;
    }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        const bool has_skips = in.size() == 5 || (!m_normalization_form.empty() && in.size() == 4);          // src/charsmap_normalization.cpp:35
        std::call_once(m_state->once, [&] {                                // :38-59
            std::string blob;
            if (!m_normalization_form.empty()) {
                blob = get_precompiled_charsmap(m_normalization_form, m_case_fold);                      // the reference's generated header (src/precompiled_charsmap.hpp), built with the extension
                OPENVINO_ASSERT(!blob.empty(), "Unsupported normalization form: `", m_normalization_form, "` with case_fold=", m_case_fold);
            } else {
                blob.assign(in[3 + has_skips].data<const char>(), in[3 + has_skips].get_size());
            }
            check(b200tok_charsmap_create(reinterpret_cast<const uint8_t*>(blob.data()), (int64_t)blob.size(), m_add_dummy_prefix,
                                          m_remove_extra_whitespaces, m_escape_whitespaces, device(), &m_state->h));
        });
        return run_normalizer(m_state->h, out, in, has_skips, 3);
    }
private:
    bool m_add_dummy_prefix = false, m_remove_extra_whitespaces = true, m_escape_whitespaces = false, m_case_fold = false;
    std::string m_normalization_form;
    bool m_nmt = false;
    mutable std::shared_ptr<Handle> m_state = std::make_shared<Handle>();
};

// ---- byte-level shims and detokenizer tail: BytesToChars (src/bytes_to_chars.cpp:272-339), CharsToBytes (src/chars_to_bytes.cpp:10-68),
// ---- FuzeRagged (src/fuze.cpp:10-40), UTF8Validate (src/utf8_validate.cpp:13-137) ------------------------------------------------
class BytesToChars : public ov::op::Op {
public:
    OPENVINO_OP("BytesToChars");
    BytesToChars() = default;
    explicit BytesToChars(const ov::OutputVector& args) : ov::op::Op(args) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        const auto n = get_input_size();
        OPENVINO_ASSERT(n == 5 || n == 6, "supported input sizes are 5 or 6");
        for (size_t i = 0; i < 4; ++i) set_output_type(i, ov::element::i32, i < 2 ? get_input_partial_shape(0) : ov::PartialShape{ov::Dimension()});
        set_output_type(4, ov::element::u8, ov::PartialShape{ov::Dimension()});
        if (n == 6) set_output_type(5, get_input_element_type(5), get_input_partial_shape(5));
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override { return std::make_shared<BytesToChars>(in); }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        const bool has_skips = in.size() == 6;
        out[0] = in[0]; out[1] = in[1];                                      // :296-297
        out[2].set_shape(in[2].get_shape()); out[3].set_shape(in[3].get_shape());
        const size_t cap = in[4].get_size() * 2;                             // :300
        out[4].set_shape({cap});
        if (has_skips) out[5] = in[5];
        auto rin = ragged_of(in, has_skips ? reinterpret_cast<const uint8_t*>(in[5].data<bool>()) : nullptr);
        int64_t n_chars = 0;
        check(b200tok_bytes_to_chars_run(device(), &rin, out[2].data<int32_t>(), out[3].data<int32_t>(), out[4].data<uint8_t>(), (int64_t)cap, &n_chars, nullptr));
        out[4].set_shape({(size_t)n_chars});
        return true;
    }
};

class CharsToBytes : public ov::op::Op {
public:
    OPENVINO_OP("CharsToBytes");
    CharsToBytes() = default;
    explicit CharsToBytes(const ov::OutputVector& args) : ov::op::Op(args) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        set_output_type(0, ov::element::i32, get_input_partial_shape(0));
        set_output_type(1, ov::element::i32, get_input_partial_shape(0));
        set_output_type(2, ov::element::u8, ov::PartialShape{ov::Dimension()});
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override { return std::make_shared<CharsToBytes>(in); }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        out[0].set_shape(in[0].get_shape()); out[1].set_shape(in[1].get_shape());
        const size_t cap = in[4].get_size();                                 // :41
        out[2].set_shape({cap});
        auto rin = ragged_of(in);
        int64_t n_chars = 0;
        check(b200tok_chars_to_bytes_run(device(), &rin, out[0].data<int32_t>(), out[1].data<int32_t>(), out[2].data<uint8_t>(), (int64_t)cap, &n_chars, nullptr));
        out[2].set_shape({(size_t)n_chars});
        return true;
    }
};

class FuzeRagged : public ov::op::Op {
public:
    OPENVINO_OP("FuzeRagged");
    FuzeRagged() = default;
    explicit FuzeRagged(const ov::OutputVector& args) : ov::op::Op(args) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        for (size_t i = 0; i < 4; ++i) OPENVINO_ASSERT(get_input_element_type(i) == ov::element::i32, "Expected i32 tensors as the decomposed ragged string representation");
        set_output_type(0, ov::element::i32, get_input_partial_shape(0));
        set_output_type(1, ov::element::i32, get_input_partial_shape(0));
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override { return std::make_shared<FuzeRagged>(in); }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        out[0].set_shape(in[0].get_shape()); out[1].set_shape(in[1].get_shape());
        check(b200tok_fuze_ragged_run(device(), in[0].data<const int32_t>(), in[1].data<const int32_t>(), (int64_t)in[0].get_size(), in[2].data<const int32_t>(),
                                      in[3].data<const int32_t>(), (int64_t)in[2].get_size(), out[0].data<int32_t>(), out[1].data<int32_t>(), B200TOK_MEM_HOST, nullptr));
        return true;
    }
};

class UTF8Validate : public ov::op::Op {
public:
    OPENVINO_OP("UTF8Validate");
    UTF8Validate() = default;
    UTF8Validate(const ov::OutputVector& args, bool replace_mode = false) : ov::op::Op(args), m_replace_mode(replace_mode) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        set_output_type(0, ov::element::i32, get_input_partial_shape(0));
        set_output_type(1, ov::element::i32, get_input_partial_shape(0));
        set_output_type(2, ov::element::u8, ov::PartialShape{ov::Dimension()});
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override { return std::make_shared<UTF8Validate>(in, m_replace_mode); }
    bool visit_attributes(ov::AttributeVisitor& v) override { v.on_attribute("replace_mode", m_replace_mode); return true; }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override {
        out[0].set_shape(in[0].get_shape()); out[1].set_shape(in[0].get_shape());
        const size_t cap = in[2].get_size() * 3;                             // :29-35 — the reference leaves the chars tensor at this size
        out[2].set_shape({cap});
        int64_t n_chars = 0;
        check(b200tok_utf8_validate_run(device(), in[0].data<const int32_t>(), in[1].data<const int32_t>(), (int64_t)in[0].get_size(), in[2].data<const uint8_t>(),
                                        (int64_t)in[2].get_size(), m_replace_mode, out[0].data<int32_t>(), out[1].data<int32_t>(), out[2].data<uint8_t>(), (int64_t)cap,
                                        &n_chars, B200TOK_MEM_HOST, nullptr));
        return true;
    }
private:
    bool m_replace_mode = false;
};

// ---- fused layers created at IR-load time (see the header comment) -------------------------------------------------------------------
// B200SplitBPE: inputs = RegexSplit's ragged strings [0..4] (+ skips [5]) followed by BPETokenizer's constants (its inputs [5..]);
// outputs = BPETokenizer's.  The two wrapped ops keep their own lazily built device state, shared with clones like everywhere else.
class B200SplitBPE : public ov::op::Op {
public:
    OPENVINO_OP("B200SplitBPE");
    B200SplitBPE() = default;
    B200SplitBPE(const ov::OutputVector& args, bool has_skips, std::shared_ptr<RegexSplit> split, std::shared_ptr<BPETokenizer> bpe, std::string pattern)
        : ov::op::Op(args), m_has_skips(has_skips), m_split(std::move(split)), m_bpe(std::move(bpe)), m_pattern(std::move(pattern)) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        set_output_type(0, ov::element::i32, get_input_partial_shape(0));
        set_output_type(1, ov::element::i32, get_input_partial_shape(0));
        set_output_type(2, ov::element::i32, ov::PartialShape{ov::Dimension()});
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override { return std::make_shared<B200SplitBPE>(in, m_has_skips, m_split, m_bpe, m_pattern); }
    // a model that was fused at load time can be serialised and read back: the layer carries both wrapped ops' attributes
    bool visit_attributes(ov::AttributeVisitor& v) override {
        if (!m_split) m_split = std::make_shared<RegexSplit>();
        if (!m_bpe) m_bpe = std::make_shared<BPETokenizer>();
        v.on_attribute("has_skips", m_has_skips);
        v.on_attribute("split_pattern", m_pattern);
        return m_split->visit_prefixed(v, "split_") && m_bpe->visit_attributes(v);
    }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override;
private:
    bool m_has_skips = false;
    std::shared_ptr<RegexSplit> m_split;
    std::shared_ptr<BPETokenizer> m_bpe;
    std::string m_pattern;
};
// B200SplitWordpiece: inputs = the first RegexSplit's ragged strings [0..4] followed by WordpieceTokenizer's vocab [5..7] and unk id [8].
class B200SplitWordpiece : public ov::op::Op {
public:
    OPENVINO_OP("B200SplitWordpiece");
    B200SplitWordpiece() = default;
    B200SplitWordpiece(const ov::OutputVector& args, std::shared_ptr<RegexSplit> s1, std::shared_ptr<RegexSplit> s2, std::shared_ptr<WordpieceTokenizer> wp, std::string p1, std::string p2)
        : ov::op::Op(args), m_s1(std::move(s1)), m_s2(std::move(s2)), m_wp(std::move(wp)), m_p1(std::move(p1)), m_p2(std::move(p2)) { constructor_validate_and_infer_types(); }
    void validate_and_infer_types() override {
        set_output_type(0, ov::element::i32, get_input_partial_shape(0));
        set_output_type(1, ov::element::i32, get_input_partial_shape(0));
        set_output_type(2, ov::element::i32, ov::PartialShape{ov::Dimension()});
    }
    std::shared_ptr<ov::Node> clone_with_new_inputs(const ov::OutputVector& in) const override { return std::make_shared<B200SplitWordpiece>(in, m_s1, m_s2, m_wp, m_p1, m_p2); }
    bool visit_attributes(ov::AttributeVisitor& v) override {
        bool two = (bool)m_s2;
        v.on_attribute("two_splitters", two);
        if (!m_s1) m_s1 = std::make_shared<RegexSplit>();
        if (two && !m_s2) m_s2 = std::make_shared<RegexSplit>();
        if (!m_wp) m_wp = std::make_shared<WordpieceTokenizer>();
        v.on_attribute("split_pattern", m_p1);
        bool ok = m_s1->visit_prefixed(v, "split_");
        if (two) { v.on_attribute("split2_pattern", m_p2); ok = ok && m_s2->visit_prefixed(v, "split2_"); }
        return ok && m_wp->visit_attributes(v);
    }
    bool has_evaluate() const override { return true; }
    bool evaluate(ov::TensorVector& out, const ov::TensorVector& in) const override;
private:
    std::shared_ptr<RegexSplit> m_s1, m_s2;
    std::shared_ptr<WordpieceTokenizer> m_wp;
    std::string m_p1, m_p2;
};

bool B200SplitBPE::evaluate(ov::TensorVector& out, const ov::TensorVector& in) const {
    const size_t k = 5 + (m_has_skips ? 1 : 0);
    m_split->ensure(m_pattern.data(), m_pattern.size());
    ov::TensorVector bin(in.begin(), in.begin() + 5);                     // the tokenizer's own input list: strings + its constants
    bin.insert(bin.end(), in.begin() + k, in.end());
    m_bpe->ensure(bin);
    auto rin = ragged_of(in, m_has_skips ? reinterpret_cast<const uint8_t*>(in[5].data<bool>()) : nullptr);
    ragged_ids_out(out, in, m_bpe->capacity(bin),
                   [&](b200tok_ragged_ids* r) { return b200tok_split_bpe_run(m_split->handle(), m_bpe->handle(), &rin, r, nullptr); });
    return true;
}
bool B200SplitWordpiece::evaluate(ov::TensorVector& out, const ov::TensorVector& in) const {
    const bool has_skips = in.size() == 10;                               // strings [0..4] (+ skips) + vocab [3] + unk id
    const size_t k = 5 + (has_skips ? 1 : 0);
    m_s1->ensure(m_p1.data(), m_p1.size());
    if (m_s2) m_s2->ensure(m_p2.data(), m_p2.size());
    ov::TensorVector win(in.begin(), in.begin() + 5);
    win.insert(win.end(), in.begin() + k, in.end());
    m_wp->ensure(win);
    const int32_t unk = *win[8].data<const int32_t>();
    auto rin = ragged_of(in, has_skips ? reinterpret_cast<const uint8_t*>(in[5].data<bool>()) : nullptr);
    ragged_ids_out(out, in, (int64_t)(in[4].get_size() + in[2].get_size()),
                   [&](b200tok_ragged_ids* r) { return b200tok_split_wordpiece_run(m_s1->handle(), m_s2 ? m_s2->handle() : nullptr, m_wp->handle(), &rin, unk, r, nullptr); });
    return true;
}

// ---- the load-time fusion --------------------------------------------------------------------------------------------------------------
// `outs[0..4]` are outputs 0..4 of one of this library's RegexSplit layers whose pattern is a Constant: returns it (and the pattern).
inline std::shared_ptr<RegexSplit> producing_split(const ov::OutputVector& outs, std::string& pattern) {
    if (outs.size() < 5) return nullptr;
    auto rs = std::dynamic_pointer_cast<RegexSplit>(outs[0].get_node_shared_ptr());
    if (!rs) return nullptr;
    for (size_t i = 0; i < 5; ++i)
        if (outs[i].get_node() != rs.get() || outs[i].get_index() != i) return nullptr;
    const size_t n = rs->get_input_size();
    if (n == 9) return nullptr;      // the legacy skip-token form stays a layer of its own
    auto pc = std::dynamic_pointer_cast<ov::op::v0::Constant>(rs->input_value(n - 1).get_node_shared_ptr());
    if (!pc || (rs->behaviour() != "isolate" && rs->behaviour() != "remove") || rs->max_splits() != -1) return nullptr;
    pattern.assign(static_cast<const char*>(pc->get_data_ptr()), pc->get_byte_size());
    return rs;
}
inline bool fusion_enabled() {
    static const bool on = [] { const char* e = std::getenv("B200TOK_FUSE"); return !e || std::atoi(e) != 0; }();
    return on;
}

class FusingBPEExtension : public ov::OpExtension<BPETokenizer> {
public:
    ov::OutputVector create(const ov::OutputVector& inputs, ov::AttributeVisitor& visitor) const override {
        ov::OutputVector plain = ov::OpExtension<BPETokenizer>::create(inputs, visitor);      // attributes read, inputs validated
        auto bpe = std::dynamic_pointer_cast<BPETokenizer>(plain.at(0).get_node_shared_ptr());
        std::string pattern;
        auto rs = fusion_enabled() && bpe ? producing_split(inputs, pattern) : nullptr;
        if (!rs) return plain;
        const bool has_skips = rs->get_input_size() == 7;
        ov::OutputVector args;
        for (size_t i = 0; i < 5u + (has_skips ? 1u : 0u); ++i) args.push_back(rs->input_value(i));
        args.insert(args.end(), inputs.begin() + 5, inputs.end());
        return std::make_shared<B200SplitBPE>(args, has_skips, rs, bpe, pattern)->outputs();
    }
};
class FusingWordpieceExtension : public ov::OpExtension<WordpieceTokenizer> {
public:
    ov::OutputVector create(const ov::OutputVector& inputs, ov::AttributeVisitor& visitor) const override {
        ov::OutputVector plain = ov::OpExtension<WordpieceTokenizer>::create(inputs, visitor);
        auto wp = std::dynamic_pointer_cast<WordpieceTokenizer>(plain.at(0).get_node_shared_ptr());
        std::string p2, p1;
        auto s2 = fusion_enabled() && wp ? producing_split(inputs, p2) : nullptr;
        if (!s2) return plain;
        // a second splitter in front (the BERT chain: whitespace removed, then punctuation isolated)?  Its skip flags must be the ones
        // the first splitter hands on, so that one skips tensor describes the chain.
        std::shared_ptr<RegexSplit> s1;
        {
            ov::OutputVector up;
            for (size_t i = 0; i < 5; ++i) up.push_back(s2->input_value(i));
            s1 = producing_split(up, p1);
            if (s1 && (s1->get_input_size() == 7) != (s2->get_input_size() == 7)) s1 = nullptr;
            if (s1 && s2->get_input_size() == 7 && (s2->input_value(5).get_node() != s1.get() || s2->input_value(5).get_index() != 5)) s1 = nullptr;
            if (s1 && !(s1->behaviour() == "remove" && s2->behaviour() == "isolate")) s1 = nullptr;      // the pair the fused kernel implements
        }
        const auto& first = s1 ? s1 : s2;
        const bool has_skips = first->get_input_size() == 7;
        ov::OutputVector args;
        for (size_t i = 0; i < 5u + (has_skips ? 1u : 0u); ++i) args.push_back(first->input_value(i));
        args.insert(args.end(), inputs.begin() + 5, inputs.end());
        return std::make_shared<B200SplitWordpiece>(args, s1 ? s1 : s2, s1 ? s2 : nullptr, wp, s1 ? p1 : p2, s1 ? p2 : std::string())->outputs();
    }
};

}  // namespace b200

// Same registration entry point as the reference (src/ov_extension.cpp:72); only the hot-path ops are provided here —
// load the reference extension as well for the remaining ops, ours registered last so that these names resolve here.
OPENVINO_CREATE_EXTENSIONS(std::vector<ov::Extension::Ptr>({
    std::make_shared<ov::OpExtension<b200::RegexSplit>>(),
    std::make_shared<b200::FusingBPEExtension>(),
    std::make_shared<b200::FusingWordpieceExtension>(),
    std::make_shared<ov::OpExtension<b200::VocabEncoder>>(),
    std::make_shared<ov::OpExtension<b200::VocabDecoder>>(),
    std::make_shared<ov::OpExtension<b200::ByteFallback>>(),
    std::make_shared<ov::OpExtension<b200::SpecialTokensSplit>>(),
    std::make_shared<ov::OpExtension<b200::Truncate>>(),
    std::make_shared<ov::OpExtension<b200::CombineSegments>>(),
    std::make_shared<ov::OpExtension<b200::RaggedToDense>>(),
    std::make_shared<ov::OpExtension<b200::RegexNormalization>>(),
    std::make_shared<ov::OpExtension<b200::CharsMapNormalization>>(),
    std::make_shared<ov::OpExtension<b200::BytesToChars>>(),
    std::make_shared<ov::OpExtension<b200::CharsToBytes>>(),
    std::make_shared<ov::OpExtension<b200::FuzeRagged>>(),
    std::make_shared<ov::OpExtension<b200::UTF8Validate>>(),
    std::make_shared<ov::OpExtension<b200::B200SplitBPE>>(),
    std::make_shared<ov::OpExtension<b200::B200SplitWordpiece>>(),
}));

// GenAI's GGUF path dlsym()s this factory (src/tokenizers_factory.hpp:32-33); the signature is frozen.
namespace ov { namespace tokenizers {
OPENVINO_API_C(ov::OutputVector)
create_tokenizer_node(const std::string& op_type, const ov::OutputVector& inputs, const ov::AnyMap& attributes) {
    auto get = [&](const char* k, auto def) { auto it = attributes.find(k); return it == attributes.end() ? def : it->second.template as<decltype(def)>(); };
    if (op_type == "RegexSplit") return std::make_shared<b200::RegexSplit>(inputs, get("behaviour", std::string("remove")), get("invert", false), get("max_splits", -1))->outputs();
    if (op_type == "BPETokenizer") return std::make_shared<b200::BPETokenizer>(inputs, get("unk_token", std::string()), get("fuse_unk", false), get("suffix_indicator", std::string()), get("end_suffix", std::string()), get("byte_fallback", false))->outputs();
    if (op_type == "WordpieceTokenizer") return std::make_shared<b200::WordpieceTokenizer>(inputs, get("suffix_indicator", std::string("##")), get("max_bytes_per_word", 100))->outputs();
    if (op_type == "VocabEncoder") return std::make_shared<b200::VocabEncoder>(inputs)->outputs();
    if (op_type == "VocabDecoder") return std::make_shared<b200::VocabDecoder>(inputs, get("skip_tokens", std::vector<int>{}))->outputs();
    if (op_type == "ByteFallback") return std::make_shared<b200::ByteFallback>(inputs)->outputs();
    if (op_type == "SpecialTokensSplit") return std::make_shared<b200::SpecialTokensSplit>(inputs)->outputs();
    if (op_type == "Truncate") return std::make_shared<b200::Truncate>(inputs, get("m_num_inputs", 1))->outputs();
    if (op_type == "CombineSegments") return std::make_shared<b200::CombineSegments>(inputs)->outputs();
    if (op_type == "RaggedToDense") return std::make_shared<b200::RaggedToDense>(inputs, get("pad_right", true), get("m_pad_max_length", false))->outputs();
    if (op_type == "RegexNormalization") return std::make_shared<b200::RegexNormalization>(inputs, get("global_replace", true))->outputs();
    if (op_type == "BytesToChars") return std::make_shared<b200::BytesToChars>(inputs)->outputs();
    if (op_type == "CharsToBytes") return std::make_shared<b200::CharsToBytes>(inputs)->outputs();
    if (op_type == "FuzeRagged") return std::make_shared<b200::FuzeRagged>(inputs)->outputs();
    if (op_type == "UTF8Validate") return std::make_shared<b200::UTF8Validate>(inputs, get("replace_mode", false))->outputs();
    if (op_type == "CharsMapNormalization") return std::make_shared<b200::CharsMapNormalization>(inputs, get("add_dummy_prefix", false), get("remove_extra_whitespaces", true), get("escape_whitespaces", false), get("case_fold", false), get("normalization_form", std::string()), get("nmt", false))->outputs();
    OPENVINO_THROW("Unsupported operation type in the B200 hot-path extension: ", op_type);
}
}}  // namespace ov::tokenizers
#endif  // __has_include(<openvino/op/op.hpp>)
