// ref_driver.cpp — registers the REFERENCE's own op classes (compiled unmodified from /root/reference/src against the
// stand-in OpenVINO API, see oracle/Makefile target `_ref`) with the stub driver, so tests and bench.py's CPU legs can call
// the reference's evaluate() bodies themselves.  TEST INFRASTRUCTURE ONLY: the product never loads this library.
// This file contains no reference code: it only names the classes the reference declares in its headers
// (src/ov_extension.cpp:72-109 registers the same ones).
#include <stub_driver.hpp>

#include "bpe_tokenizer.hpp"
#include "byte_fallback.hpp"
#include "bytes_to_chars.hpp"
#include "chars_to_bytes.hpp"
#include "combine_segments.hpp"
#include "fuze.hpp"
#include "ragged_to_dense.hpp"
#include "regex_normalization.hpp"
#include "regex_split.hpp"
#include "special_tokens_split.hpp"
#include "truncate.hpp"
#include "utf8_validate.hpp"
#include "vocab_decoder.hpp"
#include "vocab_encoder.hpp"
#include "wordpiece_tokenizer.hpp"

// byte_fallback.cpp:39 calls sentencepiece::PieceToByte (sentencepiece 0.2.1, pinned src/CMakeLists.txt:77, not vendored):
// restated from its published behaviour — the 256 pieces "<0x00>".."<0xFF>" (uppercase hex) map to their byte, anything
// else to -1.
namespace sentencepiece {
int PieceToByte(std::string_view piece) {
    auto hex = [](char c) { return (c >= '0' && c <= '9') ? c - '0' : (c >= 'A' && c <= 'F') ? c - 'A' + 10 : -1; };
    if (piece.size() != 6 || piece[0] != '<' || piece[1] != '0' || piece[2] != 'x' || piece[5] != '>') return -1;
    const int hi = hex(piece[3]), lo = hex(piece[4]);
    return (hi < 0 || lo < 0) ? -1 : hi * 16 + lo;
}
}  // namespace sentencepiece

namespace {
template <class T>
void add(ovs::Registry& r) { r[T::get_type_info_static().name] = std::make_shared<ov::OpExtension<T>>(); }
const ovs::Registry& registry() {
    static const ovs::Registry reg = [] {
        ovs::Registry r;
        add<RegexSplit>(r); add<BPETokenizer>(r); add<WordpieceTokenizer>(r); add<VocabEncoder>(r); add<VocabDecoder>(r); add<ByteFallback>(r);
        add<SpecialTokensSplit>(r); add<Truncate>(r); add<CombineSegments>(r); add<RaggedToDense>(r);
        add<BytesToChars>(r); add<CharsToBytes>(r); add<FuzeRagged>(r); add<UTF8Validate>(r); add<RegexNormalization>(r);
        return r;
    }();
    return reg;
}
}  // namespace
OVS_DEFINE_C_API(ovref, registry())
