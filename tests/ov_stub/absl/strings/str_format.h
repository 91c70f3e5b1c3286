// stand-in for absl::StrFormat (test infrastructure): printf-style formatting into a std::string, which is what the
// reference uses it for ("<0x%02X>", src/bpe_tokenizer.cpp:244)
#pragma once
#include <cstdio>
#include <string>
namespace absl {
template <class... A>
std::string StrFormat(const char* fmt, A... a) {
    char buf[256];
    const int n = std::snprintf(buf, sizeof(buf), fmt, a...);
    return std::string(buf, n < 0 ? 0 : (size_t)(n < (int)sizeof(buf) ? n : (int)sizeof(buf) - 1));
}
}  // namespace absl
