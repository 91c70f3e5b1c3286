// stand-in for absl::string_view (test infrastructure): abseil aliases it to std::string_view under C++17 as well
#pragma once
#include <string_view>
namespace absl { using string_view = std::string_view; }
