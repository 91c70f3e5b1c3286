// stand-in for absl::flat_hash_map (test infrastructure): same interface subset on std::unordered_map — iteration order
// is never observable in the ops that use it (lookups only)
#pragma once
#include <unordered_map>
namespace absl {
template <class K, class V, class H = std::hash<K>, class E = std::equal_to<K>>
using flat_hash_map = std::unordered_map<K, V, H, E>;
}
