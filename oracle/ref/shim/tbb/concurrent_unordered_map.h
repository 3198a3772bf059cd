// Shim: the reference's include/visnav/common_types.h:43 includes TBB, which is
// not installed here.  The BA path never uses concurrency on these containers,
// so alias them to the std equivalents (SURVEY.md §8(c) gotcha 2).  TBB's
// default hasher casts the key to size_t, which FrameCamId supports
// (common_types.h:93-98); std::hash<FrameCamId> is only specialised later in
// that header (:344), so it cannot be the default here.
#pragma once
#include <cstddef>
#include <functional>
#include <unordered_map>
namespace tbb {
template <class K>
struct tbb_hash {
  std::size_t operator()(const K& k) const { return static_cast<std::size_t>(k); }
};
template <class K, class V, class H = tbb_hash<K>, class E = std::equal_to<K>,
          class A = std::allocator<std::pair<const K, V>>>
using concurrent_unordered_map = std::unordered_map<K, V, H, E, A>;
}  // namespace tbb
