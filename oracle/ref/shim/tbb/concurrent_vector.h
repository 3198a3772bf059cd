// Shim for common_types.h:44 (see concurrent_unordered_map.h).
#pragma once
#include <vector>
namespace tbb {
template <class T, class A = std::allocator<T>>
using concurrent_vector = std::vector<T, A>;
}
