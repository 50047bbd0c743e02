// boost/functional/hash.hpp stand-in — TEST INFRASTRUCTURE ONLY (oracle/shim).
// boost::hash_combine as published for Boost < 1.81 (hash<size_t> is the identity).  Only the memo
// dictionary's key order depends on it (TreeNode.h:62-65); no arithmetic does.
#pragma once
#include <cstddef>
namespace boost {
template <class T>
inline void hash_combine(std::size_t& seed, const T& v) {
  seed ^= static_cast<std::size_t>(v) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
}
}  // namespace boost
