// boost/range.hpp stand-in — TEST INFRASTRUCTURE ONLY (oracle/shim); see zip_iterator.hpp.
#pragma once
namespace boost {
template <class It> struct iterator_range { It b, e; };
template <class It> iterator_range<It> make_iterator_range(It b, It e) { return iterator_range<It>{b, e}; }
}  // namespace boost
