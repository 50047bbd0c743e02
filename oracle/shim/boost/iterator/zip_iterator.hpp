// boost/iterator/zip_iterator.hpp stand-in — TEST INFRASTRUCTURE ONLY (oracle/shim).
// Only named by the never-instantiated `zip` template in helper.h:9-16.
#pragma once
#include <tuple>
namespace boost {
template <class T> struct zip_iterator { T t; };
template <class... A> std::tuple<A...> make_tuple(A... a) { return std::tuple<A...>(a...); }
template <class T> zip_iterator<T> make_zip_iterator(T t) { return zip_iterator<T>{t}; }
}  // namespace boost
