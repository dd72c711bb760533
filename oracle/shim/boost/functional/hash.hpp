// Stand-in for boost/functional/hash.hpp (Boost is absent from this image; the reference
// does not pin a version).  hash_combine for a 64-bit std::size_t as shipped by Boost
// 1.56-1.80 (hash_combine_impl(uint64_t&, uint64_t), a MurmurHash2-style mix);
// boost::hash<int> is the sign-extending cast.  TEST INFRASTRUCTURE.
#ifndef MPLX_BOOST_HASH_SHIM
#define MPLX_BOOST_HASH_SHIM
#include <cstddef>
#include <cstdint>
namespace boost {
// boost::hash<T> for a user type calls hash_value(v) found by argument-dependent lookup
// (waypoint.h:93 defines hash_value(Waypoint<Dim>)).
template <typename T>
struct hash {
  std::size_t operator()(const T &v) const { return hash_value(v); }
};
template <>
struct hash<int> {
  std::size_t operator()(int v) const { return static_cast<std::size_t>(v); }
};
inline void hash_combine(std::size_t &seed, int v) {
  std::uint64_t k = hash<int>()(v);
  const std::uint64_t m = 0xc6a4a7935bd1e995ULL;
  const int r = 47;
  k *= m;
  k ^= k >> r;
  k *= m;
  seed ^= k;
  seed *= m;
  seed += 0xe6546b64;
}
}  // namespace boost
#endif
