// Stand-in for boost/unordered_map.hpp: the reference uses it as a plain node-based hash map
// (state_space.h:77-79).  TEST INFRASTRUCTURE.
#ifndef MPLX_BOOST_UMAP_SHIM
#define MPLX_BOOST_UMAP_SHIM
#include <boost/functional/hash.hpp>
#include <unordered_map>
namespace boost {
template <typename K, typename V, typename H = boost::hash<K>, typename E = std::equal_to<K>>
using unordered_map = std::unordered_map<K, V, H, E>;
}
#endif
