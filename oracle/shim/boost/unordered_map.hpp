// Stand-in for boost/unordered_map.hpp: the reference uses it as a node-based hash map from
// Waypoint to StatePtr (state_space.h:77-79).  TEST INFRASTRUCTURE.
//
// Iteration order of boost::unordered_map is unspecified (and Boost is unpinned in the reference);
// two reference loops depend on it — getSubStateSpace's re-queueing of the open states
// (state_space.h:184-192) and getLinkedNodes (map_planner.cpp:128) — so this stand-in fixes it to
// INSERTION order, the order the host planner of this repository defines for the same loops.
// Everything else (lookup, operator[] default-inserting, copy) is plain hash-map behaviour.
#ifndef MPLX_BOOST_UMAP_SHIM
#define MPLX_BOOST_UMAP_SHIM
#include <boost/functional/hash.hpp>
#include <list>
#include <unordered_map>
#include <utility>
namespace boost {
template <typename K, typename V, typename H = boost::hash<K>, typename E = std::equal_to<K>>
class unordered_map {
 public:
  typedef K key_type;
  typedef V mapped_type;
  typedef std::pair<const K, V> value_type;
  typedef typename std::list<value_type>::iterator iterator;
  typedef typename std::list<value_type>::const_iterator const_iterator;

  unordered_map() {}
  unordered_map(const unordered_map &o) : items_(o.items_) { reindex(); }
  unordered_map &operator=(const unordered_map &o) {
    if (this != &o) {
      items_ = std::list<value_type>(o.items_.begin(), o.items_.end());
      reindex();
    }
    return *this;
  }
  iterator begin() { return items_.begin(); }
  iterator end() { return items_.end(); }
  const_iterator begin() const { return items_.begin(); }
  const_iterator end() const { return items_.end(); }
  std::size_t size() const { return items_.size(); }
  bool empty() const { return items_.empty(); }
  void clear() {
    items_.clear();
    idx_.clear();
  }
  V &operator[](const K &k) {
    auto it = idx_.find(k);
    if (it != idx_.end()) return it->second->second;
    items_.push_back(value_type(k, V()));
    iterator pos = --items_.end();
    idx_.emplace(k, pos);
    return pos->second;
  }
  iterator find(const K &k) {
    auto it = idx_.find(k);
    return it == idx_.end() ? items_.end() : it->second;
  }
  const_iterator find(const K &k) const {
    auto it = idx_.find(k);
    return it == idx_.end() ? items_.end() : const_iterator(it->second);
  }
  std::size_t count(const K &k) const { return idx_.count(k); }

 private:
  void reindex() {
    idx_.clear();
    for (iterator it = items_.begin(); it != items_.end(); ++it) idx_.emplace(it->first, it);
  }
  std::list<value_type> items_;
  std::unordered_map<K, iterator, H, E> idx_;
};
}  // namespace boost
#endif
