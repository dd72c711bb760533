// Stand-in for boost/heap/d_ary_heap.hpp (Boost is absent from this image).  TEST INFRASTRUCTURE.
// Written from the documented semantics of boost::heap::d_ary_heap<T, mutable_<true>, arity<D>,
// compare<Cmp>>: an array D-ary max-heap w.r.t. Cmp whose elements keep stable handles;
// push appends and sifts up, pop swaps the root with the last element and sifts down,
// increase(handle) sifts up, erase(handle) moves the element to the root and pops; a child replaces its
// parent unless it compares strictly lower, and among equal children the first one is taken.
// Tie-breaking between fully equal keys may differ from a given Boost release (unpinned in the
// reference); the host planner of this repository implements the same rules.
#ifndef MPLX_BOOST_DARY_SHIM
#define MPLX_BOOST_DARY_SHIM
#include <cstddef>
#include <list>
#include <utility>
#include <vector>

namespace boost {
namespace heap {
template <bool B>
struct mutable_ {};
template <int N>
struct arity {
  static const int value = N;
};
template <typename C>
struct compare {
  typedef C type;
};

namespace detail {
template <typename... Opts>
struct find_compare;
template <typename C, typename... Rest>
struct find_compare<compare<C>, Rest...> {
  typedef C type;
};
template <typename O, typename... Rest>
struct find_compare<O, Rest...> : find_compare<Rest...> {};
template <typename... Opts>
struct find_arity;
template <int N, typename... Rest>
struct find_arity<arity<N>, Rest...> {
  static const int value = N;
};
template <typename O, typename... Rest>
struct find_arity<O, Rest...> : find_arity<Rest...> {};
}  // namespace detail

template <typename T, typename... Opts>
class d_ary_heap {
  struct Node {
    T value;
    std::size_t index;
  };
  typedef typename std::list<Node>::iterator It;
  typedef typename detail::find_compare<Opts...>::type Cmp;
  static const int D = detail::find_arity<Opts...>::value;

 public:
  class handle_type {
   public:
    handle_type() : valid_(false) {}
    T &operator*() const { return it_->value; }

   private:
    friend class d_ary_heap;
    explicit handle_type(It it) : it_(it), valid_(true) {}
    It it_;
    bool valid_;
  };

  bool empty() const { return q_.empty(); }
  std::size_t size() const { return q_.size(); }
  const T &top() const { return q_.front()->value; }
  void clear() {
    q_.clear();
    nodes_.clear();
  }
  handle_type push(const T &v) {
    nodes_.push_back(Node{v, q_.size()});
    It it = --nodes_.end();
    q_.push_back(it);
    siftup(it->index);
    return handle_type(it);
  }
  void pop() { erase_at(0); }
  void increase(handle_type h) { siftup(h.it_->index); }
  void decrease(handle_type h) { siftdown(h.it_->index); }
  void update(handle_type h) {
    const std::size_t i = h.it_->index;
    if (i > 0 && cmp_(q_[(i - 1) / D]->value, q_[i]->value))
      siftup(i);
    else
      siftdown(i);
  }
  // erase = move the element to the root without comparisons, then pop (as boost's d_ary_heap does)
  void erase(handle_type h) {
    std::size_t i = h.it_->index;
    while (i != 0) {
      const std::size_t p = (i - 1) / D;
      swap_at(p, i);
      i = p;
    }
    pop();
  }

 private:
  void swap_at(std::size_t a, std::size_t b) {
    std::swap(q_[a], q_[b]);
    q_[a]->index = a;
    q_[b]->index = b;
  }
  void erase_at(std::size_t i) {
    const std::size_t last = q_.size() - 1;
    It victim = q_[i];
    if (i != last) swap_at(i, last);
    q_.pop_back();
    nodes_.erase(victim);
    if (i < q_.size()) {
      if (i > 0 && cmp_(q_[(i - 1) / D]->value, q_[i]->value))
        siftup(i);
      else
        siftdown(i);
    }
  }
  void siftup(std::size_t i) {
    while (i != 0) {
      const std::size_t p = (i - 1) / D;
      if (cmp_(q_[p]->value, q_[i]->value)) {
        swap_at(p, i);
        i = p;
      } else
        return;
    }
  }
  void siftdown(std::size_t i) {
    const std::size_t n = q_.size();
    while (D * i + 1 < n) {
      std::size_t c = D * i + 1;
      const std::size_t end = c + D < n ? c + D : n;
      for (std::size_t k = c + 1; k < end; k++)
        if (cmp_(q_[c]->value, q_[k]->value)) c = k;  // first maximum among the children
      if (!cmp_(q_[c]->value, q_[i]->value)) {
        swap_at(c, i);
        i = c;
      } else
        return;
    }
  }
  std::list<Node> nodes_;
  std::vector<It> q_;
  Cmp cmp_;
};
}  // namespace heap
}  // namespace boost
#endif
