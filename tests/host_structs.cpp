// Unit checks of the host planner's own containers (mpl_host.hpp) against simple references:
//   KeyMap          vs std::unordered_map (random inserts / look-ups, growth, swap)
//   SmallVec        vs std::vector
//   PriorityQueue   push / pop / increase / erase vs a brute-force "extract best" with the
//                   reference's compare_pair ordering (state_space.h:16-27): the popped KEY sequence
//                   must be non-decreasing and every state must come out exactly once.
// Compiled and run by tests/test_host_structs_cpu.py (needs only the header; libmplx is not called).
#include <cstdio>
#include <random>
#include <unordered_map>

#include "../motion_primitive_library_b200/host/mpl_host.hpp"

static int fails = 0;
#define CHECK(c)                                              \
  do {                                                        \
    if (!(c)) {                                               \
      std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); \
      fails++;                                                \
    }                                                         \
  } while (0)

int main() {
  std::mt19937_64 rng(7);
  {  // KeyMap
    MPL::KeyMap<int> km, other;
    std::unordered_map<std::size_t, int *> ref;
    std::vector<int> store(200000);
    for (int i = 0; i < 200000; i++) {
      // clustered keys too: the slot function must spread them
      const std::size_t k = (i % 3 == 0) ? (std::size_t)i * 4096 : rng();
      bool created;
      int *&slot = km.obtain(k, created);
      CHECK(created == (ref.find(k) == ref.end()));
      if (created) {
        slot = &store[i];
        ref[k] = &store[i];
      }
      CHECK(slot == ref[k]);
    }
    CHECK(km.size() == ref.size());
    for (const auto &e : ref) CHECK(km.find(e.first) == e.second);
    for (int i = 0; i < 1000; i++) {
      const std::size_t k = rng() | 1;
      if (!ref.count(k)) CHECK(km.find(k) == nullptr);
    }
    km.swap(other);
    CHECK(km.size() == 0 && other.size() == ref.size() && km.find(ref.begin()->first) == nullptr);
    CHECK(other.find(ref.begin()->first) == ref.begin()->second);
  }
  {  // SmallVec
    struct R { void *p; double c; int a; };
    MPL::SmallVec<R, 1> sv;
    std::vector<R> ref;
    CHECK(sv.empty());
    for (int i = 0; i < 100; i++) {
      R r{(void *)(std::size_t)i, i * 0.5, i};
      sv.push_back(r);
      ref.push_back(r);
      CHECK(sv.size() == ref.size());
      for (std::size_t k = 0; k < ref.size(); k++) CHECK(sv[k].a == ref[k].a && sv[k].c == ref[k].c && sv[k].p == ref[k].p);
    }
    int n = 0;
    for (const auto &r : sv) CHECK(r.a == n++);
    sv.clear();
    CHECK(sv.empty() && sv.size() == 0);
    sv.push_back(R{nullptr, 1.0, 9});
    CHECK(sv.size() == 1 && sv[0].a == 9);
  }
  {  // PriorityQueue
    using S = MPL::State<2>;
    const int N = 3000;
    std::deque<S> states;
    for (int i = 0; i < N; i++) states.emplace_back(Waypoint<2>(), (std::size_t)i);
    MPL::PriorityQueue<2> pq;
    std::vector<double> key(N, 0);
    std::vector<char> in(N, 0);
    std::uniform_int_distribution<int> coarse(0, 40);  // many ties, as with lattice costs
    int live = 0;
    double last = -1;
    for (int step = 0; step < 40000; step++) {
      const int op = (int)(rng() % 10);
      const int i = (int)(rng() % N);
      if (op < 5 && !in[i]) {
        key[i] = coarse(rng);
        states[i].g = states[i].rhs = (double)coarse(rng);
        pq.push(key[i], &states[i]);
        in[i] = 1;
        live++;
      } else if (op < 7 && in[i] && key[i] > 0) {  // increase priority = smaller key
        key[i] -= 1;
        pq.increase(&states[i], key[i]);
      } else if (op < 8 && in[i]) {
        pq.erase(&states[i]);
        in[i] = 0;
        live--;
        CHECK(states[i].heap_idx == -1);
      } else if (op >= 8 && !pq.empty()) {
        // the top must be a best element under compare_pair
        const auto top = pq.top();
        for (int k = 0; k < N; k++)
          if (in[k]) CHECK(!MPL::PriorityQueue<2>::lower(top, std::make_pair(key[k], &states[k])));
        const int t = (int)top.second->key;
        CHECK(in[t] && top.first == key[t]);
        pq.pop();
        in[t] = 0;
        live--;
      }
      CHECK((int)pq.size() == live);
      for (std::size_t q = 0; q < pq.raw().size() && q < 4; q++) CHECK(pq.raw()[q].second->heap_idx == (int)q);
    }
    while (!pq.empty()) {  // drain: keys come out in non-decreasing order
      CHECK(pq.top().first >= last);
      last = pq.top().first;
      in[pq.top().second->key] = 0;
      pq.pop();
    }
    for (int k = 0; k < N; k++) CHECK(!in[k]);
  }
  std::printf("host_structs fails %d\n", fails);
  return fails ? 1 : 0;
}
