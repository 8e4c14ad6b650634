// Structural check of oracle/rstar_port.hpp (test infrastructure): after every 50th of 6000 random insertions a 50-nearest query
// must return exactly the 50 smallest distances in non-decreasing order, every point must be stored once, and every node but
// the root must hold 3..6 children (rstar's default parameters).  Built and run by tests/test_oracle_aux.py.
#include "rstar_port.hpp"
#include <cstdio>
#include <random>
int main() {
    std::mt19937 rng(1);
    rstar_port::RTree t;
    std::vector<std::pair<int,int>> pts;
    std::vector<char> used(100*100, 0);
    int bad = 0;
    for (int it = 0; it < 6000; ++it) {
        int x, y;
        do { x = rng() % 100; y = rng() % 100; } while (used[y*100+x]);
        used[y*100+x] = 1;
        t.insert(x, y); pts.push_back({x, y});
        if (it % 50 == 0) {
            int qx = rng() % 100, qy = rng() % 100;
            std::vector<int> out; t.nearest(qx, qy, 50, out);
            std::vector<int> d; for (auto& p : pts) d.push_back((p.first-qx)*(p.first-qx)+(p.second-qy)*(p.second-qy));
            std::sort(d.begin(), d.end());
            size_t n = std::min<size_t>(50, pts.size());
            if (out.size()/2 != n) { ++bad; continue; }
            for (size_t i = 0; i < n; ++i) { int dd = (out[2*i]-qx)*(out[2*i]-qx)+(out[2*i+1]-qy)*(out[2*i+1]-qy); if (dd != d[i]) { ++bad; break; } }
        }
    }
    // count nodes / check child counts
    size_t leaves = 0, maxc = 0, minc = 99; int depth = 0;
    std::vector<std::pair<const rstar_port::Node*, int>> st{{t.root.get(), 0}};
    while (!st.empty()) { auto [n, d] = st.back(); st.pop_back(); if (n->leaf) { ++leaves; depth = std::max(depth, d); continue; } if (n != t.root.get()) { maxc = std::max(maxc, n->ch.size()); minc = std::min(minc, n->ch.size()); } for (auto& c : n->ch) st.push_back({c.get(), d+1}); }
    printf("bad queries %d, leaves %zu (inserted %zu), child counts %zu..%zu, depth %d\n", bad, leaves, pts.size(), minc, maxc, depth);
    return (bad == 0 && leaves == pts.size() && minc >= 3 && maxc <= 6) ? 0 : 1;
}
