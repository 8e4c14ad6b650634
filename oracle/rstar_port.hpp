// rstar_port.hpp -- TEST INFRASTRUCTURE ONLY (selected with ORC_KNN=rstar): a restatement of the parts of rstar 0.7.1 the
// reference touches (RTree<[i32; 2]>::new / insert / nearest_neighbor_iter at ms.rs:1321-1531, default parameters MIN 3 / MAX 6 /
// REINSERTION 2, R* insertion strategy) plus the std BinaryHeap that orders its nearest-neighbour iterator.
//
// Why: the order in which rstar yields EQUIDISTANT points depends on the shape of the tree and on the heap, and which of them make
// the k-cut changes the candidate set.  The crate source is not available offline; this is written from its published algorithm
// (choose_subtree by inclusion / overlap increase / area increase, forced reinsertion of the two children farthest from the node
// centre once per level, split along the axis of minimum perimeter sum at the position of minimum overlap then area).  Five details
// I could not recall with certainty are switchable through ORC_RSTAR_VARIANT (see variant()); all 32 combinations were run against
// the nine hash constants of lib/tests/diff.rs; the default is the one that reproduces six of them exactly on Pillow decodes
// (0 0 1 0 16 0 9 0 0) and ALL NINE on inputs decoded as jpeg-decoder 0.1.22 does (oracle/jpeg_port.py;
// tests/test_oracle_pin.py::test_oracle_reproduces_every_reference_hash).
// The CUDA path and its parity tests use the oracle's canonical order, not this one.
#pragma once
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstdint>
#include <memory>
#include <utility>
#include <vector>

namespace rstar_port {

struct AABB {
    int lo[2], hi[2];
    static AABB empty() { return AABB{{INT_MAX, INT_MAX}, {INT_MIN, INT_MIN}}; }
    static AABB point(int x, int y) { return AABB{{x, y}, {x, y}}; }
    void merge(const AABB& o) {
        for (int a = 0; a < 2; ++a) { lo[a] = std::min(lo[a], o.lo[a]); hi[a] = std::max(hi[a], o.hi[a]); }
    }
    bool contains(const AABB& o) const {
        return lo[0] <= o.lo[0] && lo[1] <= o.lo[1] && hi[0] >= o.hi[0] && hi[1] >= o.hi[1];
    }
    // wrapping i32 arithmetic like release-mode Rust
    static int wsub(int a, int b) { return (int)((unsigned)a - (unsigned)b); }
    static int wadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }
    static int wmul(int a, int b) { return (int)((unsigned)a * (unsigned)b); }
    int area() const {
        int acc = 1;
        for (int a = 0; a < 2; ++a) acc = wmul(std::max(wsub(hi[a], lo[a]), 0), acc);
        return acc;
    }
    int perimeter_value() const {
        int s = 0;
        for (int a = 0; a < 2; ++a) s = wadd(s, wsub(hi[a], lo[a]));
        return std::max(s, 0);
    }
    int intersection_area(const AABB& o) const {
        AABB r;
        for (int a = 0; a < 2; ++a) { r.lo[a] = std::max(lo[a], o.lo[a]); r.hi[a] = std::min(hi[a], o.hi[a]); }
        return r.area();
    }
    void center(int c[2]) const { for (int a = 0; a < 2; ++a) c[a] = wadd(lo[a], hi[a]) / 2; }
    int distance_2(int x, int y) const {
        const int p[2] = {x, y};
        int d = 0;
        for (int a = 0; a < 2; ++a) {
            const int q = std::min(hi[a], std::max(lo[a], p[a]));  // min_point: the point of the box closest to p
            d = wadd(d, wmul(wsub(q, p[a]), wsub(q, p[a])));
        }
        return d;
    }
};

struct Node {
    bool leaf = false;
    int pt[2] = {0, 0};                         // leaf
    AABB env = AABB::empty();                   // parent
    std::vector<std::unique_ptr<Node>> ch;      // parent
    AABB envelope() const { return leaf ? AABB::point(pt[0], pt[1]) : env; }
};
using NodePtr = std::unique_ptr<Node>;

constexpr size_t MIN_SIZE = 3, MAX_SIZE = 6, REINSERTION_COUNT = 2;
// Details this restatement was unsure about can be flipped for measurement (ORC_RSTAR_VARIANT bit mask).  The default (0) is the
// combination that reproduces the reference's golden hashes: of the 32 combinations scanned, it alone brings six of the nine
// diff.rs constants to distance 0 (all three PNG-only ones) -- see tests/test_oracle_pin.py and DESIGN.md section 2.
// bit 0: heap extend rebuilds (newer std) | bit 1: sift-down tie goes left | bit 2: farther reinsertion candidate first |
// bit 3: split axis keeps the last perimeter of axis 0 | bit 4: overlap increase also above the leaf level
inline int variant() { static int v = getenv("ORC_RSTAR_VARIANT") ? atoi(getenv("ORC_RSTAR_VARIANT")) : 0; return v; }

inline AABB envelope_for_children(const std::vector<NodePtr>& ch) {
    AABB r = AABB::empty();
    for (auto& c : ch) r.merge(c->envelope());
    return r;
}

struct InsertionResult {
    enum Kind { Complete, Split, Reinsert } kind = Complete;
    NodePtr split;
    std::vector<NodePtr> reinsert;
    size_t height = 0;
};

inline size_t choose_subtree(const Node& node, const Node& to_insert) {
    if (node.ch.empty() || node.ch[0]->leaf) return SIZE_MAX;
    const Node& first = *node.ch[0];
    const bool all_leaves = first.ch.empty() ? true : first.ch[0]->leaf;
    const AABB ins = to_insert.envelope();
    size_t inclusion_count = 0, min_index = 0;
    int min_area = INT_MAX;
    for (size_t i = 0; i < node.ch.size(); ++i) {
        const AABB e = node.ch[i]->envelope();
        if (e.contains(ins)) {
            ++inclusion_count;
            const int area = e.area();
            if (area < min_area) { min_area = area; min_index = i; }
        }
    }
    if (inclusion_count == 0) {
        int m0 = 0, m1 = 0, m2 = 0;  // (overlap increase, area increase, area), lexicographic
        for (size_t i = 0; i < node.ch.size(); ++i) {
            const AABB e = node.ch[i]->envelope();
            AABB ne = e;
            ne.merge(ins);
            int overlap_increase = 0;
            if (all_leaves || (variant() & 16)) {
                int overlap = 0, new_overlap = 0;
                for (size_t j = 0; j < node.ch.size(); ++j) {
                    if (j == i) continue;
                    const AABB ce = node.ch[j]->envelope();
                    overlap = AABB::wadd(overlap, e.intersection_area(ce));
                    new_overlap = AABB::wadd(new_overlap, ne.intersection_area(ce));
                }
                overlap_increase = AABB::wsub(new_overlap, overlap);
            }
            const int area = ne.area();
            const int area_increase = AABB::wsub(area, e.area());
            const bool less = overlap_increase != m0 ? overlap_increase < m0 : (area_increase != m1 ? area_increase < m1 : area < m2);
            if (less || i == 0) { m0 = overlap_increase; m1 = area_increase; m2 = area; min_index = i; }
        }
    }
    return min_index;
}

inline void sort_envelopes(size_t axis, std::vector<NodePtr>& ch) {  // slice::sort_by is stable
    std::stable_sort(ch.begin(), ch.end(), [axis](const NodePtr& l, const NodePtr& r) { return l->envelope().lo[axis] < r->envelope().lo[axis]; });
}

inline size_t get_split_axis(Node& node) {
    int best_goodness = 0;
    size_t best_axis = 0;
    const size_t until = node.ch.size() - MIN_SIZE + 1;
    for (size_t axis = 0; axis < 2; ++axis) {
        sort_envelopes(axis, node.ch);
        AABB first = AABB::empty(), second = AABB::empty();
        for (size_t i = 0; i < MIN_SIZE; ++i) first.merge(node.ch[i]->envelope());
        for (size_t i = until; i < node.ch.size(); ++i) second.merge(node.ch[i]->envelope());
        for (size_t k = MIN_SIZE; k < until; ++k) {
            AABB fm = first, sm = second;
            for (size_t i = 0; i < k; ++i) fm.merge(node.ch[i]->envelope());
            for (size_t i = k; i < node.ch.size(); ++i) sm.merge(node.ch[i]->envelope());
            const int perimeter = AABB::wadd(fm.perimeter_value(), sm.perimeter_value());
            // (recalled as `|| axis == 0`, which would keep the LAST perimeter of axis 0; the reference's hashes say it is the minimum)
            const bool first_k = (variant() & 8) ? (axis == 0) : (axis == 0 && k == MIN_SIZE);
            if (best_goodness > perimeter || first_k) { best_axis = axis; best_goodness = perimeter; }
        }
    }
    return best_axis;
}

inline NodePtr split(Node& node) {
    const size_t axis = get_split_axis(node);
    sort_envelopes(axis, node.ch);
    int b0 = 0, b1 = 0;
    size_t best_index = MIN_SIZE;
    for (size_t k = MIN_SIZE; k <= node.ch.size() - MIN_SIZE; ++k) {
        AABB fe = node.ch[k - 1]->envelope(), se = node.ch[k]->envelope();
        for (size_t i = 0; i < k; ++i) fe.merge(node.ch[i]->envelope());
        for (size_t i = k; i < node.ch.size(); ++i) se.merge(node.ch[i]->envelope());
        const int overlap = fe.intersection_area(se);
        const int area = AABB::wadd(fe.area(), se.area());
        const bool less = overlap != b0 ? overlap < b0 : area < b1;
        if (less || k == MIN_SIZE) { b0 = overlap; b1 = area; best_index = k; }
    }
    NodePtr off(new Node());
    for (size_t i = best_index; i < node.ch.size(); ++i) off->ch.push_back(std::move(node.ch[i]));
    node.ch.resize(best_index);
    node.env = envelope_for_children(node.ch);
    off->env = envelope_for_children(off->ch);
    return off;
}

inline std::vector<NodePtr> get_nodes_for_reinsertion(Node& node) {
    int c[2];
    node.env.center(c);
    auto dist = [&](const NodePtr& n) {
        int nc[2];
        n->envelope().center(nc);
        return AABB::wadd(AABB::wmul(AABB::wsub(nc[0], c[0]), AABB::wsub(nc[0], c[0])), AABB::wmul(AABB::wsub(nc[1], c[1]), AABB::wsub(nc[1], c[1])));
    };
    std::stable_sort(node.ch.begin(), node.ch.end(), [&](const NodePtr& l, const NodePtr& r) { return dist(l) < dist(r); });
    std::vector<NodePtr> out;
    const size_t keep = node.ch.size() - REINSERTION_COUNT;
    for (size_t i = keep; i < node.ch.size(); ++i) out.push_back(std::move(node.ch[i]));
    node.ch.resize(keep);
    node.env = envelope_for_children(node.ch);
    return out;
}

inline InsertionResult resolve_overflow_without_reinsertion(Node& node) {
    InsertionResult r;
    if (node.ch.size() > MAX_SIZE) { r.kind = InsertionResult::Split; r.split = split(node); }
    return r;
}
inline InsertionResult resolve_overflow(Node& node, size_t current_height) {
    InsertionResult r;
    if (node.ch.size() > MAX_SIZE) {
        r.kind = InsertionResult::Reinsert;
        r.reinsert = get_nodes_for_reinsertion(node);
        r.height = current_height;
    }
    return r;
}

inline InsertionResult forced_insertion(Node& node, NodePtr t, size_t target_height) {
    node.env.merge(t->envelope());
    const size_t idx = choose_subtree(node, *t);
    if (target_height == 0 || idx >= node.ch.size()) {
        node.ch.push_back(std::move(t));
        return resolve_overflow_without_reinsertion(node);
    }
    InsertionResult sub = forced_insertion(*node.ch[idx], std::move(t), target_height - 1);
    if (sub.kind == InsertionResult::Split) {
        node.env.merge(sub.split->envelope());
        node.ch.push_back(std::move(sub.split));
        return resolve_overflow_without_reinsertion(node);
    }
    return sub;
}

inline InsertionResult recursive_insert(Node& node, NodePtr t, size_t current_height) {
    node.env.merge(t->envelope());
    const size_t idx = choose_subtree(node, *t);
    if (idx >= node.ch.size()) {
        node.ch.push_back(std::move(t));
        return resolve_overflow(node, current_height);
    }
    InsertionResult sub = recursive_insert(*node.ch[idx], std::move(t), current_height + 1);
    if (sub.kind == InsertionResult::Split) {
        node.env.merge(sub.split->envelope());
        node.ch.push_back(std::move(sub.split));
        return resolve_overflow(node, current_height);
    }
    if (sub.kind == InsertionResult::Reinsert) node.env = envelope_for_children(node.ch);
    return sub;
}

// std::collections::BinaryHeap (max-heap) of (distance, node) with the iterator's INVERTED comparison (smaller distance = greater)
struct HeapItem { int distance; const Node* node; };
struct MinHeap {
    std::vector<HeapItem> d;
    static bool le(const HeapItem& a, const HeapItem& b) { return !(a.distance < b.distance); }  // a <= b in heap order
    static bool gt(const HeapItem& a, const HeapItem& b) { return a.distance < b.distance; }     // a > b in heap order
    void sift_up(size_t start, size_t pos) {
        HeapItem e = d[pos];
        while (pos > start) {
            const size_t parent = (pos - 1) / 2;
            if (le(e, d[parent])) break;
            d[pos] = d[parent];
            pos = parent;
        }
        d[pos] = e;
    }
    void push(const HeapItem& it) { d.push_back(it); sift_up(0, d.size() - 1); }
    void sift_down(size_t pos) {  // sift_down_range(pos, len)
        const size_t end = d.size();
        HeapItem e = d[pos];
        size_t child = 2 * pos + 1;
        while (child < end) {
            const size_t right = child + 1;
            if (right < end && !gt(d[child], d[right])) child = right;
            if (!gt(d[child], e)) break;   // hole.element() >= hole.get(child)
            d[pos] = d[child];
            pos = child;
            child = 2 * pos + 1;
        }
        d[pos] = e;
    }
    bool pop(HeapItem& out) {
        if (d.empty()) return false;
        HeapItem item = d.back();
        d.pop_back();
        if (!d.empty()) {
            std::swap(item, d[0]);
            // sift_down_to_bottom(0)
            const size_t end = d.size();
            size_t pos = 0;
            HeapItem e = d[0];
            size_t child = 1;
            while (child < end) {
                const size_t right = child + 1;
                if (right < end && ((variant() & 2) ? gt(d[right], d[child]) : !gt(d[child], d[right]))) child = right;
                d[pos] = d[child];
                pos = child;
                child = 2 * pos + 1;
            }
            d[pos] = e;
            sift_up(0, pos);
        }
        out = item;
        return true;
    }
};

struct RTree {
    NodePtr root;
    RTree() : root(new Node()) {}
    void insert(int x, int y) {
        NodePtr leaf(new Node());
        leaf->leaf = true; leaf->pt[0] = x; leaf->pt[1] = y;
        InsertionResult first = recursive_insert(*root, std::move(leaf), 0);
        size_t target_height = 0;
        struct Action { bool is_split; NodePtr node; };
        std::vector<Action> stack;
        if (first.kind == InsertionResult::Split) stack.push_back({true, std::move(first.split)});
        else if (first.kind == InsertionResult::Reinsert) {
            // (recalled as "extend, then pop from the back"; the reference's hashes say the nearer of the two goes back in first)
            if (!(variant() & 4)) std::reverse(first.reinsert.begin(), first.reinsert.end());
            for (auto& n : first.reinsert) stack.push_back({false, std::move(n)});
            target_height = first.height;
        }
        while (!stack.empty()) {
            Action next = std::move(stack.back());
            stack.pop_back();
            if (next.is_split) {  // the root was split: new root one level up
                NodePtr new_root(new Node());
                NodePtr old_root = std::move(root);
                AABB e = old_root->env;
                e.merge(next.node->envelope());
                new_root->env = e;
                new_root->ch.push_back(std::move(old_root));
                new_root->ch.push_back(std::move(next.node));
                root = std::move(new_root);
                target_height += 1;
            } else {
                InsertionResult r = forced_insertion(*root, std::move(next.node), target_height);
                if (r.kind == InsertionResult::Split) stack.push_back({true, std::move(r.split)});
            }
        }
    }
    // nearest_neighbor_iter(&[x, y]).take(k): points in the order the iterator yields them
    void nearest(int x, int y, int k, std::vector<int>& out) const {
        MinHeap h;
        h.d.reserve(20);
        auto extend = [&](const Node& parent) {
            if ((variant() & 1) && h.d.size() < parent.ch.size()) {  // newer std: append + rebuild when the tail outweighs the heap
                for (auto& c : parent.ch) h.d.push_back(HeapItem{c->envelope().distance_2(x, y), c.get()});
                for (size_t n = h.d.size() / 2; n-- > 0;) h.sift_down(n);
                return;
            }
            for (auto& c : parent.ch) h.push(HeapItem{c->envelope().distance_2(x, y), c.get()});
        };
        extend(*root);
        HeapItem cur;
        while ((int)(out.size() / 2) < k && h.pop(cur)) {
            if (cur.node->leaf) { out.push_back(cur.node->pt[0]); out.push_back(cur.node->pt[1]); }
            else extend(*cur.node);
        }
    }
};

}  // namespace rstar_port
