// bvh_build.h — host BVH construction restated from the reference so that the object order, tree
// topology, time-sampled bounds and traversal codes (and therefore primitive ids and tie-breaks)
// match what `psychopath` itself would build.
//
// Reference: src/accel/bvh_base.rs:96-215 (recursive_build, acc_bounds, 1.4x union collapse),
//            src/accel/objects_split.rs:153-238 (sah_split), :291-341 (median_split),
//            src/algorithm.rs:44-81 (partition), :203-231 (quick_select),
//            src/accel/bvh4.rs:186-358 (construct_from_base), sub_crates/bvh_order/src/lib.rs:30-37.
//
// Differences that do not change the result: SAH bins are accumulated as 13 buckets + prefix/suffix
// unions instead of 12 left/right pairs per object (identical boxes and counts because the split
// planes are monotone), and large subtrees are built on several threads.
#pragma once
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

#include "hmath.h"

namespace psyh {

constexpr size_t BVH_MAX_DEPTH = 42;
constexpr float USE_UNION_FACTOR = 1.4f;
constexpr int SAH_BIN_COUNT = 13;

// Bounds of object `id` are all[begin[id] .. begin[id+1]).
struct ObjBoundsView {
    const BBox* all;
    const uint32_t* begin;
    const BBox* ptr(uint32_t id) const { return all + begin[id]; }
    size_t count(uint32_t id) const { return begin[id + 1] - begin[id]; }
};

inline uint64_t log2_64(uint64_t n) { return n == 0 ? 0 : (uint64_t)(63 - __builtin_clzll(n)); }

// f(begin, end) over [0, n) on the host's cores (chunks of at least `grain` items)
template <typename F>
inline void parallel_for(size_t n, size_t grain, F f) {
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = std::min<size_t>(hw ? hw : 1, (n + grain - 1) / std::max<size_t>(grain, 1));
    if (nt <= 1) {
        f((size_t)0, n);
        return;
    }
    std::vector<std::thread> th;
    size_t per = (n + nt - 1) / nt;
    for (size_t t = 0; t < nt; ++t) {
        size_t b = t * per, e = std::min(n, b + per);
        if (b >= e) break;
        th.emplace_back([=]() { f(b, e); });
    }
    for (auto& x : th) x.join();
}

// partition (src/algorithm.rs:44-81): same scan/swap sequence, predicate called once per element.
template <typename T, typename P>
inline size_t partition_ref(T* slc, size_t len, P pred) {
    T* a = slc;
    T* b = slc + len;
    for (;;) {
        for (;;) {
            if (a == b) return (size_t)(a - slc);
            if (!pred(*a)) break;
            ++a;
        }
        for (;;) {
            --b;
            if (a == b) return (size_t)(a - slc);
            if (pred(*b)) break;
        }
        std::swap(*a, *b);
        ++a;
    }
}

// quick_select (src/algorithm.rs:203-231); `less(a, b)` == (order(a,b) == Ordering::Less)
template <typename T, typename L>
inline void quick_select_ref(T* slc, size_t len, size_t n, L less) {
    size_t left = 0, right = len;
    uint64_t seed = (uint64_t)n;
    for (;;) {
        size_t i = left + (size_t)(hash_u64((uint64_t)right, seed) % (uint64_t)(right - left));
        std::swap(slc[i], slc[right - 1]);
        T val = slc[right - 1];
        size_t ii = left + partition_ref(slc + left, right - 1 - left, [&](const T& e) { return less(e, val); });
        std::swap(slc[ii], slc[right - 1]);
        if (ii == n) return;
        if (ii > n)
            right = ii;
        else
            left = ii + 1;
        seed += 1;
    }
}

// sah_split (objects_split.rs:153-238) over object ids; tb[id] = bounds of the object at time 0.5.
inline std::pair<size_t, int> sah_split_ids(uint32_t* objs, size_t n, const BBox* tb_) {
    BBox cb;
    for (size_t i = 0; i < n; ++i) cb = bb_union_pt(cb, bb_center(tb_[objs[i]]));

    float divs[3][SAH_BIN_COUNT - 1];
    for (int d = 0; d < 3; ++d) {
        float extent = cb.max[d] - cb.min[d];
        for (int dv = 0; dv < SAH_BIN_COUNT - 1; ++dv) {
            float part = extent * ((float)(dv + 1) / (float)SAH_BIN_COUNT);
            divs[d][dv] = cb.min[d] + part;
        }
    }
    // 13 buckets per axis: bucket k = first div with centroid <= div (12 if none).
    BBox bucket_bb[3][SAH_BIN_COUNT];
    uint32_t bucket_n[3][SAH_BIN_COUNT] = {};
    for (size_t i = 0; i < n; ++i) {
        const BBox& tb = tb_[objs[i]];
        V3 centroid = (tb.min + tb.max) * 0.5f;
        for (int d = 0; d < 3; ++d) {
            float c = centroid[d];
            int k = 0;
            while (k < SAH_BIN_COUNT - 1 && !(c <= divs[d][k])) ++k;
            bucket_bb[d][k] = bb_union(bucket_bb[d][k], tb);
            bucket_n[d][k] += 1;
        }
    }
    int dim = 0;
    float div_n = 0.0f;
    float smallest = std::numeric_limits<float>::infinity();
    for (int d = 0; d < 3; ++d) {
        BBox lbb[SAH_BIN_COUNT - 1], rbb[SAH_BIN_COUNT - 1];
        uint32_t ln[SAH_BIN_COUNT - 1], rn[SAH_BIN_COUNT - 1];
        BBox acc;
        uint32_t cnt = 0;
        for (int dv = 0; dv < SAH_BIN_COUNT - 1; ++dv) {
            acc = bb_union(acc, bucket_bb[d][dv]);
            cnt += bucket_n[d][dv];
            lbb[dv] = acc;
            ln[dv] = cnt;
        }
        acc = BBox();
        cnt = 0;
        for (int dv = SAH_BIN_COUNT - 2; dv >= 0; --dv) {
            acc = bb_union(acc, bucket_bb[d][dv + 1]);
            cnt += bucket_n[d][dv + 1];
            rbb[dv] = acc;
            rn[dv] = cnt;
        }
        for (int dv = 0; dv < SAH_BIN_COUNT - 1; ++dv) {
            float left_cost = bb_area(lbb[dv]) * (float)ln[dv];
            float right_cost = bb_area(rbb[dv]) * (float)rn[dv];
            float left_diag = bb_diagonal(lbb[dv]);
            float right_diag = bb_diagonal(rbb[dv]);
            float tot = (left_cost * left_diag) + (right_cost * right_diag);
            if (tot < smallest) {
                dim = d;
                div_n = divs[d][dv];
                smallest = tot;
            }
        }
    }
    size_t split_i = partition_ref(objs, n, [&](uint32_t id) {
        const BBox& tb = tb_[id];
        float centroid = (tb.min[dim] + tb.max[dim]) * 0.5f;
        return centroid < div_n;
    });
    if (split_i < 1)
        split_i = 1;
    else if (split_i >= n)
        split_i = n - 1;
    return {split_i, dim};
}


struct BaseNode {
    bool leaf = false;
    uint8_t split_axis = 0;
    BaseNode* c[2] = {nullptr, nullptr};
    uint32_t obj_begin = 0, obj_end = 0;
    std::vector<BBox>* bounds_store = nullptr;  // chunk-owned storage
    uint32_t b0 = 0, b1 = 0;
    const BBox* bounds() const { return bounds_store->data() + b0; }
    size_t nbounds() const { return b1 - b0; }
};

struct BaseChunk {
    std::deque<BaseNode> nodes;
    std::vector<BBox> bounds;
};

class BVHBaseBuilder {
  public:
    // ids: the object list to be reordered in place.  tb[id]: bounds at time 0.5 (lerp_slice(bounder(obj), 0.5)).
    BVHBaseBuilder(uint32_t* ids, size_t n, ObjBoundsView ob, size_t objects_per_leaf)
        : ids_(ids), n_(n), ob_(ob), per_leaf_(objects_per_leaf) {
        tb_.resize(max_id_plus1());
        for (size_t i = 0; i < n_; ++i) {
            uint32_t id = ids_[i];
            tb_[id] = lerp_slice(ob_.ptr(id), ob_.count(id), 0.5f);
        }
        unsigned hw = std::thread::hardware_concurrency();
        threads_left_ = (int)(hw > 1 ? hw - 1 : 0);
    }

    BaseNode* build() {
        chunks_.emplace_back(new BaseChunk());
        BaseNode* root = recurse(0, 0, n_, chunks_.back().get());
        return root;
    }
    size_t depth() const { return depth_.load(); }

  private:
    size_t max_id_plus1() const {
        uint32_t m = 0;
        for (size_t i = 0; i < n_; ++i) m = ids_[i] > m ? ids_[i] : m;
        return (size_t)m + 1;
    }

    void note_depth(size_t d) {
        size_t cur = depth_.load();
        while (cur < d && !depth_.compare_exchange_weak(cur, d)) {
        }
    }

    // acc_bounds (bvh_base.rs:70-94)
    void acc_bounds(size_t begin, size_t end, std::vector<BBox>& cache) const {
        size_t max_len = 0;
        for (size_t i = begin; i < end; ++i) max_len = std::max(max_len, ob_.count(ids_[i]));
        cache.assign(max_len, BBox());
        for (size_t i = begin; i < end; ++i) {
            const BBox* b = ob_.ptr(ids_[i]);
            size_t nb = ob_.count(ids_[i]);
            if (nb == max_len) {
                for (size_t k = 0; k < nb; ++k) cache[k] = bb_union(cache[k], b[k]);
            } else {
                float s = (float)(max_len - 1);
                for (size_t k = 0; k < max_len; ++k) cache[k] = bb_union(cache[k], lerp_slice(b, nb, (float)k / s));
            }
        }
    }

    static void push_collapsed(const std::vector<BBox>& keys, std::vector<BBox>& out) {
        BBox u;
        for (auto& b : keys) u = bb_union(u, b);
        float area = 0.0f;
        for (auto& b : keys) area = area + bb_area(b);
        float average_area = area / (float)keys.size();
        if (bb_area(u) <= (average_area * USE_UNION_FACTOR)) {
            out.push_back(u);
        } else {
            out.insert(out.end(), keys.begin(), keys.end());
        }
    }

    std::pair<size_t, int> sah_split(size_t begin, size_t end) { return sah_split_ids(ids_ + begin, end - begin, tb_.data()); }

    // median_split (objects_split.rs:291-341)
    std::pair<size_t, int> median_split(size_t begin, size_t end) {
        uint32_t* objs = ids_ + begin;
        size_t n = end - begin;
        BBox bb;
        for (size_t i = 0; i < n; ++i) bb = bb_union(bb, tb_[objs[i]]);
        int axis = 0;
        float largest = -std::numeric_limits<float>::infinity();
        for (int i = 0; i < 3; ++i) {
            float extent = bb.max[i] - bb.min[i];
            if (extent > largest) {
                largest = extent;
                axis = i;
            }
        }
        size_t place = n / 2;
        if (place == 0) place = 1;
        quick_select_ref(objs, n, place, [&](uint32_t a, uint32_t b) {
            float ca = (tb_[a].min[axis] + tb_[a].max[axis]) * 0.5f;
            float cb2 = (tb_[b].min[axis] + tb_[b].max[axis]) * 0.5f;
            return ca < cb2;
        });
        return {place, axis};
    }

    BaseNode* recurse(size_t depth, size_t begin, size_t end, BaseChunk* chunk) {
        size_t n = end - begin;
        chunk->nodes.emplace_back();
        BaseNode* me = &chunk->nodes.back();
        me->bounds_store = &chunk->bounds;
        if (n <= per_leaf_) {
            std::vector<BBox> cache;
            acc_bounds(begin, end, cache);
            me->b0 = (uint32_t)chunk->bounds.size();
            push_collapsed(cache, chunk->bounds);
            me->b1 = (uint32_t)chunk->bounds.size();
            me->leaf = true;
            me->obj_begin = (uint32_t)begin;
            me->obj_end = (uint32_t)end;
            note_depth(depth);
            return me;
        }
        std::pair<size_t, int> sp;
        if ((size_t)log2_64((uint64_t)n) < (BVH_MAX_DEPTH - depth))
            sp = sah_split(begin, end);
        else
            sp = median_split(begin, end);
        size_t mid = begin + sp.first;
        me->split_axis = (uint8_t)sp.second;

        BaseNode *c0 = nullptr, *c1 = nullptr;
        bool spawned = false;
        if (n >= 65536) {
            int left = threads_left_.fetch_sub(1);
            if (left > 0) {
                spawned = true;
                BaseChunk* nc;
                {
                    std::lock_guard<std::mutex> g(mu_);
                    chunks_.emplace_back(new BaseChunk());
                    nc = chunks_.back().get();
                }
                std::thread th([&]() { c0 = recurse(depth + 1, begin, mid, nc); });
                c1 = recurse(depth + 1, mid, end, chunk);
                th.join();
                threads_left_.fetch_add(1);
            } else {
                threads_left_.fetch_add(1);
            }
        }
        if (!spawned) {
            c0 = recurse(depth + 1, begin, mid, chunk);
            c1 = recurse(depth + 1, mid, end, chunk);
        }
        me->c[0] = c0;
        me->c[1] = c1;
        std::vector<BBox> merged;
        merge_slices_append(c0->bounds(), c0->nbounds(), c1->bounds(), c1->nbounds(), merged,
                            [](const BBox& a, const BBox& b) { return bb_union(a, b); });
        me->b0 = (uint32_t)chunk->bounds.size();
        push_collapsed(merged, chunk->bounds);
        me->b1 = (uint32_t)chunk->bounds.size();
        return me;
    }

    uint32_t* ids_;
    size_t n_;
    ObjBoundsView ob_;
    size_t per_leaf_;
    std::vector<BBox> tb_;
    std::vector<std::unique_ptr<BaseChunk>> chunks_;
    std::mutex mu_;
    std::atomic<int> threads_left_{0};
    std::atomic<size_t> depth_{0};
};

// A binary tree whose topology comes from elsewhere (the GPU LBVH build): internal node i has children
// left[i] / right[i] and a split axis.  A child >= 0 is an internal node; a child < 0 is a leaf: code = ~child holds
// ids[code >> 2 .. (code >> 2) + (code & 3)) of the ALREADY ORDERED id list.
struct Bvh2Topology {
    uint32_t n_internal = 0;
    const int32_t* left = nullptr;
    const int32_t* right = nullptr;
    const uint8_t* axis = nullptr;
    int32_t root = 0;  // (>= 0 unless the tree is a single leaf)
    uint32_t depth = 0;
};

// The bounds side of BVHBaseBuilder::recurse on a given topology: leaves accumulate their objects' time keys
// (acc_bounds), inner nodes merge their children's (merge_slices_append), both collapse under the 1.4x union rule —
// so everything downstream (bvh4_emit, traversal) sees the same kind of tree as from the SAH builder.
class TopologyBoundsBuilder {
  public:
    TopologyBoundsBuilder(const uint32_t* ids, size_t n, ObjBoundsView ob, size_t per_leaf, const Bvh2Topology& t)
        : ids_(ids), n_(n), ob_(ob), per_leaf_(per_leaf), t_(t) {
        unsigned hw = std::thread::hardware_concurrency();
        threads_left_ = (int)(hw > 1 ? hw - 1 : 0);
    }
    BaseNode* build() {
        chunks_.emplace_back(new BaseChunk());
        return node(t_.root, chunks_.back().get(), n_);
    }

  private:
    void acc_bounds(size_t begin, size_t end, std::vector<BBox>& cache) const {
        size_t max_len = 0;
        for (size_t i = begin; i < end; ++i) max_len = std::max(max_len, ob_.count(ids_[i]));
        cache.assign(max_len, BBox());
        for (size_t i = begin; i < end; ++i) {
            const BBox* b = ob_.ptr(ids_[i]);
            size_t nb = ob_.count(ids_[i]);
            if (nb == max_len) {
                for (size_t k = 0; k < nb; ++k) cache[k] = bb_union(cache[k], b[k]);
            } else {
                float s = (float)(max_len - 1);
                for (size_t k = 0; k < max_len; ++k) cache[k] = bb_union(cache[k], lerp_slice(b, nb, (float)k / s));
            }
        }
    }
    static void push_collapsed(const std::vector<BBox>& keys, std::vector<BBox>& out) {
        BBox u;
        for (auto& b : keys) u = bb_union(u, b);
        float area = 0.0f;
        for (auto& b : keys) area = area + bb_area(b);
        float average_area = area / (float)keys.size();
        if (bb_area(u) <= (average_area * USE_UNION_FACTOR))
            out.push_back(u);
        else
            out.insert(out.end(), keys.begin(), keys.end());
    }
    // `weight`: an upper bound of the leaves below (only used to decide where to fork a thread)
    BaseNode* node(int32_t ref, BaseChunk* chunk, size_t weight) {
        chunk->nodes.emplace_back();
        BaseNode* me = &chunk->nodes.back();
        me->bounds_store = &chunk->bounds;
        if (ref < 0) {
            const uint32_t code = (uint32_t)(~ref);
            const size_t begin = code >> 2, end = begin + (code & 3u);
            if ((code & 3u) == 0u || (code & 3u) > per_leaf_ || end > n_) throw std::runtime_error("bad leaf in external BVH topology");
            std::vector<BBox> cache;
            acc_bounds(begin, end, cache);
            me->b0 = (uint32_t)chunk->bounds.size();
            push_collapsed(cache, chunk->bounds);
            me->b1 = (uint32_t)chunk->bounds.size();
            me->leaf = true;
            me->obj_begin = (uint32_t)begin;
            me->obj_end = (uint32_t)end;
            return me;
        }
        if ((uint32_t)ref >= t_.n_internal) throw std::runtime_error("bad node in external BVH topology");
        me->split_axis = t_.axis[ref];
        BaseNode *c0 = nullptr, *c1 = nullptr;
        bool spawned = false;
        if (weight >= 32768) {
            int left = threads_left_.fetch_sub(1);
            if (left > 0) {
                spawned = true;
                BaseChunk* nc;
                {
                    std::lock_guard<std::mutex> g(mu_);
                    chunks_.emplace_back(new BaseChunk());
                    nc = chunks_.back().get();
                }
                std::thread th([&]() { c0 = node(t_.left[ref], nc, weight / 2); });
                c1 = node(t_.right[ref], chunk, weight / 2);
                th.join();
                threads_left_.fetch_add(1);
            } else {
                threads_left_.fetch_add(1);
            }
        }
        if (!spawned) {
            c0 = node(t_.left[ref], chunk, weight / 2);
            c1 = node(t_.right[ref], chunk, weight / 2);
        }
        me->c[0] = c0;
        me->c[1] = c1;
        std::vector<BBox> merged;
        merge_slices_append(c0->bounds(), c0->nbounds(), c1->bounds(), c1->nbounds(), merged,
                            [](const BBox& a, const BBox& b) { return bb_union(a, b); });
        me->b0 = (uint32_t)chunk->bounds.size();
        push_collapsed(merged, chunk->bounds);
        me->b1 = (uint32_t)chunk->bounds.size();
        return me;
    }

    const uint32_t* ids_;
    size_t n_;
    ObjBoundsView ob_;
    size_t per_leaf_;
    Bvh2Topology t_;
    std::vector<std::unique_ptr<BaseChunk>> chunks_;
    std::mutex mu_;
    std::atomic<int> threads_left_{0};
};

// Flattened BVH4 with node indices local to this tree.
struct Bvh4Flat {
    uint32_t root = PSY_NONE;  // child reference
    std::vector<PsyBvh4Node> nodes;
    std::vector<PsyBBox4> motion;
    std::vector<BBox> bounds;  // root bounds (BVH4::_bounds)
    uint32_t depth = 0;
};

inline uint8_t calc_traversal_code(int kind, uint8_t top, uint8_t a, uint8_t b) {
    // kind: 0 Full(top,left,right) 1 Left(top,left) 2 Right(top,right) 3 TopOnly(top)
    switch (kind) {
        case 0: return (uint8_t)(top + a * 3 + b * 9);
        case 1: return (uint8_t)(top + a * 3 + 27);
        case 2: return (uint8_t)(top + a * 3 + (27 + 9));
        default: return (uint8_t)(top + (27 + 9 + 9));
    }
}

inline PsyBBox4 bbox4_from(const BBox b[4]) {
    PsyBBox4 o;
    for (int i = 0; i < 4; ++i) {
        o.v[0][i] = b[i].min.x;
        o.v[1][i] = b[i].max.x;
        o.v[2][i] = b[i].min.y;
        o.v[3][i] = b[i].max.y;
        o.v[4][i] = b[i].min.z;
        o.v[5][i] = b[i].max.z;
    }
    return o;
}

// One BVH4 node of construct_from_base (bvh4.rs:186-358): picks the 2-4 children of binary node `node` (its
// children, or their children where those are internal), the traversal code and the time-sampled child bounds, and
// appends the node with its child references still unset.  Returns its index.
inline uint32_t bvh4_make_node(const BaseNode* node, const BaseNode* children[4], int& child_count, Bvh4Flat& out) {
    const BaseNode* cl = node->c[0];
    const BaseNode* cr = node->c[1];
    children[0] = children[1] = children[2] = children[3] = nullptr;
    uint8_t code;
    if (!cl->leaf) {
        if (!cr->leaf) {
            child_count = 4;
            children[0] = cl->c[0];
            children[1] = cl->c[1];
            children[2] = cr->c[0];
            children[3] = cr->c[1];
            code = calc_traversal_code(0, node->split_axis, cl->split_axis, cr->split_axis);
        } else {
            child_count = 3;
            children[0] = cl->c[0];
            children[1] = cl->c[1];
            children[2] = cr;
            code = calc_traversal_code(1, node->split_axis, cl->split_axis, 0);
        }
    } else {
        if (!cr->leaf) {
            child_count = 3;
            children[0] = cl;
            children[1] = cr->c[0];
            children[2] = cr->c[1];
            code = calc_traversal_code(2, node->split_axis, cr->split_axis, 0);
        } else {
            child_count = 2;
            children[0] = cl;
            children[1] = cr;
            code = calc_traversal_code(3, node->split_axis, 0, 0);
        }
    }
    size_t bounds_len = 0;
    for (int i = 0; i < 4; ++i)
        if (children[i]) bounds_len = std::max(bounds_len, children[i]->nbounds());

    uint32_t my_index = (uint32_t)out.nodes.size();
    PsyBvh4Node nd{};
    nd.traversal_code = code;
    nd.n_children = (uint32_t)child_count;
    nd.n_keys = (uint32_t)bounds_len;
    nd.key_base = 0;
    if (bounds_len < 2) {
        BBox b[4];
        for (int i = 0; i < 4; ++i)
            if (children[i]) b[i] = children[i]->bounds()[0];
        nd.bounds = bbox4_from(b);
    } else {
        nd.key_base = (uint32_t)out.motion.size();
        for (size_t k = 0; k < bounds_len; ++k) {
            float time = (float)k / (float)(bounds_len - 1);
            BBox b[4];
            for (int i = 0; i < 4; ++i)
                if (children[i]) b[i] = lerp_slice(children[i]->bounds(), children[i]->nbounds(), time);
            PsyBBox4 key = bbox4_from(b);
            if (k == 0) nd.bounds = key;
            out.motion.push_back(key);
        }
    }
    for (int i = 0; i < 4; ++i) nd.child[i] = PSY_NONE;
    out.nodes.push_back(nd);
    return my_index;
}

// construct_from_base (bvh4.rs:186-358).  `obj_base` is added to leaf object starts so that leaves
// address the global triangle / instance arrays.
inline uint32_t bvh4_emit(const BaseNode* node, uint32_t obj_base, Bvh4Flat& out) {
    if (node->leaf) return PSY_CHILD_LEAF(obj_base + node->obj_begin, node->obj_end - node->obj_begin);
    const BaseNode* children[4];
    int child_count = 0;
    const uint32_t my_index = bvh4_make_node(node, children, child_count, out);
    for (int i = 0; i < child_count; ++i) {
        const uint32_t c = bvh4_emit(children[i], obj_base, out);
        out.nodes[my_index].child[i] = c;
    }
    return my_index;
}

// construct_from_base on the host's cores: the BVH4 nodes of the first levels are emitted here, every subtree below
// them by its own task into its own Bvh4Flat, and the pieces are then appended (child references and key bases
// shifted).  Node NUMBERING differs from the serial emit (subtrees are contiguous either way); nothing reads it.
struct Bvh4Task {
    const BaseNode* node;
    uint32_t parent;  // node index in the top part
    int slot;         // child slot to patch
};
inline uint32_t bvh4_emit_top(const BaseNode* node, uint32_t obj_base, Bvh4Flat& out, int depth, int cut,
                              std::vector<Bvh4Task>& tasks, uint32_t parent, int slot) {
    if (node->leaf) return PSY_CHILD_LEAF(obj_base + node->obj_begin, node->obj_end - node->obj_begin);
    if (depth == cut) {
        tasks.push_back(Bvh4Task{node, parent, slot});
        return PSY_NONE;  // patched once the subtree is built
    }
    const BaseNode* children[4];
    int child_count = 0;
    const uint32_t my_index = bvh4_make_node(node, children, child_count, out);
    for (int i = 0; i < child_count; ++i) {
        const uint32_t c = bvh4_emit_top(children[i], obj_base, out, depth + 1, cut, tasks, my_index, i);
        out.nodes[my_index].child[i] = c;
    }
    return my_index;
}
inline uint32_t bvh4_emit_parallel(const BaseNode* root, uint32_t obj_base, Bvh4Flat& out) {
    std::vector<Bvh4Task> tasks;
    uint32_t r = bvh4_emit_top(root, obj_base, out, 0, 3, tasks, PSY_NONE, 0);
    if (tasks.empty()) return r;
    std::vector<Bvh4Flat> parts(tasks.size());
    std::vector<uint32_t> roots(tasks.size());
    std::atomic<size_t> next{0};
    unsigned hw = std::thread::hardware_concurrency();
    size_t nt = std::min<size_t>(hw ? hw : 1, tasks.size());
    std::vector<std::thread> th;
    for (size_t t = 0; t < nt; ++t)
        th.emplace_back([&]() {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= tasks.size()) break;
                roots[i] = bvh4_emit(tasks[i].node, obj_base, parts[i]);
            }
        });
    for (auto& x : th) x.join();
    // append the parts: sizes first, then copy in parallel
    std::vector<size_t> node_off(tasks.size()), key_off(tasks.size());
    size_t n_nodes = out.nodes.size(), n_keys = out.motion.size();
    for (size_t i = 0; i < tasks.size(); ++i) {
        node_off[i] = n_nodes, key_off[i] = n_keys;
        n_nodes += parts[i].nodes.size(), n_keys += parts[i].motion.size();
    }
    out.nodes.resize(n_nodes);
    out.motion.resize(n_keys);
    parallel_for(tasks.size(), 1, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            const uint32_t no = (uint32_t)node_off[i], ko = (uint32_t)key_off[i];
            for (size_t j = 0; j < parts[i].nodes.size(); ++j) {
                PsyBvh4Node nd = parts[i].nodes[j];
                for (int c = 0; c < 4; ++c)
                    if (nd.child[c] != PSY_NONE && !PSY_CHILD_IS_LEAF(nd.child[c])) nd.child[c] += no;
                if (nd.n_keys > 1) nd.key_base += ko;
                out.nodes[no + j] = nd;
            }
            std::copy(parts[i].motion.begin(), parts[i].motion.end(), out.motion.begin() + ko);
        }
    });
    for (size_t i = 0; i < tasks.size(); ++i) out.nodes[tasks[i].parent].child[tasks[i].slot] = roots[i] + (uint32_t)node_off[i];
    return r;
}

// BVH4::from_objects (bvh4.rs:58-95): reorders ids in place.
inline Bvh4Flat build_bvh4(uint32_t* ids, size_t n, ObjBoundsView ob, size_t objects_per_leaf, uint32_t obj_base) {
    Bvh4Flat out;
    if (n == 0) return out;
    const bool timing = getenv("PSY_BUILD_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    BVHBaseBuilder b(ids, n, ob, objects_per_leaf);
    BaseNode* root = b.build();
    auto t1 = std::chrono::steady_clock::now();
    out.root = n >= 65536 ? bvh4_emit_parallel(root, obj_base, out) : bvh4_emit(root, obj_base, out);
    if (timing && n > 100000)
        std::fprintf(stderr, "[build] %zu objects: binary tree %.3fs, BVH4 collapse %.3fs\n", n,
                     std::chrono::duration<double>(t1 - t0).count(),
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
    out.depth = (uint32_t)(b.depth() / 2 + 1);
    out.bounds.assign(root->bounds(), root->bounds() + root->nbounds());
    return out;
}

// The same BVH4 from an externally built topology over the ordered id list (GPU LBVH).
inline Bvh4Flat build_bvh4_from_topology(const uint32_t* ids, size_t n, ObjBoundsView ob, size_t objects_per_leaf,
                                         uint32_t obj_base, const Bvh2Topology& topo) {
    Bvh4Flat out;
    if (n == 0) return out;
    const bool timing = getenv("PSY_BUILD_TIMING") != nullptr;
    auto t0 = std::chrono::steady_clock::now();
    TopologyBoundsBuilder b(ids, n, ob, objects_per_leaf, topo);
    BaseNode* root = b.build();
    auto t1 = std::chrono::steady_clock::now();
    out.root = n >= 65536 ? bvh4_emit_parallel(root, obj_base, out) : bvh4_emit(root, obj_base, out);
    if (timing && n > 100000)
        std::fprintf(stderr, "[build] %zu objects: bounds on the external topology %.3fs, BVH4 collapse %.3fs\n", n,
                     std::chrono::duration<double>(t1 - t0).count(),
                     std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count());
    out.depth = (uint32_t)(topo.depth / 2 + 1);
    out.bounds.assign(root->bounds(), root->bounds() + root->nbounds());
    return out;
}

}  // namespace psyh
