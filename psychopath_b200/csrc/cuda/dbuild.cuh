// dbuild.cuh — SURVEY.md N3: a mesh's BVH built on the GPU.
//
// The reference builds its binary BVH top-down on one thread with a binned SAH (src/accel/bvh_base.rs:96-215,
// objects_split.rs:153-238): seconds at 10 M triangles.  This is the build for scenes whose build time matters
// (previews, animation): triangles are ordered along a 63-bit Morton curve of their centroids (cub radix sort); the
// binary tree over the sorted triangles comes from PLOC (below; or Karras' radix tree, 2012, in one parallel pass), and
// every subtree of at most `tris_per_leaf` triangles (the reference's 3, triangle_mesh.rs:17) becomes one leaf.  The
// second half of the file turns that topology into the reference's BVH4 — time-sampled bounds, the 1.4x union rule,
// the BVH2 -> BVH4 collapse — in the host builder's arithmetic (bvh_build.h), so the trace kernels and the oracle read
// the same node format whichever side built it.  A different tree means different visit counts and tie-breaks, never
// different hits: tests/test_gpu_build.py.
#pragma once
#include <cub/cub.cuh>

namespace psyd {

PSY_DEV uint32_t lbvh_order_bits(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);  // unsigned order == float order
}
PSY_DEV float lbvh_from_order_bits(uint32_t u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

// centroid of each triangle's bounds at shutter time 0.5 (the host builder's split criterion: lerp_slice(bounds, 0.5),
// bvh_base.rs:57-68) and the box of all centroids (min / max as order-preserving bits, via atomics)
__global__ void __launch_bounds__(256) k_lbvh_centroids(const float* __restrict__ verts, uint32_t n_keys,
                                                        const uint32_t* __restrict__ tri_idx, uint32_t n_tris,
                                                        float3* __restrict__ cent, uint32_t* __restrict__ box6) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float3 c = make_float3(0, 0, 0);
    const bool on = i < n_tris;
    if (on) {
        const uint32_t v0 = tri_idx[3 * (size_t)i], v1 = tri_idx[3 * (size_t)i + 1], v2 = tri_idx[3 * (size_t)i + 2];
        uint32_t k0 = 0, k1 = 0;
        float a = 0.0f;
        if (n_keys > 1) {
            const float tmp = 0.5f * (float)(n_keys - 1);
            k0 = (uint32_t)tmp;
            a = tmp - (float)k0;
            k1 = min(k0 + 1u, n_keys - 1u);
        }
        float lo[3], hi[3];
        for (int d = 0; d < 3; ++d) {
            float mn[2], mx[2];
            const uint32_t ks[2] = {k0, k1};
            for (int q = 0; q < 2; ++q) {
                const float p0 = verts[((size_t)v0 * n_keys + ks[q]) * 3 + d], p1 = verts[((size_t)v1 * n_keys + ks[q]) * 3 + d],
                            p2 = verts[((size_t)v2 * n_keys + ks[q]) * 3 + d];
                mn[q] = fminf(p0, fminf(p1, p2));
                mx[q] = fmaxf(p0, fmaxf(p1, p2));
            }
            lo[d] = (mn[0] * (1.0f - a)) + (mn[1] * a);
            hi[d] = (mx[0] * (1.0f - a)) + (mx[1] * a);
        }
        c = make_float3((lo[0] + hi[0]) * 0.5f, (lo[1] + hi[1]) * 0.5f, (lo[2] + hi[2]) * 0.5f);
        cent[i] = c;
    }
    // block-wide min / max, then one atomic per block and component
    typedef cub::BlockReduce<float, 256> Reduce;
    __shared__ typename Reduce::TempStorage tmp;
    const float big = 3.402823466e+38f;
    const float vals[6] = {on ? c.x : big, on ? c.y : big, on ? c.z : big, on ? c.x : -big, on ? c.y : -big, on ? c.z : -big};
    for (int q = 0; q < 6; ++q) {
        const float r = q < 3 ? Reduce(tmp).Reduce(vals[q], cub::Min()) : Reduce(tmp).Reduce(vals[q], cub::Max());
        if (threadIdx.x == 0) {
            if (q < 3)
                atomicMin(box6 + q, lbvh_order_bits(r));
            else
                atomicMax(box6 + q, lbvh_order_bits(r));
        }
        __syncthreads();
    }
}

PSY_DEV uint64_t lbvh_spread21(uint32_t v) {  // 21 bits -> every third bit of 63
    uint64_t x = v & 0x1fffffu;
    x = (x | (x << 32)) & 0x1f00000000ffffull;
    x = (x | (x << 16)) & 0x1f0000ff0000ffull;
    x = (x | (x << 8)) & 0x100f00f00f00f00full;
    x = (x | (x << 4)) & 0x10c30c30c30c30c3ull;
    x = (x | (x << 2)) & 0x1249249249249249ull;
    return x;
}
__global__ void __launch_bounds__(256) k_lbvh_morton(const float3* __restrict__ cent, uint32_t n_tris,
                                                     const uint32_t* __restrict__ box6, uint64_t* __restrict__ keys,
                                                     uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_tris) return;
    const float3 c = cent[i];
    const float lo[3] = {lbvh_from_order_bits(box6[0]), lbvh_from_order_bits(box6[1]), lbvh_from_order_bits(box6[2])};
    const float hi[3] = {lbvh_from_order_bits(box6[3]), lbvh_from_order_bits(box6[4]), lbvh_from_order_bits(box6[5])};
    const float p[3] = {c.x, c.y, c.z};
    // one scale for all axes (the largest extent): cells are cubes, a flat mesh does not get split along its thin axis
    const float ext = fmaxf(hi[0] - lo[0], fmaxf(hi[1] - lo[1], hi[2] - lo[2]));
    uint32_t q[3];
    for (int d = 0; d < 3; ++d) {
        const float u = ext > 0.0f ? (p[d] - lo[d]) / ext : 0.0f;
        q[d] = (uint32_t)fminf(fmaxf(u * 2097152.0f, 0.0f), 2097151.0f);
    }
    keys[i] = (lbvh_spread21(q[0]) << 2) | (lbvh_spread21(q[1]) << 1) | lbvh_spread21(q[2]);
    vals[i] = i;
}
// length of the common prefix of sorted triangles i and j (keys made unique by the position); -1 out of range
PSY_DEV int lbvh_delta(const uint64_t* __restrict__ k, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = k[i], b = k[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}
// Karras, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees" (2012): internal node i of
// the binary radix tree over the n sorted keys, each found independently.  Children: >= 0 internal node, < 0 the
// single triangle at sorted position ~c.  rng[i] = (first, last) sorted position below node i.
__global__ void __launch_bounds__(256) k_lbvh_tree(const uint64_t* __restrict__ keys, int n, int32_t* __restrict__ left,
                                                   int32_t* __restrict__ right, uint8_t* __restrict__ axis,
                                                   int2* __restrict__ rng, int32_t* __restrict__ node_parent,
                                                   int32_t* __restrict__ leaf_parent) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n - 1) return;
    const int d = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) / 2;
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int32_t cl = lo == gamma ? ~gamma : gamma;
    const int32_t cr = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    left[i] = cl, right[i] = cr;
    rng[i] = make_int2(lo, hi);
    if (cl < 0)
        leaf_parent[~cl] = i;
    else
        node_parent[cl] = i;
    if (cr < 0)
        leaf_parent[~cr] = i;
    else
        node_parent[cr] = i;
    if (i == 0) node_parent[0] = -1;
    // split axis = the axis of the first key bit the two halves differ in (x above y above z in every triple)
    uint8_t ax = 0;
    if (dnode >= 1 && dnode < 64) {
        const int bit = 63 - dnode;  // 0 .. 62
        ax = (uint8_t)(2 - (bit % 3));
    }
    axis[i] = ax;
}
// A radix-tree node that spans at most `per_leaf` triangles becomes ONE leaf of the BVH (a Morton cell: spatially
// compact, unlike a fixed cut of the sorted list); the nodes above stay internal and are renumbered densely.
__global__ void __launch_bounds__(256) k_lbvh_flag(const int2* __restrict__ rng, int n_nodes, int per_leaf,
                                                   uint32_t* __restrict__ keep) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i < n_nodes) keep[i] = (rng[i].y - rng[i].x + 1) > per_leaf ? 1u : 0u;
}
PSY_DEV int32_t lbvh_child_ref(int32_t c, const int2* __restrict__ rng, const uint32_t* __restrict__ keep,
                               const uint32_t* __restrict__ new_id) {
    if (c < 0) return ~(int32_t)((((uint32_t)~c) << 2) | 1u);  // a single triangle
    if (keep[c]) return (int32_t)new_id[c];
    const int2 r = rng[c];
    return ~(int32_t)((((uint32_t)r.x) << 2) | (uint32_t)(r.y - r.x + 1));  // leaf: (first position << 2) | count
}
__global__ void __launch_bounds__(256) k_lbvh_emit(const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                                                   const uint8_t* __restrict__ axis, const int2* __restrict__ rng,
                                                   const uint32_t* __restrict__ keep, const uint32_t* __restrict__ new_id,
                                                   int n_nodes, int32_t* __restrict__ left2, int32_t* __restrict__ right2,
                                                   uint8_t* __restrict__ axis2) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n_nodes || !keep[i]) return;
    const uint32_t me = new_id[i];
    left2[me] = lbvh_child_ref(left[i], rng, keep, new_id);
    right2[me] = lbvh_child_ref(right[i], rng, keep, new_id);
    axis2[me] = axis[i];
}
// depth of the deepest leaf (root = 0) = the largest number of kept nodes above a triangle, like BVHBase::depth
__global__ void __launch_bounds__(256) k_lbvh_depth(const int32_t* __restrict__ node_parent, const int32_t* __restrict__ leaf_parent,
                                                    const uint32_t* __restrict__ keep, int n, uint32_t* __restrict__ depth) {
    const int j = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    uint32_t dep = 0;
    if (j < n) {
        int p = leaf_parent[j];
        while (p >= 0 && dep < 4096u) {
            dep += keep[p];
            p = node_parent[p];
        }
    }
    dep = __reduce_max_sync(0xffffffffu, dep);
    if ((threadIdx.x & 31u) == 0u && dep) atomicMax(depth, dep);
}

// ---------------------------------------------------------------------------------------------------------------
// PLOC (Meister & Bittner, "Parallel Locally-Ordered Clustering for Bounding Volume Hierarchy Construction", 2018):
// the binary tree of better quality than the radix tree, from the same Morton-sorted triangles.  Clusters start as
// the sorted triangles (bounds at shutter time 0.5, the host builder's criterion); every round each cluster looks
// PLOC_RADIUS places to either side for the partner whose union with it has the smallest surface area, mutual choices
// merge into a new node, the cluster list is compacted in order, until one cluster is left.  Everything is
// deterministic: partners are compared by (area, lower place, higher place), new nodes are numbered by a scan.
// Ids: < n a triangle (its place in the sorted list), >= n the internal node id - n.
#ifndef PSY_PLOC_RADIUS
#define PSY_PLOC_RADIUS 16
#endif
constexpr int PLOC_RADIUS = PSY_PLOC_RADIUS;

struct PlocBox {
    float mn[3], mx[3];
};
PSY_DEV float ploc_union_area(const PlocBox& a, const PlocBox& b) {
    const float dx = fmaxf(a.mx[0], b.mx[0]) - fminf(a.mn[0], b.mn[0]);
    const float dy = fmaxf(a.mx[1], b.mx[1]) - fminf(a.mn[1], b.mn[1]);
    const float dz = fmaxf(a.mx[2], b.mx[2]) - fminf(a.mn[2], b.mn[2]);
    return (dx * dy) + (dy * dz) + (dz * dx);
}
// bounds of the sorted triangles at shutter time 0.5
__global__ void __launch_bounds__(256) k_ploc_init(const float* __restrict__ verts, uint32_t n_keys,
                                                   const uint32_t* __restrict__ tri_idx, const uint32_t* __restrict__ order,
                                                   uint32_t n, PlocBox* __restrict__ box, uint32_t* __restrict__ cid) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = order[i];
    const uint32_t v[3] = {tri_idx[3 * (size_t)t], tri_idx[3 * (size_t)t + 1], tri_idx[3 * (size_t)t + 2]};
    uint32_t k0 = 0, k1 = 0;
    float a = 0.0f;
    if (n_keys > 1) {
        const float tmp = 0.5f * (float)(n_keys - 1);
        k0 = (uint32_t)tmp;
        a = tmp - (float)k0;
        k1 = min(k0 + 1u, n_keys - 1u);
    }
    PlocBox b;
    for (int d = 0; d < 3; ++d) {
        float mn[2], mx[2];
        const uint32_t ks[2] = {k0, k1};
        for (int q = 0; q < 2; ++q) {
            const float p0 = verts[((size_t)v[0] * n_keys + ks[q]) * 3 + d], p1 = verts[((size_t)v[1] * n_keys + ks[q]) * 3 + d],
                        p2 = verts[((size_t)v[2] * n_keys + ks[q]) * 3 + d];
            mn[q] = fminf(p0, fminf(p1, p2));
            mx[q] = fmaxf(p0, fmaxf(p1, p2));
        }
        b.mn[d] = (mn[0] * (1.0f - a)) + (mn[1] * a);
        b.mx[d] = (mx[0] * (1.0f - a)) + (mx[1] * a);
    }
    box[i] = b;
    cid[i] = i;
}
// nearest partner of every cluster within the radius; the block's clusters and their halo are staged in shared memory
__global__ void __launch_bounds__(256) k_ploc_nearest(const PlocBox* __restrict__ box, uint32_t m, uint32_t* __restrict__ nn) {
    __shared__ PlocBox tile[256 + 2 * PLOC_RADIUS];
    const int base = (int)(blockIdx.x * blockDim.x) - PLOC_RADIUS;
    for (int s = (int)threadIdx.x; s < 256 + 2 * PLOC_RADIUS; s += 256) {
        const int g = base + s;
        if (g >= 0 && g < (int)m) tile[s] = box[g];
    }
    __syncthreads();
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const PlocBox me = tile[threadIdx.x + PLOC_RADIUS];
    float best = CUDART_INF_F;
    uint32_t best_j = 0xFFFFFFFFu;
    const int lo = max((int)i - PLOC_RADIUS, 0), hi = min((int)i + PLOC_RADIUS, (int)m - 1);
    for (int j = lo; j <= hi; ++j) {
        if (j == (int)i) continue;
        const float a = ploc_union_area(me, tile[j - base]);
        // order of candidate pairs: (area, lower place, higher place) — the same from both ends of a pair
        bool better = a < best;
        if (a == best && best_j != 0xFFFFFFFFu) {
            const uint32_t lo_new = min(i, (uint32_t)j), hi_new = max(i, (uint32_t)j);
            const uint32_t lo_old = min(i, best_j), hi_old = max(i, best_j);
            better = lo_new < lo_old || (lo_new == lo_old && hi_new < hi_old);
        }
        if (better) best = a, best_j = (uint32_t)j;
    }
    nn[i] = best_j;
}
// lead[i] = 1 where cluster i merges with its partner and is the lower of the two; gone[i] = 1 for the higher one
__global__ void __launch_bounds__(256) k_ploc_flags(const uint32_t* __restrict__ nn, uint32_t m, uint32_t* __restrict__ lead,
                                                    uint32_t* __restrict__ stay) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t j = nn[i];
    const bool mutual = j != 0xFFFFFFFFu && nn[j] == i;
    lead[i] = (mutual && i < j) ? 1u : 0u;
    stay[i] = (mutual && i > j) ? 0u : 1u;
}
// merge the pairs and compact the cluster list in order
__global__ void __launch_bounds__(256) k_ploc_merge(const PlocBox* __restrict__ box, const uint32_t* __restrict__ cid,
                                                    const uint32_t* __restrict__ nn, const uint32_t* __restrict__ lead,
                                                    const uint32_t* __restrict__ lead_at, const uint32_t* __restrict__ stay,
                                                    const uint32_t* __restrict__ stay_at, uint32_t m, uint32_t n,
                                                    uint32_t node_base, PlocBox* __restrict__ box_out, uint32_t* __restrict__ cid_out,
                                                    uint32_t* __restrict__ left, uint32_t* __restrict__ right,
                                                    uint8_t* __restrict__ axis, uint32_t* __restrict__ cnt,
                                                    uint32_t* __restrict__ parent) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m || !stay[i]) return;
    PlocBox b = box[i];
    uint32_t id = cid[i];
    if (lead[i]) {
        const uint32_t j = nn[i];
        const PlocBox o = box[j];
        uint32_t a_id = id, b_id = cid[j];
        // split axis = where the two centres are furthest apart; the child with the lower centre goes left
        float sep[3];
        for (int d = 0; d < 3; ++d) sep[d] = (o.mn[d] + o.mx[d]) - (b.mn[d] + b.mx[d]);
        int ax = 0;
        if (fabsf(sep[1]) > fabsf(sep[ax])) ax = 1;
        if (fabsf(sep[2]) > fabsf(sep[ax])) ax = 2;
        if (sep[ax] < 0.0f) {
            const uint32_t t = a_id;
            a_id = b_id, b_id = t;
        }
        const uint32_t node = node_base + lead_at[i];  // internal node index
        left[node] = a_id, right[node] = b_id, axis[node] = (uint8_t)ax;
        cnt[node] = (a_id < n ? 1u : cnt[a_id - n]) + (b_id < n ? 1u : cnt[b_id - n]);
        parent[a_id] = node, parent[b_id] = node;
        for (int d = 0; d < 3; ++d) b.mn[d] = fminf(b.mn[d], o.mn[d]), b.mx[d] = fmaxf(b.mx[d], o.mx[d]);
        id = n + node;
    }
    const uint32_t at = stay_at[i];
    box_out[at] = b;
    cid_out[at] = id;
}
// The place of every triangle / the first place of every node in the leaf order of the finished tree: walking up,
// every time the node is a right child its left sibling's triangles come first.  ids as above; parent[] is indexed by id.
PSY_DEV uint32_t ploc_place(uint32_t id, uint32_t n, uint32_t root_node, const uint32_t* __restrict__ left,
                            const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ parent, uint32_t* kept_above,
                            uint32_t per_leaf) {
    uint32_t pos = 0, above = 0;
    while (id != n + root_node) {
        const uint32_t p = parent[id];
        const uint32_t l = left[p];
        if (l != id) pos += l < n ? 1u : cnt[l - n];
        above += cnt[p] > per_leaf ? 1u : 0u;
        id = n + p;
    }
    if (kept_above) *kept_above = above;
    return pos;
}
__global__ void __launch_bounds__(256) k_ploc_order(const uint32_t* __restrict__ order_in, uint32_t n, uint32_t root_node,
                                                    const uint32_t* __restrict__ left, const uint32_t* __restrict__ cnt,
                                                    const uint32_t* __restrict__ parent, uint32_t per_leaf,
                                                    uint32_t* __restrict__ order_out, uint32_t* __restrict__ depth) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t above = 0;
    if (i < n) {
        const uint32_t pos = ploc_place(i, n, root_node, left, cnt, parent, &above, per_leaf);
        order_out[pos] = order_in[i];
    }
    above = __reduce_max_sync(0xffffffffu, above);
    if ((threadIdx.x & 31u) == 0u && above) atomicMax(depth, above);
}
// keep[node] = 1 for the nodes that stay internal (more than per_leaf triangles below); start[node] = first place
__global__ void __launch_bounds__(256) k_ploc_keep(uint32_t n, uint32_t n_nodes, uint32_t root_node,
                                                   const uint32_t* __restrict__ left, const uint32_t* __restrict__ cnt,
                                                   const uint32_t* __restrict__ parent, uint32_t per_leaf,
                                                   uint32_t* __restrict__ keep, uint32_t* __restrict__ start) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes) return;
    keep[v] = cnt[v] > per_leaf ? 1u : 0u;
    start[v] = ploc_place(n + v, n, root_node, left, cnt, parent, nullptr, per_leaf);
}
// kept nodes, renumbered in reverse creation order (the root is created last: it becomes node 0)
__global__ void __launch_bounds__(256) k_ploc_emit(uint32_t n, uint32_t n_nodes, uint32_t n_keep, uint32_t root_node,
                                                   const uint32_t* __restrict__ left, const uint32_t* __restrict__ right,
                                                   const uint8_t* __restrict__ axis, const uint32_t* __restrict__ cnt,
                                                   const uint32_t* __restrict__ parent, const uint32_t* __restrict__ keep,
                                                   const uint32_t* __restrict__ keep_at, const uint32_t* __restrict__ start,
                                                   uint32_t per_leaf, int32_t* __restrict__ left2, int32_t* __restrict__ right2,
                                                   uint8_t* __restrict__ axis2) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_nodes || !keep[v]) return;
    const uint32_t me = n_keep - 1u - keep_at[v];
    auto ref = [&](uint32_t id) -> int32_t {
        if (id < n) {
            const uint32_t pos = ploc_place(id, n, root_node, left, cnt, parent, nullptr, per_leaf);
            return ~(int32_t)((pos << 2) | 1u);
        }
        const uint32_t c = id - n;
        if (keep[c]) return (int32_t)(n_keep - 1u - keep_at[c]);
        return ~(int32_t)((start[c] << 2) | cnt[c]);
    };
    left2[me] = ref(left[v]);
    right2[me] = ref(right[v]);
    axis2[me] = axis[v];
}

// ---------------------------------------------------------------------------------------------------------------
// The rest of a mesh BVH on the device (psy_mesh_bvh_build): time-sampled bounds of every node of the binary tree,
// the 1.4x union rule, the BVH2 -> BVH4 collapse and the traversal codes, in the host builder's arithmetic
// (csrc/host/bvh_build.h: TopologyBoundsBuilder, bvh4_make_node; reference bvh_base.rs:96-215, bvh4.rs:186-358), so
// that the same topology gives the same nodes bit for bit whichever side built them (tests/test_gpu_build.py).
//
// Every node of the binary tree has either one bounds key or the mesh's n_keys (a leaf unions its triangles key by
// key; a parent merges two children of 1 or n keys; the union rule then keeps all keys or one).  The bounds of a
// node are stored AT ITS PARENT (cb[parent][side][key]): the collapse reads the children of a node together.

constexpr int BUILD_MAX_KEYS = 8;  // more time samples than this: the host builds the mesh

struct BBox6 {
    float mn[3], mx[3];
};
// SSE minps / maxps (hmath.h: sse_min, sse_max): the second operand unless the first is strictly smaller / larger
PSY_DEV float b_min(float a, float b) { return a < b ? a : b; }
PSY_DEV float b_max(float a, float b) { return a > b ? a : b; }
PSY_DEV BBox6 b_empty() {
    BBox6 b;
    for (int d = 0; d < 3; ++d) b.mn[d] = CUDART_INF_F, b.mx[d] = -CUDART_INF_F;
    return b;
}
PSY_DEV BBox6 b_union(const BBox6& a, const BBox6& b) {
    BBox6 o;
    for (int d = 0; d < 3; ++d) o.mn[d] = b_min(a.mn[d], b.mn[d]), o.mx[d] = b_max(a.mx[d], b.mx[d]);
    return o;
}
PSY_DEV float b_area(const BBox6& b) {
    const float dx = b.mx[0] - b.mn[0], dy = b.mx[1] - b.mn[1], dz = b.mx[2] - b.mn[2];
    return ((dx * dy) + (dy * dz) + (dz * dx)) * 2.0f;
}
PSY_DEV BBox6 b_lerp(const BBox6& a, const BBox6& b, float alpha) {
    BBox6 o;
    for (int d = 0; d < 3; ++d) {
        o.mn[d] = (a.mn[d] * (1.0f - alpha)) + (b.mn[d] * alpha);
        o.mx[d] = (a.mx[d] * (1.0f - alpha)) + (b.mx[d] * alpha);
    }
    return o;
}
PSY_DEV BBox6 b_load(const float* p) {  // (written by another thread of this launch: read through L2)
    BBox6 b;
    for (int d = 0; d < 3; ++d) b.mn[d] = __ldcg(p + d), b.mx[d] = __ldcg(p + 3 + d);
    return b;
}
PSY_DEV void b_store(float* p, const BBox6& b) {
    for (int d = 0; d < 3; ++d) p[d] = b.mn[d], p[3 + d] = b.mx[d];
}
// lerp_slice (lerp.rs:20-35) over n stored keys
PSY_DEV BBox6 b_lerp_slice(const float* keys, uint32_t n, float alpha) {
    if (n == 1u || alpha == 1.0f) return b_load(keys + 6 * (size_t)(n - 1u));
    const float tmp = alpha * (float)(n - 1u);
    const uint32_t i1 = (uint32_t)tmp;
    const float alpha2 = tmp - (float)i1;
    return b_lerp(b_load(keys + 6 * (size_t)i1), b_load(keys + 6 * (size_t)(i1 + 1u)), alpha2);
}
// push_collapsed: one key (the union) when the union's area is at most 1.4x the keys' average area (bvh_base.rs:119-132)
PSY_DEV uint32_t b_collapse(BBox6* keys, uint32_t n) {
    BBox6 u = b_empty();
    for (uint32_t k = 0; k < n; ++k) u = b_union(u, keys[k]);
    float area = 0.0f;
    for (uint32_t k = 0; k < n; ++k) area = area + b_area(keys[k]);
    const float average_area = area / (float)n;
    if (b_area(u) <= (average_area * 1.4f)) {
        keys[0] = u;
        return 1u;
    }
    return n;
}

// parent[c] = 2 * parent + side for every internal node but the root (-1)
__global__ void __launch_bounds__(256) k_b2_parents(const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                                                    int n_int, int32_t* __restrict__ parent) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n_int) return;
    if (i == 0) parent[0] = -1;
    const int32_t cl = left[i], cr = right[i];
    if (cl >= 0) parent[cl] = 2 * i;
    if (cr >= 0) parent[cr] = 2 * i + 1;
}

// Bottom-up: one thread per leaf reference computes the leaf's bounds and climbs; at every node the second child to
// arrive merges the two children, applies the union rule and goes on (Karras 2012, section 4).
__global__ void __launch_bounds__(128) k_b2_refit(const float* __restrict__ verts, uint32_t n_keys,
                                                  const uint32_t* __restrict__ tri_idx, const uint32_t* __restrict__ order,
                                                  const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                                                  const int32_t* __restrict__ parent, int n_int, float* cb, uint8_t* cnk,
                                                  uint32_t* ready, float* root_bounds, uint32_t* root_nk) {
    const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (t >= 2 * n_int) return;
    int i = t >> 1, side = t & 1;
    const int32_t c = side ? right[i] : left[i];
    if (c >= 0) return;
    const uint32_t code = (uint32_t)~c, first = code >> 2, count = code & 3u;
    BBox6 keys[BUILD_MAX_KEYS];
    for (uint32_t k = 0; k < n_keys; ++k) keys[k] = b_empty();
    for (uint32_t j = first; j < first + count; ++j) {
        const uint32_t tri = order[j];
        const uint32_t v0 = tri_idx[3 * (size_t)tri], v1 = tri_idx[3 * (size_t)tri + 1], v2 = tri_idx[3 * (size_t)tri + 2];
        for (uint32_t k = 0; k < n_keys; ++k) {
            const float* p0 = verts + ((size_t)v0 * n_keys + k) * 3;
            const float* p1 = verts + ((size_t)v1 * n_keys + k) * 3;
            const float* p2 = verts + ((size_t)v2 * n_keys + k) * 3;
            BBox6 b;  // vmin(p0, vmin(p1, p2)) (triangle_mesh.rs:86-99)
            for (int d = 0; d < 3; ++d) b.mn[d] = b_min(p0[d], b_min(p1[d], p2[d])), b.mx[d] = b_max(p0[d], b_max(p1[d], p2[d]));
            keys[k] = b_union(keys[k], b);
        }
    }
    uint32_t nk = b_collapse(keys, n_keys);
    for (;;) {
        float* dst = cb + ((size_t)i * 2 + side) * n_keys * 6;
        for (uint32_t k = 0; k < nk; ++k) b_store(dst + 6 * k, keys[k]);
        cnk[(size_t)i * 2 + side] = (uint8_t)nk;
        __threadfence();
        if (atomicAdd(ready + i, 1u) == 0u) return;  // the sibling has not arrived: it will do the rest
        __threadfence();
        // merge_slices_append(c0, c1, union) (algorithm.rs:234-262)
        const float* s1 = cb + ((size_t)i * 2) * n_keys * 6;
        const float* s2 = cb + ((size_t)i * 2 + 1) * n_keys * 6;
        const uint32_t n1 = ((volatile uint8_t*)cnk)[(size_t)i * 2], n2 = ((volatile uint8_t*)cnk)[(size_t)i * 2 + 1];
        if (n1 == n2) {
            nk = n1;
            for (uint32_t k = 0; k < nk; ++k) keys[k] = b_union(b_load(s1 + 6 * k), b_load(s2 + 6 * k));
        } else if (n1 > n2) {
            nk = n1;
            const float s = (float)(n1 - 1u);
            for (uint32_t k = 0; k < nk; ++k) keys[k] = b_union(b_load(s1 + 6 * k), b_lerp_slice(s2, n2, (float)k / s));
        } else {
            nk = n2;
            const float s = (float)(n2 - 1u);
            for (uint32_t k = 0; k < nk; ++k) keys[k] = b_union(b_lerp_slice(s1, n1, (float)k / s), b_load(s2 + 6 * k));
        }
        nk = b_collapse(keys, nk);
        const int32_t p = parent[i];
        if (p < 0) {
            for (uint32_t k = 0; k < nk; ++k) b_store(root_bounds + 6 * k, keys[k]);
            *root_nk = nk;
            return;
        }
        i = p >> 1, side = p & 1;
    }
}

// construct_from_base takes a node's grandchildren wherever its children are internal: the BVH4 nodes are the
// internal nodes at even depth.  mcount = the motion keys the node will append (its longest child, if that is > 1).
__global__ void __launch_bounds__(256) k_b4_mark(const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                                                 const int32_t* __restrict__ parent, const uint8_t* __restrict__ cnk,
                                                 int n_int, uint32_t* __restrict__ is4, uint32_t* __restrict__ mcount) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n_int) return;
    uint32_t depth = 0;
    for (int32_t p = parent[i]; p >= 0; p = parent[p >> 1]) ++depth;
    const uint32_t mine = (depth & 1u) ? 0u : 1u;
    is4[i] = mine;
    uint32_t len = 0;
    if (mine) {
        const int32_t cs[2] = {left[i], right[i]};
        for (int s = 0; s < 2; ++s) {
            if (cs[s] >= 0)
                len = max(len, (uint32_t)max(cnk[(size_t)cs[s] * 2], cnk[(size_t)cs[s] * 2 + 1]));
            else
                len = max(len, (uint32_t)cnk[(size_t)i * 2 + s]);
        }
    }
    mcount[i] = len >= 2u ? len : 0u;
}

PSY_DEV uint32_t b4_traversal_code(int kind, uint32_t top, uint32_t a, uint32_t b) {
    // calc_traversal_code (sub_crates/bvh_order/src/lib.rs:30-37): 0 Full 1 Left 2 Right 3 TopOnly
    switch (kind) {
        case 0: return top + a * 3u + b * 9u;
        case 1: return top + a * 3u + 27u;
        case 2: return top + a * 3u + (27u + 9u);
        default: return top + (27u + 9u + 9u);
    }
}

// bvh4_make_node per BVH4 node: children, traversal code, time-sampled child bounds, child references
__global__ void __launch_bounds__(128) k_b4_emit(const int32_t* __restrict__ left, const int32_t* __restrict__ right,
                                                 const uint8_t* __restrict__ axis, const float* __restrict__ cb,
                                                 const uint8_t* __restrict__ cnk, const uint32_t* __restrict__ is4,
                                                 const uint32_t* __restrict__ idx4, const uint32_t* __restrict__ kbase,
                                                 int n_int, uint32_t n_keys, uint32_t obj_base, uint32_t node_base,
                                                 uint32_t key_base, PsyBvh4Node* __restrict__ nodes,
                                                 PsyBBox4* __restrict__ motion) {
    const int i = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    if (i >= n_int || !is4[i]) return;
    const int32_t cl = left[i], cr = right[i];
    int32_t ref[4] = {0, 0, 0, 0};
    const float* bp[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t nk[4] = {0, 0, 0, 0};
    int cc = 0;
    const size_t stride = (size_t)n_keys * 6;
    auto take = [&](int node, int side) {
        ref[cc] = side ? right[node] : left[node];
        bp[cc] = cb + ((size_t)node * 2 + side) * stride;
        nk[cc] = cnk[(size_t)node * 2 + side];
        ++cc;
    };
    uint32_t code;
    if (cl >= 0) {
        take(cl, 0), take(cl, 1);
        if (cr >= 0) {
            take(cr, 0), take(cr, 1);
            code = b4_traversal_code(0, axis[i], axis[cl], axis[cr]);
        } else {
            take(i, 1);
            code = b4_traversal_code(1, axis[i], axis[cl], 0);
        }
    } else {
        take(i, 0);
        if (cr >= 0) {
            take(cr, 0), take(cr, 1);
            code = b4_traversal_code(2, axis[i], axis[cr], 0);
        } else {
            take(i, 1);
            code = b4_traversal_code(3, axis[i], 0, 0);
        }
    }
    uint32_t len = 0;
    for (int j = 0; j < cc; ++j) len = max(len, nk[j]);
    PsyBvh4Node nd;
    nd.traversal_code = code;
    nd.n_children = (uint32_t)cc;
    nd.n_keys = len;
    const uint32_t my_keys = len >= 2u ? kbase[i] : 0u;  // (local: where this node's keys go in `motion`)
    nd.key_base = len >= 2u ? key_base + my_keys : 0u;
    for (int j = 0; j < 4; ++j) {
        if (j >= cc) {
            nd.child[j] = PSY_NONE;
        } else if (ref[j] < 0) {
            const uint32_t lc = (uint32_t)~ref[j];
            nd.child[j] = PSY_CHILD_LEAF(obj_base + (lc >> 2), lc & 3u);
        } else {
            nd.child[j] = node_base + idx4[ref[j]];
        }
    }
    const uint32_t n_out = len >= 2u ? len : 1u;
    for (uint32_t k = 0; k < n_out; ++k) {
        const float time = len >= 2u ? (float)k / (float)(len - 1u) : 0.0f;
        PsyBBox4 key;
        for (int j = 0; j < 4; ++j) {
            BBox6 b = b_empty();
            if (j < cc) b = len >= 2u ? b_lerp_slice(bp[j], nk[j], time) : b_load(bp[j]);
            key.v[0][j] = b.mn[0], key.v[1][j] = b.mx[0];
            key.v[2][j] = b.mn[1], key.v[3][j] = b.mx[1];
            key.v[4][j] = b.mn[2], key.v[5][j] = b.mx[2];
        }
        if (k == 0) nd.bounds = key;
        if (len >= 2u) motion[(size_t)my_keys + k] = key;
    }
    nodes[idx4[i]] = nd;
}

}  // namespace psyd
