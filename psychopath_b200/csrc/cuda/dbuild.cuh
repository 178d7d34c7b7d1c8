// dbuild.cuh — SURVEY.md N3: the topology of a mesh BVH built on the GPU (LBVH).
//
// The reference builds its binary BVH top-down on one thread with a binned SAH (src/accel/bvh_base.rs:96-215,
// objects_split.rs:153-238): seconds at 10 M triangles.  This is the build for scenes whose build time matters
// (previews, animation): triangles are ordered along a 63-bit Morton curve of their centroids (cub radix sort), the
// binary radix tree over the sorted triangles is built in one parallel pass (Karras 2012), and every subtree of at
// most `tris_per_leaf` triangles (the reference's 3, triangle_mesh.rs:17) becomes one leaf.  Only the TOPOLOGY comes from
// here — triangle order, children, split axes, depth; the host then derives the time-sampled bounds, the 1.4x
// union rule and the BVH2 -> BVH4 collapse exactly as for its own SAH trees (bvh_build.h: build_bvh4_from_topology),
// so the trace kernels and the oracle read the same node format either way.  A different tree means different
// visit counts and tie-breaks, never different hits: tests/test_gpu_build.py.
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

}  // namespace psyd
