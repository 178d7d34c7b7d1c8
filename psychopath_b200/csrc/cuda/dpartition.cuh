// dpartition.cuh — slot-ordered lists for the wavefront.
//
// The trace kernel finishes rays in an arbitrary order, and appending them to the shading queues (and the
// survivors of shading to the next wave's live list) in completion order interleaves the samples of ~3000
// different pixels: the shade kernels then read their 16 B pool columns from scattered slots, and the next wave's
// pools of 80 rays come from 80 different places of the image.  Measured with the lists sorted by slot (radix
// sort, experiment): shade stage -15 %, trace -3 % (profiles/README.md).
//
// So the kernels only write one byte per slot (the ray's shading class / "path continues"), and these three
// small kernels turn the byte column into per-class lists IN SLOT ORDER: a stable partition = count per block,
// scan the block counts, scatter.  A pass reads n bytes and writes <= 4 n bytes; the bytes are reset to 0xFF on
// the way, so the column is clean for the next wave.
#pragma once

namespace psyd {

constexpr int PART_THREADS = 256;
constexpr int PART_PER_THREAD = 16;  // one uint4 of key bytes
constexpr int PART_PER_BLOCK = PART_THREADS * PART_PER_THREAD;
constexpr uint8_t PART_NONE = 0xFF;

PSY_DEV uint32_t part_key(const uint4& v, int j) {
    const uint32_t w = j < 4 ? v.x : (j < 8 ? v.y : (j < 12 ? v.z : v.w));
    return (w >> ((j & 3) * 8)) & 255u;
}

// block_counts[c * nb + block] = number of slots of class c in the block's 4096 slots
template <int K>
__global__ void __launch_bounds__(PART_THREADS) k_part_count(const uint8_t* __restrict__ key, uint32_t n,
                                                             uint32_t* __restrict__ block_counts, uint32_t nb) {
    __shared__ uint32_t warp_cnt[K][PART_THREADS / 32];
    const uint32_t first = (blockIdx.x * PART_THREADS + threadIdx.x) * PART_PER_THREAD;
    uint32_t cnt[K];
#pragma unroll
    for (int c = 0; c < K; ++c) cnt[c] = 0;
    if (first < n) {
        const uint4 v = *reinterpret_cast<const uint4*>(key + first);
#pragma unroll
        for (int j = 0; j < PART_PER_THREAD; ++j) {
            const uint32_t k = part_key(v, j);
#pragma unroll
            for (int c = 0; c < K; ++c) cnt[c] += (k == (uint32_t)c && first + j < n) ? 1u : 0u;
        }
    }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < K; ++c) {
        const uint32_t s = __reduce_add_sync(0xffffffffu, cnt[c]);
        if (lane == 0) warp_cnt[c][warp] = s;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        uint32_t s = 0;
        for (int w = 0; w < PART_THREADS / 32; ++w) s += warp_cnt[threadIdx.x][w];
        block_counts[threadIdx.x * nb + blockIdx.x] = s;
    }
}

struct PartLists {
    uint32_t* list[4];
    uint32_t* total[4];
};

// exclusive scan of every class's block counts in place; *lists.total[c] = list length
template <int K>
__global__ void __launch_bounds__(1024) k_part_scan(uint32_t* __restrict__ block_counts, uint32_t nb, PartLists lists) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry_s;
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int c = 0; c < K; ++c) {
        if (threadIdx.x == 0) carry_s = 0;
        __syncthreads();
        for (uint32_t base = 0; base < nb; base += 1024u) {
            const uint32_t i = base + threadIdx.x;
            const uint32_t v = i < nb ? block_counts[c * nb + i] : 0u;
            uint32_t x = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if ((int)lane >= d) x += y;
            }
            if (lane == 31) warp_sum[warp] = x;
            __syncthreads();
            if (warp == 0) {
                uint32_t s = warp_sum[lane];
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(0xffffffffu, s, d);
                    if ((int)lane >= d) s += y;
                }
                warp_sum[lane] = s;
            }
            __syncthreads();
            const uint32_t carry = carry_s;
            const uint32_t incl = x + (warp ? warp_sum[warp - 1] : 0u);
            if (i < nb) block_counts[c * nb + i] = carry + incl - v;
            __syncthreads();
            if (threadIdx.x == 1023) carry_s = carry + incl;
            __syncthreads();
        }
        if (threadIdx.x == 0) *lists.total[c] = carry_s;
        __syncthreads();
    }
}

// writes the slots of class c, in slot order, to lists.list[c]; resets the key bytes
template <int K>
__global__ void __launch_bounds__(PART_THREADS) k_part_scatter(uint8_t* __restrict__ key, uint32_t n,
                                                               const uint32_t* __restrict__ block_offsets, uint32_t nb,
                                                               PartLists lists) {
    __shared__ uint32_t warp_cnt[K][PART_THREADS / 32];
    const uint32_t first = (blockIdx.x * PART_THREADS + threadIdx.x) * PART_PER_THREAD;
    uint4 v = make_uint4(~0u, ~0u, ~0u, ~0u);
    uint32_t cnt[K];
#pragma unroll
    for (int c = 0; c < K; ++c) cnt[c] = 0;
    if (first < n) {
        v = *reinterpret_cast<const uint4*>(key + first);
        *reinterpret_cast<uint4*>(key + first) = make_uint4(~0u, ~0u, ~0u, ~0u);
#pragma unroll
        for (int j = 0; j < PART_PER_THREAD; ++j) {
            const uint32_t k = part_key(v, j);
#pragma unroll
            for (int c = 0; c < K; ++c) cnt[c] += (k == (uint32_t)c && first + j < n) ? 1u : 0u;
        }
    }
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t pos[K];  // exclusive prefix of this thread inside the block
#pragma unroll
    for (int c = 0; c < K; ++c) {
        uint32_t x = cnt[c];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if ((int)lane >= d) x += y;
        }
        if (lane == 31) warp_cnt[c][warp] = x;
        pos[c] = x - cnt[c];
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < K; ++c) {
        uint32_t before = 0;
        for (uint32_t w = 0; w < warp; ++w) before += warp_cnt[c][w];
        pos[c] += before + block_offsets[c * nb + blockIdx.x];
    }
    if (first < n) {
#pragma unroll
        for (int j = 0; j < PART_PER_THREAD; ++j) {
            const uint32_t k = part_key(v, j);
            if (first + j < n) {
#pragma unroll
                for (int c = 0; c < K; ++c)
                    if (k == (uint32_t)c) lists.list[c][pos[c]++] = first + j;
            }
        }
    }
}

}  // namespace psyd
