// Exclusive scans: block helper, three-kernel device-wide scan over a loader functor, the loaders.
// Part of the CUDA engine (engine.cu includes these files in this order; they form one translation unit).
#pragma once

namespace gpugrep {

// ------------------------------------------------------------------------------------------------------------
// Exclusive scan over u64 values produced by a loader functor: three kernels (block sums, scan of sums, write).
// ------------------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long* s_warp, unsigned long long* s_total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    unsigned long long incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        unsigned long long w = lane < nw ? s_warp[lane] : 0ull;
        unsigned long long wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, wi, d);
            if (lane >= d) wi += t;
        }
        if (lane < nw) s_warp[lane] = wi - w;
        if (lane == 31) *s_total = wi;
    }
    __syncthreads();
    return s_warp[wid] + incl - v;
}

// Prefix sums are kept per GROUP of four 512-byte blocks (the unit one warp step of k_stream writes): a quarter of the
// scan work; consumers add the in-group part from the (adjacent) meta words.
constexpr int kGroupBlocks = 4;
struct LoadMetaGroup {   // sum over the group's blocks of: candidates << 32 | newlines
    const unsigned long long* meta;
    size_t nblk;
    __device__ unsigned long long operator()(size_t g) const {
        unsigned long long acc = 0;
        size_t b0 = g * kGroupBlocks;
#pragma unroll
        for (int u = 0; u < kGroupBlocks; u++) {
            if (b0 + u < nblk) {
                unsigned long long m = meta[b0 + u];
                acc += ((unsigned long long)__popc((uint32_t)m) << 32) | (m >> 32);
            }
        }
        return acc;
    }
};
// newlines before block `blk`: group prefix + the earlier blocks of its group
__device__ __forceinline__ uint32_t newlines_before_block(const unsigned long long* __restrict__ prefix_g, const unsigned long long* __restrict__ meta, size_t blk) {
    uint32_t c = (uint32_t)prefix_g[blk / kGroupBlocks];
    for (size_t b = blk - blk % kGroupBlocks; b < blk; b++) c += (uint32_t)(meta[b] >> 32);
    return c;
}
struct LoadU64 {
    const unsigned long long* p;
    __device__ unsigned long long operator()(size_t i) const { return p[i]; }
};
struct LoadU8 {
    const uint8_t* p;
    __device__ unsigned long long operator()(size_t i) const { return p[i]; }
};
struct LoadU32 {
    const uint32_t* p;
    __device__ unsigned long long operator()(size_t i) const { return p[i]; }
};

// `limit` (optional): device word that bounds the meaningful prefix of the input (value >> limit_shift: 32 selects the
// candidate count of Totals::meta_total, 0 a plain count); tiles entirely beyond it contribute zero and are skipped.
template <class Load>
__global__ void __launch_bounds__(kScanThreads) k_scan_sums(Load load, size_t n, unsigned long long* __restrict__ sums, size_t ntiles,
                                                            const unsigned long long* limit, int limit_shift) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_total;
    if (limit) {
        size_t lim = (size_t)(*limit >> limit_shift);
        if (lim < n) n = lim;
    }
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        if (tile * kScanTile >= n) { if (threadIdx.x == 0) sums[tile] = 0; continue; }
        size_t base = tile * kScanTile + (size_t)threadIdx.x * kScanItems;
        unsigned long long acc = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; k++) if (base + k < n) acc += load(base + k);
        block_exclusive_scan(acc, s_warp, &s_total);
        if (threadIdx.x == 0) sums[tile] = s_total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) k_scan_top(unsigned long long* __restrict__ sums, size_t nb, unsigned long long* __restrict__ total) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_total;
    size_t per = (nb + blockDim.x - 1) / blockDim.x;
    size_t lo = (size_t)threadIdx.x * per, hi = lo + per < nb ? lo + per : nb;
    unsigned long long acc = 0;
    for (size_t i = lo; i < hi; i++) acc += sums[i];
    unsigned long long run = block_exclusive_scan(acc, s_warp, &s_total);
    for (size_t i = lo; i < hi; i++) {
        unsigned long long v = sums[i];
        sums[i] = run;
        run += v;
    }
    if (threadIdx.x == 0) *total = s_total;
}

template <class Load>
__global__ void __launch_bounds__(kScanThreads) k_scan_write(Load load, size_t n, const unsigned long long* __restrict__ sums,
                                                             unsigned long long* __restrict__ out, size_t ntiles, const unsigned long long* limit,
                                                             int limit_shift) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_total;
    if (limit) {
        size_t lim = (size_t)(*limit >> limit_shift);
        if (lim < n) n = lim;
    }
    for (size_t tile = blockIdx.x; tile < ntiles && tile * kScanTile < n; tile += gridDim.x) {
        size_t base = tile * kScanTile + (size_t)threadIdx.x * kScanItems;
        unsigned long long vals[kScanItems];
        unsigned long long acc = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            vals[k] = base + k < n ? load(base + k) : 0ull;
            acc += vals[k];
        }
        unsigned long long run = block_exclusive_scan(acc, s_warp, &s_total) + sums[tile];
#pragma unroll
        for (int k = 0; k < kScanItems; k++) {
            if (base + k < n) out[base + k] = run;
            run += vals[k];
        }
        __syncthreads();
    }
}

}  // namespace gpugrep
