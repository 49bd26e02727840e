// Exact confirmation of the gram hits of one 16-byte chunk: shared by k_confirm (one thread per candidate) and by the
// streaming kernel's CONFIRM variant (hit lanes, for very dense candidate sets).
// Part of the CUDA engine (engine.cu includes these files in this order; they form one translation unit).
#pragma once

namespace gpugrep {

// hash parameters of the streaming kernel's lookup tables (see k_stream.cuh)
struct ProbeParams {
    uint32_t mul, mul2;   // hash multipliers (mul2: second choice of the exact table)
    int shift;            // bloom: 32 - log2(bits).  exact: shift that turns the product into a BYTE offset (see below)
    uint32_t amask;       // exact: keeps the slot bits of the byte offset, clears the replica / word bits
    uint32_t half_bytes;  // exact: byte offset of the second half of the table
    int rshift;           // exact: log2 of the replication factor (copies interleaved across banks)
    // mixed sampling (Prefilter::odd): gram * odd_mul[k] + odd_add[k] == 0 at text offsets = 2 (mod 4).  Unused entries repeat
    // a used one.  The multipliers come from here (the parameter bank) so that the test stays ONE multiply-add on the FMA pipe.
    uint32_t odd_mul[2], odd_add[2];
};

// The exact gram set in global memory (two-choice table, Prefilter::confirm_keys), to find the hit positions inside a
// candidate chunk: the bloom table of k_stream only says "some sampled gram of this chunk MAY be in the set".
struct ReprobeParams {
    const uint32_t* keys;   // null: walk the whole chunk.  A gram lives in keys[h1] or keys[half + h2]
    const uint32_t* groups; // per slot of keys: the DFA groups (bit g mod 32) that can match around this gram
    uint32_t mul, mul2;     // h = (gram * mul) >> shift
    int shift;
    uint32_t half;
    int stride;
    int fold;
    int nodd;               // mixed sampling: compares at offsets 2 mod 4 (see ProbeParams)
    uint32_t odd_mul[2], odd_add[2];
    // extended confirmation (Prefilter::confirm_ext): per slot of keys the variants (bytes in front of the gram, 6 or 8
    // bytes in all); the text around the hit has to be in ext_keys (two-choice table of 64-bit keys) as well
    const uint32_t* ext_info;   // null: off
    const unsigned long long* ext_keys;
    unsigned long long ext_mul, ext_mul2;
    int ext_shift;
    uint32_t ext_half;
};

// 8 text bytes from an arbitrary offset (little endian); pos + 8 <= n
__device__ __forceinline__ unsigned long long load64_unaligned(const uint8_t* __restrict__ data, size_t pos, size_t n) {
    if (pos + 16 <= n) {
        const unsigned long long* p = reinterpret_cast<const unsigned long long*>(data + (pos & ~(size_t)7));
        const unsigned long long lo = p[0], hi = p[1];
        const uint32_t sh = 8u * (uint32_t)(pos & 7);
        return sh ? (lo >> sh) | (hi << (64u - sh)) : lo;
    }
    unsigned long long v = 0;
    for (int k = 0; k < 8; k++) v |= (unsigned long long)data[pos + k] << (8 * k);
    return v;
}

// Is the text around a gram hit at `q` one of the exact stretches the gram stands for?  info: Prefilter::confirm_ext.
__device__ bool ext_confirmed(const ReprobeParams& rp, const uint8_t* __restrict__ data, size_t n, size_t q, uint32_t info) {
    for (; info; info >>= 5) {
        const uint32_t before = info & 7u, len = (info & 8u) ? 8u : 6u;
        if (q < before || q - before + len > n) continue;   // a real occurrence lies inside the segment
        const size_t pos = q - before;
        unsigned long long v = pos + 8 <= n ? load64_unaligned(data, pos, n) : 0ull;
        if (pos + 8 > n) for (uint32_t k = 0; k < len; k++) v |= (unsigned long long)data[pos + k] << (8 * k);
        v |= 0x2020202020202020ull;   // the extended keys are folded on every byte (prefilter.cpp extension_of)
        if (len == 6u) v = (v & 0x0000ffffffffffffull) | 0xA5A5000000000000ull;
        const unsigned long long e1 = rp.ext_keys[(uint32_t)((v * rp.ext_mul) >> rp.ext_shift)];
        const unsigned long long e2 = rp.ext_keys[rp.ext_half + (uint32_t)((v * rp.ext_mul2) >> rp.ext_shift)];
        if (e1 == v || e2 == v) return true;
    }
    return false;
}

// Which sampled grams of the chunk at text offset `off` are really in the set, and which DFA groups own them?
// w: the four words of the chunk and the first word of the next one (case-folded if rp.fold); bloom: the byte table of
// the streaming kernel in shared memory (only the positions that pass it are looked up in the exact tables).
__device__ __forceinline__ void confirm_chunk(const uint32_t (&w)[5], size_t off, const uint8_t* bloom, const ProbeParams& pp, const ReprobeParams& rp,
                                              const uint8_t* __restrict__ data, size_t n, uint32_t& hits, uint32_t& group_mask) {
    uint32_t maybe = 0;   // bit = byte offset of a sampled gram that passes the bloom table
#pragma unroll
    for (int k = 0; k < 4; k++) {
#pragma unroll
        for (int sft = 0; sft < 4; sft++) {
            if (sft % rp.stride) continue;
            const uint32_t gram = sft == 0 ? w[k] : __funnelshift_r(w[k], w[k + 1], 8 * sft);
            const uint32_t p = gram * pp.mul;
            maybe |= ((bloom[p >> pp.shift] >> (p & 7u)) & 1u) << (4 * k + sft);
        }
    }
    hits = 0;
    group_mask = 0;
    while (maybe) {
        const uint32_t at = __ffs(maybe) - 1;
        maybe &= maybe - 1;
        // (the gram again from a 64-bit pair selected by comparisons: no dynamically indexed local array)
        const uint32_t k = at >> 2, sft = at & 3u;
        const uint32_t lo = k == 0 ? w[0] : (k == 1 ? w[1] : (k == 2 ? w[2] : w[3]));
        const uint32_t hi = k == 0 ? w[1] : (k == 1 ? w[2] : (k == 2 ? w[3] : w[4]));
        const uint32_t gram = __funnelshift_r(lo, hi, 8 * sft);
        const uint32_t h1 = (gram * rp.mul) >> rp.shift, h2 = rp.half + ((gram * rp.mul2) >> rp.shift);
        const uint32_t e1 = rp.keys[h1], e2 = rp.keys[h2];
        if (e1 == gram || e2 == gram) {
            const uint32_t slot = e1 == gram ? h1 : h2;
            const uint32_t info = rp.ext_info ? rp.ext_info[slot] : 0u;
            if (info == 0u || ext_confirmed(rp, data, n, off + at, info)) {
                hits |= 1u << at;
                group_mask |= rp.groups[slot];
            }
        }
    }
    if (rp.nodd) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const uint32_t gram = __funnelshift_r(w[k], w[k + 1], 16);
            for (int c = 0; c < rp.nodd; c++)
                if (gram * rp.odd_mul[c] + rp.odd_add[c] == 0u) { hits |= 1u << (4 * k + 2); group_mask = 0xffffffffu; }
        }
    }
}

}  // namespace gpugrep
