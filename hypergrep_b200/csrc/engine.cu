// CUDA engine for sm_100a: streaming newline/prefilter kernel, scans, candidate verification (DFA), compaction.
//
// Replaces, for one device-resident segment of file bytes, the reference's per-line loop
//   gzgets (line split)  -> hs_scan (match)          -> hs_callback (record)
//   hyperscanner.c:199      hyperscanner.c:217           hyperscanner.c:83-102
// Two paths produce identical results:
//   FAST    (simple mode + literal prefilter + no over-long lines):
//           k_stream -> scan -> k_list_candidates -> k_verify_simple -> scan -> k_emit_simple
//   GENERAL (everything else, and the fallback when a fast-path capacity bound is hit):
//           k_stream(no filter) -> scan -> k_newline_positions -> pseudo-line table -> k_match_pl_* -> scan -> emit
// All byte offsets inside a segment are 32-bit (segments are < 4 GiB); line numbers are rebased on the host.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "engine.hpp"
#include "nfa_sim.hpp"

namespace gpugrep {

// ------------------------------------------------------------------------------------------------------------
// device-side views
// ------------------------------------------------------------------------------------------------------------
struct GroupDev {
    const uint16_t* trans;      // [states][stride]
    const uint8_t* cls;         // [256] byte -> class
    const uint32_t* accept_of;  // [states] (general mode)
    const uint16_t* flat;       // [states][256] byte-indexed transitions (local verification: one load per byte), or null
    const uint16_t* eod_next;   // [states] transition on end-of-data (with `flat`)
    uint32_t stride, eod, first_accept, dead, accept_base, idle_end, mid_other, mid_word;
};

struct DbView {
    const GroupDev* groups;
    int ngroups;
    const NfaView* nfas;   // patterns simulated as bit-parallel NFAs (general path only)
    int nnfa;
};

constexpr uint32_t kInvalidLen = 0xffffffffu;   // LineRec.len of a record the host must drop (NUL re-check failed)
constexpr uint32_t kHasNulBit = 0x80000000u;    // LineRec.len flag: the line contains NUL bytes (host applies the strip/cut rule)

struct Totals {
    unsigned long long meta_total;   // candidates << 32 | newlines
    unsigned long long rec_total;    // records to emit (fast path) / generic scan totals
    unsigned long long aux_total;
    unsigned int flags;              // bit0: a 64 KiB super-block without newline; bit1: candidate overflow; bit2: record overflow
    unsigned int last_byte;
    unsigned int max_line;           // general path: longest line
    unsigned int pad;
};

#define CUDA_TRY(expr)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            error = std::string(#expr) + ": " + cudaGetErrorString(e_);                     \
            return 7;                                                                       \
        }                                                                                   \
    } while (0)

// ------------------------------------------------------------------------------------------------------------
// SWAR helpers
// ------------------------------------------------------------------------------------------------------------
// 0x80 in every byte of w equal to the byte replicated in `rep`.  Exact (no borrow between bytes):
// u = (w ^ rep) | 0x80 never borrows when 1 is subtracted per byte; bit 7 of the result is clear iff the low 7 bits
// matched, and ~w / rep bit 7 handling below makes the top bit exact for rep < 0x80.
__device__ __forceinline__ uint32_t eq_mask4_r(uint32_t w, uint32_t rep, uint32_t c80) {
    // the constants are operands of two three-input LOP3s: 3 instructions per word instead of 4
    uint32_t u, z;
    asm("lop3.b32 %0, %1, %2, %3, 0xBE;" : "=r"(u) : "r"(w), "r"(rep), "r"(c80));   // (w ^ rep) | 0x80808080
    uint32_t t = u - 0x01010101u;
    asm("lop3.b32 %0, %1, %2, %3, 0x02;" : "=r"(z) : "r"(t), "r"(w), "r"(c80));     // ~(t | w) & 0x80808080
    return z;   // valid for rep bytes < 0x80 ('\n' = 0x0a, NUL = 0x00)
}
__device__ __forceinline__ uint32_t eq_mask4(uint32_t w, uint32_t rep) { return eq_mask4_r(w, rep, 0x80808080u); }
// 4 flag bits (0x80 per byte) -> 4 contiguous bits
__device__ __forceinline__ uint32_t movemask4(uint32_t z) { return ((z >> 7) * 0x01020408u) >> 24 & 0xFu; }
// flag words of two consecutive words -> 8 contiguous bits (one multiply gathers both: no two partial products meet)
__device__ __forceinline__ uint32_t movemask8(uint32_t z0, uint32_t z1) { return (((z0 >> 7) | (z1 >> 3)) * 0x01020408u) >> 24; }

__device__ __forceinline__ uint32_t byte_mask16(const uint4& v, uint32_t rep) {
    return movemask8(eq_mask4(v.x, rep), eq_mask4(v.y, rep)) | (movemask8(eq_mask4(v.z, rep), eq_mask4(v.w, rep)) << 8);
}
__device__ __forceinline__ uint32_t newline_mask16(const uint4& v) { return byte_mask16(v, 0x0a0a0a0au); }
// flags of 16 bytes packed into bits 0..3 of every byte of one word (order does not matter to a count)
__device__ __forceinline__ uint32_t newline_flags16(const uint4& v) {
    uint32_t a = eq_mask4(v.x, 0x0a0a0a0au), b = eq_mask4(v.y, 0x0a0a0a0au), c = eq_mask4(v.z, 0x0a0a0a0au), d = eq_mask4(v.w, 0x0a0a0a0au);
    return (a >> 7) | (b >> 6) | (c >> 5) | (d >> 4);
}
__device__ __forceinline__ uint32_t newline_count16(const uint4& v) { return __popc(newline_flags16(v)); }

// streaming 16-byte load that does not pollute L1
__device__ __forceinline__ uint4 ld_stream16(const uint8_t* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// cached 16-byte load of the aligned chunk at `off`; bytes at or beyond n read as zero
__device__ __forceinline__ uint4 ld_chunk(const uint8_t* data, size_t off, size_t n) {
    uint4 v = *reinterpret_cast<const uint4*>(data + off);   // within the same 16-byte granule as byte n-1 at worst
    if (off + 16 > n) {
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            size_t b = off + 4 * i;
            if (b >= n) w[i] = 0;
            else if (b + 4 > n) w[i] &= (1u << (8 * (n - b))) - 1u;
        }
        v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    return v;
}

// cheap existence tests on a 16-byte chunk (exact: the borrow trick can only flag a byte above a real hit)
__device__ __forceinline__ uint32_t any_newline16(const uint4& v) {
    return eq_mask4(v.x, 0x0a0a0a0au) | eq_mask4(v.y, 0x0a0a0a0au) | eq_mask4(v.z, 0x0a0a0a0au) | eq_mask4(v.w, 0x0a0a0a0au);
}
__device__ __forceinline__ uint32_t haszero4(uint32_t w) { return (w - 0x01010101u) & ~w & 0x80808080u; }

// The line-extent searches below run one thread per matched line, 32 different lines per warp.  They are written as
// "a tight loop that only skips chunks without a hit, then the exact (expensive) look at the chunk that stopped the
// loop": the threads of a warp leave the loop at different iterations but meet again behind it, so the expensive part runs
// once per warp with every lane active instead of once per iteration with one or two lanes.  Each search can be bounded:
// a line that is not settled within `bound` bytes is handed to the warp-cooperative variants further down, which read
// 512 bytes per step (k_emit_simple: a 16 KiB JSON line is 32 steps for the warp instead of 1,000 for one thread).
constexpr size_t kNoBound = ~(size_t)0;

// Start of the line containing byte `pos` = index just past the last '\n' strictly before `pos` (0 if none).
// Returns true and the start in *out, or - after more than `bound` bytes without a newline - false and in *out a 16-byte
// aligned position p <= pos with no newline in [p, pos).
__device__ bool line_start_bounded(const uint8_t* data, size_t pos, size_t bound, size_t* out) {
    const size_t give_up = pos > bound ? pos - bound : 0;
    while (pos > 0) {
        size_t base;
        uint4 v;
        while (true) {   // skip whole chunks without a newline, four per step while that many lie below (four loads in flight)
            base = (pos - 1) & ~(size_t)15;
            if (base >= 48 && pos == base + 16) {
                if (pos <= give_up) { *out = pos; return false; }
                const uint4 a = *reinterpret_cast<const uint4*>(data + base), b = *reinterpret_cast<const uint4*>(data + base - 16);
                const uint4 c = *reinterpret_cast<const uint4*>(data + base - 32), d = *reinterpret_cast<const uint4*>(data + base - 48);
                if (any_newline16(a)) { v = a; break; }
                if (any_newline16(b)) { v = b; base -= 16; pos = base + 16; break; }
                if (any_newline16(c)) { v = c; base -= 32; pos = base + 16; break; }
                if (any_newline16(d)) { v = d; base -= 48; pos = base + 16; break; }
                pos = base - 48;
                if (pos == 0) { base = 0; v = make_uint4(0u, 0u, 0u, 0u); break; }   // reached the start of the data: no newline before
                continue;
            }
            v = *reinterpret_cast<const uint4*>(data + base);
            if (any_newline16(v) || base == 0) break;
            pos = base;
        }
        const uint32_t span = (uint32_t)(pos - base);   // bytes [base, pos) are candidates, 0..16
        uint32_t m = newline_mask16(v);
        if (span < 16) m &= (1u << span) - 1u;
        if (m) { *out = base + (32 - __clz(m)); return true; }
        pos = base;   // the newlines of this chunk lie at or behind pos (first chunk only), or base == 0
    }
    *out = 0;
    return true;
}
__device__ size_t line_start_of(const uint8_t* data, size_t pos) {
    size_t st;
    line_start_bounded(data, pos, kNoBound, &st);
    return st;
}

// End of the line that contains byte `pos` = index just past the first '\n' at or after `pos`, or n if there is none;
// *has_nul is set if a NUL byte lies in [pos, end).  Returns true and the end in *out, or - after more than `bound` bytes
// without a newline - false and in *out a 16-byte aligned position p > pos with no newline in [pos, p) (*has_nul then
// covers [pos, p)).
__device__ bool line_end_bounded(const uint8_t* data, size_t pos, size_t n, size_t bound, size_t* out, bool* has_nul) {
    size_t base = pos & ~(size_t)15;
    const size_t give_up = bound == kNoBound ? kNoBound : pos + bound;
    uint32_t skip = (uint32_t)(pos - base);
    bool nul = false;
    size_t end = n;
    bool found = true;
    while (base < n) {
        uint4 v;
        // one test for "a '\n' or a NUL may be here": with bits 1 and 3 cleared both become zero bytes (so do 0x02 and 0x08,
        // which only cost the exact look below); bytes at or beyond n read as zero and stop the loop as well
        auto maybe = [](const uint4& q) {
            const uint32_t k = 0xf5f5f5f5u;
            return (haszero4(q.x & k) | haszero4(q.y & k) | haszero4(q.z & k) | haszero4(q.w & k)) != 0;
        };
        while (true) {   // four chunks per step while that many lie inside the segment (four loads in flight)
            if (skip == 0 && base + 64 <= n) {
                if (base >= give_up) { found = false; break; }
                const uint4 a = *reinterpret_cast<const uint4*>(data + base), b = *reinterpret_cast<const uint4*>(data + base + 16);
                const uint4 c = *reinterpret_cast<const uint4*>(data + base + 32), d = *reinterpret_cast<const uint4*>(data + base + 48);
                if (maybe(a)) { v = a; break; }
                if (maybe(b)) { v = b; base += 16; break; }
                if (maybe(c)) { v = c; base += 32; break; }
                if (maybe(d)) { v = d; base += 48; break; }
                base += 64;
                if (base >= n) break;
                continue;
            }
            v = ld_chunk(data, base, n);
            if (maybe(v) || skip != 0) break;
            base += 16;
            if (base >= n) break;
        }
        if (!found) { end = base; break; }
        if (base >= n) break;
        const uint32_t valid = (base + 16 > n ? (1u << (n - base)) - 1u : 0xffffu) & ~((1u << skip) - 1u);
        const uint32_t m = newline_mask16(v) & valid;
        uint32_t zm = byte_mask16(v, 0u) & valid;
        if (m) {
            end = base + __ffs(m);
            zm &= (1u << __ffs(m)) - 1u;
        }
        if (zm) nul = true;
        if (m) break;
        skip = 0;
        base += 16;
    }
    if (has_nul) *has_nul = nul;
    *out = end;
    return found;
}
__device__ size_t line_end_of(const uint8_t* data, size_t pos, size_t n, bool* has_nul) {
    size_t en;
    line_end_bounded(data, pos, n, kNoBound, &en, has_nul);
    return en;
}

// Warp-cooperative continuations (every lane of the warp calls them with the same arguments).
// Last '\n' strictly before the 16-byte aligned `pos`: the index just past it, or 0.
__device__ size_t warp_line_start(const uint8_t* data, size_t pos) {
    const uint32_t lane = threadIdx.x & 31;
    while (pos > 0) {
        // lane 0 takes the chunk just below pos, lane 31 the one 512 bytes further down
        const bool have = pos >= (size_t)16 * (lane + 1);
        const size_t base = have ? pos - (size_t)16 * (lane + 1) : 0;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        if (have) v = *reinterpret_cast<const uint4*>(data + base);
        const uint32_t hit = __ballot_sync(0xffffffffu, have && any_newline16(v) != 0);
        if (hit) {
            const int src = __ffs(hit) - 1;   // the nearest chunk with a newline
            const size_t mine = base + (32 - __clz(newline_mask16(v) | 1u));   // only the value of lane `src` is used
            return (size_t)__shfl_sync(0xffffffffu, (unsigned long long)mine, src);
        }
        if (pos <= 512) return 0;
        pos -= 512;
    }
    return 0;
}
// First '\n' at or after the 16-byte aligned `pos`: the index just past it, or n; *has_nul: a NUL lies in [pos, end).
__device__ size_t warp_line_end(const uint8_t* data, size_t pos, size_t n, bool* has_nul) {
    const uint32_t lane = threadIdx.x & 31;
    bool nul = false;
    size_t end = n;
    while (pos < n) {
        const size_t base = pos + (size_t)16 * lane;
        const bool have = base < n;
        uint4 v = make_uint4(0u, 0u, 0u, 0u);
        uint32_t valid = 0;
        if (have) {
            v = ld_chunk(data, base, n);
            valid = base + 16 > n ? (1u << (n - base)) - 1u : 0xffffu;
        }
        const uint32_t m = have ? newline_mask16(v) & valid : 0u;
        const uint32_t zm = have ? byte_mask16(v, 0u) & valid : 0u;
        const uint32_t hit = __ballot_sync(0xffffffffu, m != 0);
        const uint32_t src = hit ? (uint32_t)__ffs(hit) - 1u : 32u;   // the first chunk with a newline
        // NULs count in the chunks before that one, and in it before the newline
        const bool counts = lane < src ? zm != 0 : (lane == src && (zm & ((1u << __ffs(m)) - 1u)) != 0);
        if (__any_sync(0xffffffffu, counts)) nul = true;
        if (hit) {
            end = (size_t)__shfl_sync(0xffffffffu, (unsigned long long)(base + __ffs(m | 0x10000u)), (int)src);
            break;
        }
        pos += 512;
    }
    *has_nul = nul;
    return end;
}

// newlines in [from, to); both ends arbitrary, to <= n.  Reads whole aligned 16-byte granules that overlap the range.
__device__ uint32_t count_newlines(const uint8_t* data, size_t from, size_t to) {
    if (from >= to) return 0;
    size_t b = from & ~(size_t)15;
    uint32_t c = 0;
    if (b != from || b + 16 > to) {   // first granule, partially inside the range
        uint32_t m = newline_mask16(*reinterpret_cast<const uint4*>(data + b)) & ~((1u << (from - b)) - 1u);
        if (b + 16 > to) m &= (1u << (to - b)) - 1u;
        c = __popc(m);
        b += 16;
    }
    for (; b + 32 <= to; b += 32)   // two granules per population count
        c += __popc(newline_flags16(*reinterpret_cast<const uint4*>(data + b)) | (newline_flags16(*reinterpret_cast<const uint4*>(data + b + 16)) << 4));
    if (b + 16 <= to) { c += newline_count16(*reinterpret_cast<const uint4*>(data + b)); b += 16; }
    if (b < to) c += __popc(newline_mask16(*reinterpret_cast<const uint4*>(data + b)) & ((1u << (to - b)) - 1u));
    return c;
}

// ------------------------------------------------------------------------------------------------------------
// K1: streaming kernel.  One warp owns four consecutive 512-byte blocks per step: 4 x (32 x 16-byte) coalesced
// loads in flight, newline count (SWAR + popc + warp reduce) and, when the prefilter is on, one gram-table lookup
// per sampled 4-byte gram (shared-memory table of exact keys, or a bloom bitmap for huge gram sets).
// Output: meta[block] = newline_count << 32 | ballot(lanes whose 16-byte chunk has a gram hit).
// STRIDE: sample every STRIDE-th byte position (4, 2, 1).  MODE: 0 no prefilter, 1 exact keys, 2 bloom bitmap.
// Algorithmic traffic: 1 byte read per input byte + 8 bytes written per 512.
// ------------------------------------------------------------------------------------------------------------
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

// Gram lookups of one 16-byte chunk.  MODE 1: two-choice table of exact 32-bit keys; the table is replicated
// 2^rshift times with the copies interleaved word by word, and a lane only ever reads copy (lane mod 2^rshift):
// with 32 copies every lane stays in its own shared-memory bank and the loads are conflict-free.
// Byte offset of slot h for this lane = ((gram * mul) >> shift) & amask | replica4, where replica4 = 4 * copy.
__device__ __forceinline__ uint32_t lds32(uint32_t shared_addr) {
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(shared_addr));
    return v;
}

__device__ __forceinline__ uint32_t lds8(uint32_t shared_addr) {
    uint32_t v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(shared_addr));
    return v;
}

// c1 / c2: shared-window address of table half 1 / 2 (aligned to the size of a half) OR-ed with 4 * copy of this lane,
// so that an address is formed by ONE logic op: ((product >> shift) & amask) | c.
template <int STRIDE, bool FOLD, int MODE, int NODD>
__device__ __forceinline__ bool probe_chunk(const uint4& v, uint32_t next, const uint32_t* __restrict__ tab, const ProbeParams& pp, uint32_t c1,
                                            uint32_t c2) {
    if (MODE == 0) return false;
    uint32_t w[5] = {v.x, v.y, v.z, v.w, next};
    if (FOLD) {
#pragma unroll
        for (int i = 0; i < 5; i++) w[i] |= 0x20202020u;
    }
    uint32_t miss = 0xffffffffu;   // min over all lookups of (key ^ gram): 0 iff some key matched
    uint32_t bits = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int s = 0; s < 4; s += STRIDE) {
            uint32_t gram = s == 0 ? w[i] : __funnelshift_r(w[i], w[i + 1], 8 * s);
            if (MODE == 1) {
                uint32_t e1 = lds32((((gram * pp.mul) >> pp.shift) & pp.amask) | c1);
                uint32_t e2 = lds32((((gram * pp.mul2) >> pp.shift) & pp.amask) | c2);
                miss = __vimin3_u32(miss, e1 - gram, e2 - gram);   // differences, not XORs: ptxas can place subtractions on the FMA pipe
            } else {
                // bloom: one byte load, bit (p & 7) of it.  The byte is replicated into all four bytes of a word (one
                // multiply on the FMA pipe) so that the wrap-around shift by p itself lands on the right bit.
                uint32_t p = gram * pp.mul;
                uint32_t b = lds8((p >> pp.shift) + c1);
                bits |= (b * 0x01010101u) >> (p & 31u);
            }
        }
    }
    if (NODD > 0) {
        // the grams at offsets 2, 6, 10, 14 of the chunk against two constants: two multiply-adds and one three-way minimum,
        // no shared-memory traffic
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t gram = __funnelshift_r(w[i], w[i + 1], 16);
            uint32_t x[2];
#pragma unroll
            for (int k = 0; k < 2; k++) asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(x[k]) : "r"(gram), "r"(pp.odd_mul[k]), "r"(pp.odd_add[k]));
            miss = __vimin3_u32(miss, x[0], x[1]);
        }
        return (MODE == 1 ? false : (bits & 1u) != 0u) || miss == 0u;
    }
    return MODE == 1 ? miss == 0u : (bits & 1u) != 0u;
}

// newlines in a 16-byte chunk: four flag words (bit 7 of matching bytes) are merged into one 64-bit word with three
// multiply-adds (FMA pipe) instead of shifts and ORs (ALU pipe, the pipe this kernel saturates first)
__device__ __forceinline__ uint32_t newline_count16_fma(const uint4& v, uint32_t cnl, uint32_t c80) {
    uint32_t a = eq_mask4_r(v.x, cnl, c80), b = eq_mask4_r(v.y, cnl, c80), c = eq_mask4_r(v.z, cnl, c80), d = eq_mask4_r(v.w, cnl, c80);
    unsigned long long acc = a;
    asm("mad.wide.u32 %0, %1, 2, %0;" : "+l"(acc) : "r"(b));
    asm("mad.wide.u32 %0, %1, 4, %0;" : "+l"(acc) : "r"(c));
    asm("mad.wide.u32 %0, %1, 8, %0;" : "+l"(acc) : "r"(d));
    return __popcll(acc);
}

constexpr int kStreamU = 4;   // 512-byte blocks per warp step

template <int STRIDE, bool FOLD, int MODE, int NODD>
__global__ void __launch_bounds__(1024) k_stream(const uint8_t* __restrict__ data, size_t n, unsigned long long* __restrict__ meta,
                                                 const uint32_t* __restrict__ table, int table_words, ProbeParams pp) {
    extern __shared__ __align__(16) uint32_t s_raw[];
    // exact tables are placed at an address aligned to the size of one half (see probe_chunk); the launch reserves the slack
    uint32_t* s_tab = s_raw;
    uint32_t saddr = (uint32_t)__cvta_generic_to_shared(s_raw);
    if (MODE == 1) {
        uint32_t aligned = (saddr + pp.half_bytes - 1u) & ~(pp.half_bytes - 1u);
        s_tab = s_raw + ((aligned - saddr) >> 2);
        saddr = aligned;
    }
    if (MODE != 0) {
        for (int i = threadIdx.x; i < table_words; i += blockDim.x) s_tab[i] = table[i];
        __syncthreads();
    }
    constexpr int U = kStreamU;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t replica4 = (lane & ((1u << pp.rshift) - 1u)) << 2;
    uint32_t c1 = MODE == 1 ? (saddr | replica4) : saddr, c2 = (saddr + pp.half_bytes) | replica4;
    asm volatile("mov.u32 %0, %0;" : "+r"(c1));   // materialise: each table address is then a single (x & amask) | c
    asm volatile("mov.u32 %0, %0;" : "+r"(c2));
    uint32_t cnl, c80;   // opaque to the optimiser so that they stay in registers (see eq_mask4_r)
    asm volatile("mov.u32 %0, 0x0a0a0a0a;" : "=r"(cnl));
    asm volatile("mov.u32 %0, 0x80808080;" : "=r"(c80));
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const size_t nblk = (n + 511) >> 9;
    const size_t nfull = n >> 9;   // blocks that lie entirely inside [0, n)

    // ---- main loop: groups of U full blocks, no bounds checks on the data loads
    for (size_t g0 = warp * U; g0 + U <= nfull; g0 += nwarps * U) {
        const uint8_t* p = data + (g0 << 9) + lane * 16;
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = ld_stream16(p + u * 512);
        uint32_t after = 0;   // first word after the group (only lane 31 needs it, for grams that straddle the end)
        if (MODE != 0 && (STRIDE < 4 || NODD > 0) && lane == 31) {
            size_t off = (g0 + U) << 9;
            if (off + 4 <= n) after = *reinterpret_cast<const uint32_t*>(data + off);
            else if (off < n) after = ld_chunk(data, off, n).x;
        }
        uint32_t cnt01, cnt23, masks[U];
        {
            uint32_t c[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                c[u] = newline_count16_fma(v[u], cnl, c80);
                uint32_t nx = 0;
                if (MODE != 0 && (STRIDE < 4 || NODD > 0)) {
                    // first word of the next chunk: lane+1's word of this block, or (lane 31) lane 0's word of the next block
                    uint32_t give = (u + 1 < U && lane == 0) ? v[u + 1 < U ? u + 1 : u].x : v[u].x;
                    nx = __shfl_sync(0xffffffffu, give, (lane + 1) & 31);
                    if (u + 1 == U && lane == 31) nx = after;
                }
                bool hit = probe_chunk<STRIDE, FOLD, MODE, NODD>(v[u], nx, s_tab, pp, c1, c2);
                masks[u] = __ballot_sync(0xffffffffu, hit);
            }
            cnt01 = __reduce_add_sync(0xffffffffu, c[0] | (c[1] << 16));
            cnt23 = __reduce_add_sync(0xffffffffu, c[2] | (c[3] << 16));
        }
        if (lane == 0) {
            uint4* out = reinterpret_cast<uint4*>(meta + g0);   // g0 is a multiple of 4: 32-byte aligned
            out[0] = make_uint4(masks[0], cnt01 & 0xffffu, masks[1], cnt01 >> 16);
            out[1] = make_uint4(masks[2], cnt23 & 0xffffu, masks[3], cnt23 >> 16);
        }
    }

    // ---- tail: the last (< U) full blocks and the partial block, one block per warp step, bounds-checked
    for (size_t g = (nfull / U) * U + warp; g < nblk; g += nwarps) {
        size_t off = (g << 9) + (size_t)lane * 16;
        uint4 v = off < n ? ld_chunk(data, off, n) : make_uint4(0, 0, 0, 0);
        uint32_t nx = 0;
        if (MODE != 0 && (STRIDE < 4 || NODD > 0)) {
            nx = __shfl_down_sync(0xffffffffu, v.x, 1);
            if (lane == 31) {
                size_t o2 = (g + 1) << 9;
                nx = o2 < n ? ld_chunk(data, o2, n).x : 0u;
            }
        }
        uint32_t cnt = newline_count16_fma(v, cnl, c80);
        bool hit = probe_chunk<STRIDE, FOLD, MODE, NODD>(v, nx, s_tab, pp, c1, c2);
        if (off >= n) hit = false;   // chunks that start at or beyond n can never be candidates
        uint32_t mask = __ballot_sync(0xffffffffu, hit);
        uint32_t total = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) meta[g] = ((unsigned long long)total << 32) | mask;
    }
}

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

// ------------------------------------------------------------------------------------------------------------
// DFA walk over one scanned block [start, lim): leading NULs are skipped and the block ends at the first later
// NUL (reference hyperscanner.c:205-217: strip loop + strlen), at '\n' (inclusive) or at lim.
// ------------------------------------------------------------------------------------------------------------
struct ByteCursor {
    const uint8_t* data;
    size_t pos, lim;
    uint32_t word;
    __device__ __forceinline__ ByteCursor(const uint8_t* d, size_t p, size_t l) : data(d), pos(p), lim(l), word(0) {
        if (p < l) word = *reinterpret_cast<const uint32_t*>(data + (p & ~(size_t)3));
    }
    __device__ __forceinline__ uint32_t get() const { return (word >> (8 * (pos & 3))) & 0xffu; }
    __device__ __forceinline__ void next() {
        pos++;
        if ((pos & 3) == 0 && pos < lim) word = *reinterpret_cast<const uint32_t*>(data + pos);
    }
};

__device__ __forceinline__ size_t skip_leading_nuls(const uint8_t* data, size_t start, size_t lim) {
    ByteCursor c(data, start, lim);
    while (c.pos < lim && c.get() == 0) c.next();
    return c.pos;
}

// end of the scanned block that starts at p0: just past the first '\n', or at the first NUL, or lim
__device__ size_t scanned_block_end(const uint8_t* data, size_t p0, size_t lim) {
    size_t e = p0;
    while (e < lim) {
        uint32_t b = data[e];
        if (b == 0) break;
        e++;
        if (b == '\n') break;
    }
    return e;
}

// simple mode: does any pattern match the block?  WITH_NFA = false keeps the (1 KiB of local memory) NFA state out of
// kernels that can never see NFA patterns (the fast path is only taken without them).
template <bool WITH_NFA>
__device__ bool block_matches(const DbView& db, const uint8_t* data, size_t start, size_t lim) {
    size_t p0 = skip_leading_nuls(data, start, lim);
    for (int g = 0; g < db.ngroups; g++) {
        const GroupDev G = db.groups[g];
        uint32_t s = 0;
        bool dead = false;
        ByteCursor c(data, p0, lim);
        while (c.pos < lim) {
            uint32_t b = c.get();
            if (b == 0) break;
            s = G.trans[s * G.stride + G.cls[b]];
            if (s >= G.first_accept) return true;
            if (s == G.dead) { dead = true; break; }
            if (b == '\n') break;
            c.next();
        }
        if (!dead) {
            s = G.trans[s * G.stride + G.eod];
            if (s >= G.first_accept) return true;
        }
    }
    if (WITH_NFA && db.nnfa) {
        const size_t e = scanned_block_end(data, p0, lim);
        for (int k = 0; k < db.nnfa; k++)
            if (nfa_scan_block(db.nfas[k], data + p0, e - p0, [](size_t) { return true; })) return true;
    }
    return false;
}

// general mode: count (out == nullptr) or write the reports of the block
__device__ uint32_t block_events(const DbView& db, const uint8_t* data, size_t start, size_t lim, uint32_t line, uint32_t pl_start,
                                 uint32_t pl_len, EventRec* out) {
    size_t p0 = skip_leading_nuls(data, start, lim);
    uint32_t k = 0;
    for (int g = 0; g < db.ngroups; g++) {
        const GroupDev G = db.groups[g];
        uint32_t s = 0;
        bool dead = false;
        ByteCursor c(data, p0, lim);
        while (c.pos < lim) {
            uint32_t b = c.get();
            if (b == 0) break;
            s = G.trans[s * G.stride + G.cls[b]];
            if (s >= G.first_accept) {
                if (out) out[k] = EventRec{line, pl_start, pl_len, (uint32_t)(c.pos - p0), G.accept_base + G.accept_of[s]};
                k++;
            }
            if (s == G.dead) { dead = true; break; }
            c.next();
            if (b == '\n') break;
        }
        if (!dead) {
            s = G.trans[s * G.stride + G.eod];
            if (s >= G.first_accept) {
                if (out) out[k] = EventRec{line, pl_start, pl_len, (uint32_t)(c.pos - p0), G.accept_base + G.accept_of[s]};
                k++;
            }
        }
    }
    if (db.nnfa) {
        const size_t e = scanned_block_end(data, p0, lim);
        for (int q = 0; q < db.nnfa; q++) {
            const uint32_t report = db.nfas[q].report;
            nfa_scan_block(db.nfas[q], data + p0, e - p0, [&](size_t end) {
                if (out) out[k] = EventRec{line, pl_start, pl_len, (uint32_t)end, report};
                k++;
                return false;
            });
        }
    }
    return k;
}

// ------------------------------------------------------------------------------------------------------------
// FAST PATH kernels
// ------------------------------------------------------------------------------------------------------------
// Flags segments that may contain a line too long for the fast path: an aligned super-block of `blocks_per_super`
// 512-byte blocks without any newline.
__global__ void k_check_long(const unsigned long long* __restrict__ prefix, size_t nblk, size_t blocks_per_super, const unsigned long long* meta_total,
                             Totals* totals) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t lo = j * blocks_per_super, hi = lo + blocks_per_super;
    if (hi > nblk) return;   // partial trailing super-block cannot hide a full one
    uint32_t a = (uint32_t)prefix[lo / kGroupBlocks];   // blocks_per_super is a multiple of kGroupBlocks
    uint32_t b = hi < nblk ? (uint32_t)prefix[hi / kGroupBlocks] : (uint32_t)*meta_total;
    if (a == b) atomicOr(&totals->flags, 1u);
}

// meta/prefix -> ordered list of candidate chunk indices
__global__ void k_list_candidates(const unsigned long long* __restrict__ meta, const unsigned long long* __restrict__ prefix, size_t nblk,
                                  uint32_t* __restrict__ cand, size_t cap, Totals* totals) {
    const size_t ngroups = (nblk + kGroupBlocks - 1) / kGroupBlocks;
    for (size_t grp = (size_t)blockIdx.x * blockDim.x + threadIdx.x; grp < ngroups; grp += (size_t)gridDim.x * blockDim.x) {
        size_t at = (size_t)(prefix[grp] >> 32);
        for (size_t g = grp * kGroupBlocks; g < nblk && g < (grp + 1) * kGroupBlocks; g++) {
            uint32_t mask = (uint32_t)meta[g];
            while (mask) {
                int b = __ffs(mask) - 1;
                mask &= mask - 1;
                if (at < cap) cand[at] = (uint32_t)(g * 32 + b);
                else atomicOr(&totals->flags, 2u);
                at++;
            }
        }
    }
}

__device__ __forceinline__ bool is_word_dev(uint32_t b) {
    return (b - '0' < 10u) || ((b | 0x20u) - 'a' < 26u) || b == '_';
}

// LOCAL verification walk of one DFA group around candidate chunk [o, o+16).
//  - starts at t (at most `lookback` bytes before the chunk, never before the line start) in the start-of-line state
//    or in the mid-line entry state that matches the previous byte;
//  - a NUL acts as end-of-data followed by a restart (lines with NULs are re-checked exactly by k_emit_simple);
//  - a '\n' ends the line: the walk continues with the next line only if that line starts inside the chunk;
//  - once past every gram hit of the chunk (idle_from: o+19, or the end of the last gram that k_verify_local found
//    again) the walk stops as soon as the automaton is idle: a match that contains a gram hit of this chunk would
//    still be in progress.
// line_bit: bit of the line that contains t (bit j = j-th line intersecting the chunk).
// Returns bit j set if the j-th line intersecting the chunk matched.
__device__ uint32_t walk_local(const GroupDev& G, const uint8_t* __restrict__ data, size_t n, size_t o, size_t t, bool at_line_start,
                               size_t idle_from, uint32_t line_bit) {
    uint32_t s = 0;
    if (!at_line_start) s = is_word_dev(data[t - 1]) ? G.mid_word : G.mid_other;
    uint32_t mask = 0;
    const size_t chunk_end = o + 16;
    const uint16_t* __restrict__ flat = G.flat;
    const uint32_t first_accept = G.first_accept, idle_end = G.idle_end;
    if (flat) {
        // Fast form: '\n' and NUL are ordinary columns of the table and "matched" is an absorbing state (see
        // engine_upload), offsets are 32-bit.  The walk advances one ALIGNED WORD per step:
        //  - a full word without a newline is four chained lookups and nothing else (no per-byte tests: a match
        //    sticks until the line ends);
        //  - a word with a newline, the first word of an unaligned start and the last word of the segment take the
        //    byte-wise form below, straight-line code without inner loops (threads of a warp diverge here, so it is short).
        // The line bit is set when the line ends in the matched state, or at the end of the walk.
        const uint32_t end = (uint32_t)n, cend = (uint32_t)chunk_end, ifrom = (uint32_t)idle_from;
        uint32_t pos = (uint32_t)t;
        if (pos >= end) return G.eod_next[s] >= first_accept ? line_bit : 0u;
        uint32_t wpos = pos & ~3u;
        uint32_t word = *reinterpret_cast<const uint32_t*>(data + wpos);   // the buffer is padded to a multiple of 16 bytes
        while (true) {
            // the next word is requested before the (dependent) table lookups of this one
            const uint32_t next_word = wpos + 4 < end ? *reinterpret_cast<const uint32_t*>(data + wpos + 4) : 0u;
            const uint32_t x = word ^ 0x0a0a0a0au;
            if (((x - 0x01010101u) & ~x & 0x80808080u) == 0 && pos == wpos && wpos + 4 <= end) {
                s = flat[(s << 8) | (word & 0xffu)];
                s = flat[(s << 8) | ((word >> 8) & 0xffu)];
                s = flat[(s << 8) | ((word >> 16) & 0xffu)];
                s = flat[(s << 8) | (word >> 24)];
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t p = wpos + k;
                    if (p >= pos && p < end) {
                        const uint32_t b = (word >> (8 * k)) & 0xffu;
                        s = flat[(s << 8) | b];
                        if (b == '\n') {
                            if (s >= first_accept) mask |= line_bit;
                            if (p + 1 >= cend) return mask;   // the next line starts outside the chunk
                            line_bit <<= 1;
                            s = 0;
                        }
                    }
                }
            }
            pos = wpos + 4;
            if (s >= first_accept) {
                if (pos >= cend) return mask | line_bit;   // matched, and no further line starts inside the chunk
            } else if (pos >= ifrom && s < idle_end) {
                return mask;
            }
            if (pos >= end) break;
            wpos = pos;
            word = next_word;
        }
        if (s >= first_accept || G.eod_next[s] >= first_accept) mask |= line_bit;
        return mask;
    }
    bool done = false;
    ByteCursor c(data, t, n);
    while (c.pos < n) {
        const uint32_t b = c.get();
        if (!done) {
            bool hit;
            if (b == 0) {
                hit = G.trans[s * G.stride + G.eod] >= first_accept;
                s = 0;
            } else {
                s = G.trans[s * G.stride + G.cls[b]];
                hit = s >= first_accept;
                if (!hit && b == '\n') hit = G.trans[s * G.stride + G.eod] >= first_accept;
            }
            if (hit) { mask |= line_bit; done = true; }
        }
        c.next();
        if (b == '\n') {
            if (c.pos >= chunk_end || c.pos >= n) return mask;
            line_bit <<= 1;
            done = false;
            s = 0;
            continue;
        }
        if (c.pos >= idle_from && (done || s < idle_end)) return mask;
        if (done && c.pos >= chunk_end) return mask;
    }
    if (!done && G.trans[s * G.stride + G.eod] >= first_accept) mask |= line_bit;
    return mask;
}

constexpr int kEmitThreads = 256;
constexpr int kEmitTile = 2048;   // candidates per emit step (and per record-offset entry): enough marked ones to keep every warp busy

// The exact gram set in global memory (two-choice table, Prefilter::confirm_keys), for k_verify_local to find the hit
// positions inside a candidate chunk again: k_stream only reports "some sampled gram of this chunk MAY be in the set".
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
};

// One thread per candidate chunk: local verification (see walk_local); writes the bitmask of matched lines.
// With the exact gram table at hand, the walk covers [first gram hit - lookback, end of the last gram hit] and then runs on
// until the automaton is idle (with one hit per chunk, the usual case, a third of walking the whole chunk), and a chunk
// that k_stream flagged only because of a bloom collision is dropped without a walk.
__global__ void __launch_bounds__(128, 16) k_verify_local(DbView db, const uint8_t* __restrict__ data, size_t n, const uint32_t* __restrict__ cand,
                                                          const unsigned long long* meta_total, size_t cap, uint32_t lookback, ReprobeParams rp,
                                                          uint32_t* __restrict__ marks, uint32_t* __restrict__ tile_records) {
    size_t ncand = (size_t)(*meta_total >> 32);
    if (ncand > cap) ncand = cap;
    // whole warps stay in the loop (a warp's 32 candidates are consecutive and lie in one emit tile): the records of the
    // tile are counted with one warp reduction and one atomic
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; (i & ~(size_t)31) < ncand; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t mask = 0;
    if (i < ncand) {
    const size_t o = (size_t)cand[i] * 16;
    size_t t;
    bool at_line_start;
    size_t idle_from = o + 19;
    uint32_t line_bit = 1u;
    uint32_t group_mask = 0xffffffffu;   // DFA groups to walk
    if (lookback == 0xffffffffu) {
        t = line_start_of(data, o);
        at_line_start = true;
    } else {
        size_t hi = o;   // the walk has to start at or before hi - lookback
        uint32_t nl_in_chunk = 0;
        if (rp.keys) {
            const uint4 v = ld_chunk(data, o, n);
            uint32_t w[5] = {v.x, v.y, v.z, v.w, o + 16 < n ? ld_chunk(data, o + 16, n).x : 0u};
            if (rp.fold) {
#pragma unroll
                for (int k = 0; k < 5; k++) w[k] |= 0x20202020u;
            }
            uint32_t hits = 0;   // bit = byte offset of a sampled gram that is in the table
            group_mask = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                for (int sft = 0; sft < 4; sft += rp.stride) {
                    const uint32_t gram = __funnelshift_r(w[k], w[k + 1], 8 * sft);
                    const uint32_t h1 = (gram * rp.mul) >> rp.shift, h2 = rp.half + ((gram * rp.mul2) >> rp.shift);
                    const uint32_t e1 = rp.keys[h1], e2 = rp.keys[h2];
                    if (e1 == gram || e2 == gram) {
                        hits |= 1u << (4 * k + sft);
                        group_mask |= rp.groups[e1 == gram ? h1 : h2];
                    }
                }
                if (rp.nodd) {
                    const uint32_t gram = __funnelshift_r(w[k], w[k + 1], 16);
                    for (int c = 0; c < rp.nodd; c++)
                        if (gram * rp.odd_mul[c] + rp.odd_add[c] == 0u) { hits |= 1u << (4 * k + 2); group_mask = 0xffffffffu; }
                }
            }
            if (hits == 0) { marks[i] = 0; goto counted; }   // a bloom collision: no gram of the set here
            const uint32_t first = __ffs(hits) - 1, last = 31 - __clz(hits);
            hi = o + first;
            idle_from = o + last + 4;
            nl_in_chunk = newline_mask16(v) & ((1u << first) - 1u);   // newlines in [o, hi)
        }
        // start: at most `lookback` bytes before the first hit, rounded down to a word, never before the line start
        size_t lo = hi > lookback ? (hi - lookback) & ~(size_t)3 : 0;
        t = lo;
        at_line_start = lo == 0;
        if (nl_in_chunk) {
            const uint32_t after = 32 - __clz(nl_in_chunk);   // offset just past the last newline before the hit
            t = o + after;
            at_line_start = true;
            line_bit = 1u << __popc(nl_in_chunk);
        } else {
            size_t p = o;   // 16-byte aligned; scan words [p-4, p) downwards for the last '\n' in [lo, o) (nothing to scan if lo >= o)
            while (p > lo) {
                uint32_t z = eq_mask4(*reinterpret_cast<const uint32_t*>(data + p - 4), 0x0a0a0a0au);
                if (p - 4 < lo) z &= ~((1u << (8 * (uint32_t)(lo - (p - 4)))) - 1u);
                if (z) {
                    t = (p - 4) + ((31 - __clz(z)) >> 3) + 1;
                    at_line_start = true;
                    break;
                }
                p -= 4;
            }
        }
    }
    for (int g = 0; g < db.ngroups; g++)
        if ((group_mask >> (g & 31)) & 1u) mask |= walk_local(db.groups[g], data, n, o, t, at_line_start, idle_from, line_bit);
    marks[i] = mask;
    }
counted:
    const uint32_t records = __reduce_add_sync(0xffffffffu, __popc(mask));
    if ((threadIdx.x & 31) == 0 && records) atomicAdd(&tile_records[i / kEmitTile], records);
    }
}

// Exclusive scan of the per-tile record counts (a few thousand entries: one block), in place; total -> *rec_total.
__global__ void __launch_bounds__(1024) k_tile_offsets(uint32_t* __restrict__ tile_records, const unsigned long long* meta_total, size_t cap,
                                                       unsigned long long* rec_total) {
    __shared__ unsigned long long s_warp[32], s_total;
    size_t ncand = (size_t)(*meta_total >> 32);
    if (ncand > cap) ncand = cap;
    const size_t ntiles = (ncand + kEmitTile - 1) / kEmitTile;
    unsigned long long running = 0;
    for (size_t base = 0; base < ntiles; base += blockDim.x) {
        const size_t k = base + threadIdx.x;
        const unsigned long long v = k < ntiles ? tile_records[k] : 0ull;
        const unsigned long long ex = block_exclusive_scan(v, s_warp, &s_total);
        if (k < ntiles) tile_records[k] = (uint32_t)(running + ex);
        running += s_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *rec_total = running;
}

// Candidates with marked lines are first compacted per block (few candidates carry a match), then one thread per
// marked candidate computes line extents, line numbers and the exact re-check of lines with NULs.
// The same line can be marked by several candidate chunks; records come out ordered by line start, so the host
// drops adjacent duplicates.
// Records of one marked candidate chunk per lane (see k_emit_simple); returns the number of valid records the lane wrote.
// Called by whole warps (`live` = this lane has a candidate): the lanes go through their marked lines round by round, and
// in every round the line extents that a lane did not settle within kEmitBound bytes are finished by the whole warp.
constexpr size_t kEmitBound = 256;
__device__ uint32_t emit_warp(const DbView& db, const uint8_t* __restrict__ data, size_t n, const uint32_t* __restrict__ cand,
                              const uint32_t* __restrict__ marks, const unsigned long long* __restrict__ meta,
                              const unsigned long long* __restrict__ prefix, bool live, size_t i, size_t at, LineRec* __restrict__ recs,
                              size_t rec_cap, Totals* totals) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t valid = 0;
    uint32_t mask = live ? marks[i] : 0u;
    const size_t o = live ? (size_t)cand[i] * 16 : 0;
    uint32_t nlm = 0;
    if (live) {
        nlm = newline_mask16(ld_chunk(data, o, n));
        if (o + 16 > n) nlm &= (1u << (n - o)) - 1u;
    }
    // line j of the chunk starts at `st` (j = 0: somewhere before the chunk, found below); `first` = still on line 0
    size_t st = 0;
    bool first = true;
    while (__any_sync(0xffffffffu, mask != 0)) {
        // skip lines of the chunk that are not marked
        while (mask != 0 && !(mask & 1u)) {
            if (!nlm) { mask = 0; break; }
            st = o + __ffs(nlm);
            nlm &= nlm - 1;
            first = false;
            mask >>= 1;
        }
        const bool work = mask != 0;
        // ---- line start (only line 0 starts before the chunk)
        bool settled = true;
        if (work && first) settled = line_start_bounded(data, o, kEmitBound, &st);
        for (uint32_t pend = __ballot_sync(0xffffffffu, work && !settled); pend; pend &= pend - 1) {
            const int src = __ffs(pend) - 1;
            const size_t found = warp_line_start(data, (size_t)__shfl_sync(0xffffffffu, (unsigned long long)st, src));
            if ((int)lane == src) st = found;
        }
        // ---- line end
        bool has_nul = false;
        size_t en = 0;
        settled = true;
        if (work) settled = line_end_bounded(data, st, n, kEmitBound, &en, &has_nul);
        for (uint32_t pend = __ballot_sync(0xffffffffu, work && !settled); pend; pend &= pend - 1) {
            const int src = __ffs(pend) - 1;
            bool more_nul = false;
            const size_t found = warp_line_end(data, (size_t)__shfl_sync(0xffffffffu, (unsigned long long)en, src), n, &more_nul);
            if ((int)lane == src) { en = found; has_nul |= more_nul; }
        }
        if (work) {
            bool ok = true;
            if (first) {
                // the line started before this chunk: an earlier candidate chunk that intersects it may have marked it
                // already (the line is the LAST line of such a chunk); only the first marking is kept.  (A repeat still
                // gets its extents and line number computed: its neighbours in the warp need that work anyway.)
                for (size_t k = i; k-- > 0;) {
                    const size_t ok_off = (size_t)cand[k] * 16;
                    if (ok_off + 16 <= st) break;
                    const uint32_t mk = marks[k];
                    if (!mk) continue;
                    uint4 pv = ld_chunk(data, ok_off, n);
                    const uint32_t last_idx = __popc(newline_mask16(pv) & 0x7fffu);   // line starts inside that chunk
                    if ((mk >> last_idx) & 1u) { ok = false; break; }
                }
            }
            if (ok && has_nul) ok = block_matches<false>(db, data, st, en);
            valid += ok ? 1u : 0u;
            // line number = newlines before the line start: whole blocks from the scan, then the part of the line's own
            // 512-byte block, counted from whichever end of the block is nearer (the block's total is in meta)
            const size_t lb = st >> 9;
            uint32_t line_no = newlines_before_block(prefix, meta, lb);
            if ((st & 511) <= 256) line_no += count_newlines(data, lb << 9, st);
            else line_no += (uint32_t)(meta[lb] >> 32) - count_newlines(data, st, min((lb + 1) << 9, n));
            if (at < rec_cap) recs[at] = LineRec{line_no, (uint32_t)st, ok ? ((uint32_t)(en - st) | (has_nul ? kHasNulBit : 0u)) : kInvalidLen};
            else atomicOr(&totals->flags, 4u);
            at++;
            // on to the next line of the chunk
            if (!nlm) mask = 0;
            else {
                st = o + __ffs(nlm);
                nlm &= nlm - 1;
                first = false;
                mask >>= 1;
            }
        }
    }
    return valid;
}

// Persistent blocks walk tiles of kEmitTile candidates.  The marked candidates of a tile (about one in six) go into a
// shared-memory queue together with their record offset (tile offset from k_tile_offsets + a block scan inside the
// tile); the block takes them out in FULL batches of one per thread and carries the remainder over to the next tile, so
// that the expensive per-record work runs with every thread busy instead of a last, mostly empty round per tile.
constexpr uint32_t kEmitQueue = 4096;   // >= kEmitTile + kEmitThreads, power of two
__global__ void __launch_bounds__(kEmitThreads) k_emit_simple(DbView db, const uint8_t* __restrict__ data, size_t n, const uint32_t* __restrict__ cand,
                                                              const uint32_t* __restrict__ marks, const uint32_t* __restrict__ tile_offsets,
                                                              const unsigned long long* __restrict__ meta, const unsigned long long* __restrict__ prefix,
                                                              const unsigned long long* meta_total, size_t cap, LineRec* __restrict__ recs, size_t rec_cap,
                                                              Totals* totals) {
    __shared__ uint32_t q_cand[kEmitQueue], q_at[kEmitQueue];
    __shared__ unsigned long long s_warp[kEmitThreads / 32], s_total;
    uint32_t valid = 0;
    uint32_t head = 0, queued = 0;   // the same in every thread of the block
    size_t ncand = (size_t)(*meta_total >> 32);
    if (ncand > cap) ncand = cap;
    constexpr int kPer = kEmitTile / kEmitThreads;   // consecutive candidates per thread in the compaction step
    for (size_t block_base = (size_t)blockIdx.x * kEmitTile; block_base < ncand; block_base += (size_t)gridDim.x * kEmitTile) {
        uint32_t mk[kPer];
        uint32_t records = 0, marked = 0;
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            const size_t i = block_base + (size_t)threadIdx.x * kPer + j;
            mk[j] = i < ncand ? marks[i] : 0u;
            records += __popc(mk[j]);
            marked += mk[j] != 0u;
        }
        // one scan for both: queue position (marked candidates before mine) and record offset (records before mine)
        const unsigned long long before = block_exclusive_scan(((unsigned long long)marked << 32) | records, s_warp, &s_total);
        uint32_t slot = head + queued + (uint32_t)(before >> 32);
        uint32_t at = tile_offsets[block_base / kEmitTile] + (uint32_t)before;
#pragma unroll
        for (int j = 0; j < kPer; j++) {
            if (mk[j]) {
                q_cand[slot & (kEmitQueue - 1)] = (uint32_t)(block_base + (size_t)threadIdx.x * kPer + j);
                q_at[slot & (kEmitQueue - 1)] = at;
                slot++;
                at += __popc(mk[j]);
            }
        }
        queued += (uint32_t)(s_total >> 32);
        __syncthreads();
        while (queued >= (uint32_t)kEmitThreads) {
            const uint32_t k = (head + threadIdx.x) & (kEmitQueue - 1);
            valid += emit_warp(db, data, n, cand, marks, meta, prefix, true, q_cand[k], q_at[k], recs, rec_cap, totals);
            head += kEmitThreads;
            queued -= kEmitThreads;
        }
        __syncthreads();   // everything taken out before the next tile overwrites queue slots / scan scratch
    }
    if (queued) {   // whole warps, some lanes without a candidate
        const uint32_t k = (head + threadIdx.x) & (kEmitQueue - 1);
        const bool live = threadIdx.x < queued;
        valid += emit_warp(db, data, n, cand, marks, meta, prefix, live, live ? q_cand[k] : 0, live ? q_at[k] : 0, recs, rec_cap, totals);
    }
    // unique valid records of the segment (count-only callers need nothing else)
    valid = __reduce_add_sync(0xffffffffu, valid);
    if ((threadIdx.x & 31) == 0 && valid) atomicAdd(&totals->aux_total, (unsigned long long)valid);
}

// Device-resident inputs: the end of segment j is the byte after a '\n' before boundary (j+1)*chunk, chosen so that the
// NEXT segment starts 16-byte aligned (the kernels use 16-byte loads): one line end in 16 qualifies on average.
// One thread per boundary scans backwards (gives up after `window` bytes -> 0 = not found).
__global__ void k_find_cuts(const uint8_t* __restrict__ data, size_t size, size_t chunk, size_t window, size_t ncuts, unsigned long long* __restrict__ cuts) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncuts) return;
    size_t b = (j + 1) * chunk;
    if (b >= size) { cuts[j] = size; return; }
    size_t lo = b > window ? b - window : 0;
    unsigned long long found = 0;
    for (size_t p = b; p > lo; p--) {
        if ((p & 15) == 0 && data[p - 1] == '\n') { found = p; break; }
    }
    cuts[j] = found;
}

// ------------------------------------------------------------------------------------------------------------
// GENERAL PATH kernels
// ------------------------------------------------------------------------------------------------------------
// warp per 512-byte block: write the offset of every '\n' at its global rank
__global__ void __launch_bounds__(256) k_newline_positions(const uint8_t* __restrict__ data, size_t n, size_t nblk, const unsigned long long* __restrict__ meta,
                                                           const unsigned long long* __restrict__ prefix, uint32_t* __restrict__ nlpos) {
    const int lane = threadIdx.x & 31;
    size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (g >= nblk) return;
    size_t off = g * 512 + (size_t)lane * 16;
    uint32_t m = 0;
    if (off < n) {
        uint4 v = ld_chunk(data, off, n);
        m = newline_mask16(v);
        if (off + 16 > n) m &= (1u << (n - off)) - 1u;
    }
    uint32_t cnt = __popc(m), incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    size_t at = (size_t)newlines_before_block(prefix, meta, g) + (incl - cnt);
    while (m) {
        int b = __ffs(m) - 1;
        m &= m - 1;
        nlpos[at++] = (uint32_t)(off + b);
    }
}

__device__ __forceinline__ void line_extent(const uint32_t* nlpos, size_t nl_total, size_t n, size_t i, uint32_t& start, uint32_t& len) {
    start = i ? nlpos[i - 1] + 1 : 0;
    uint32_t end = i < nl_total ? nlpos[i] + 1 : (uint32_t)n;
    len = end - start;
}

// pseudo-lines per line for a gzgets buffer of buffer_size (limit = buffer_size - 1 bytes per read)
__global__ void k_count_pseudo_lines(const uint32_t* __restrict__ nlpos, size_t nl_total, size_t n, size_t nlines, uint32_t limit,
                                     uint32_t* __restrict__ npl, Totals* totals) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlines) return;
    uint32_t st, len;
    line_extent(nlpos, nl_total, n, i, st, len);
    npl[i] = (len + limit - 1) / limit;
    if (len > limit) atomicMax(&totals->max_line, len);
}

__global__ void k_build_pseudo_lines(const uint32_t* __restrict__ nlpos, size_t nl_total, size_t n, size_t nlines, uint32_t limit,
                                     const unsigned long long* __restrict__ ploff, uint32_t* __restrict__ pl_start, uint32_t* __restrict__ pl_len) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlines) return;
    uint32_t st, len;
    line_extent(nlpos, nl_total, n, i, st, len);
    size_t at = ploff ? (size_t)ploff[i] : i;
    while (len > 0) {
        uint32_t take = len < limit ? len : limit;
        pl_start[at] = st;
        pl_len[at] = take;
        at++;
        st += take;
        len -= take;
    }
}

__global__ void __launch_bounds__(128) k_match_pl_simple(DbView db, const uint8_t* __restrict__ data, const uint32_t* __restrict__ pl_start,
                                                         const uint32_t* __restrict__ pl_len, size_t npl, uint8_t* __restrict__ flags) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npl) return;
    size_t st = pl_start[i];
    flags[i] = block_matches<true>(db, data, st, st + pl_len[i]) ? 1 : 0;
}

__global__ void k_emit_pl_simple(const uint32_t* __restrict__ pl_start, const uint32_t* __restrict__ pl_len, size_t npl, const uint8_t* __restrict__ flags,
                                 const unsigned long long* __restrict__ off, LineRec* __restrict__ recs) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npl || !flags[i]) return;
    recs[off[i]] = LineRec{(uint32_t)i, pl_start[i], pl_len[i]};
}

__global__ void __launch_bounds__(128) k_match_pl_events(DbView db, const uint8_t* __restrict__ data, const uint32_t* __restrict__ pl_start,
                                                         const uint32_t* __restrict__ pl_len, size_t npl, uint32_t* __restrict__ counts,
                                                         const unsigned long long* __restrict__ off, EventRec* __restrict__ out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npl) return;
    size_t st = pl_start[i];
    uint32_t len = pl_len[i];
    if (out) {
        if (counts[i]) block_events(db, data, st, st + len, (uint32_t)i, (uint32_t)st, len, out + off[i]);
    } else {
        counts[i] = block_events(db, data, st, st + len, (uint32_t)i, (uint32_t)st, len, nullptr);
    }
}

// warp per record: copy matched line bytes into a packed buffer (device-resident scans with a callback)
__global__ void k_gather_lines(const uint8_t* __restrict__ data, const uint32_t* __restrict__ starts, const uint32_t* __restrict__ lens,
                               const unsigned long long* __restrict__ outoff, size_t count, uint8_t* __restrict__ out) {
    size_t r = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= count) return;
    const uint8_t* src = data + starts[r];
    uint8_t* dst = out + outoff[r];
    uint32_t len = lens[r];
    for (uint32_t k = lane; k < len; k += 32) dst[k] = src[k];
    if (lane == 0) dst[len] = 0;
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
    ~DevBuf() { if (p) cudaFree(p); }
};
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
    ~PinBuf() { if (p) cudaFreeHost(p); }
};

int g_device = -1;
int g_num_sms = 148;
std::mutex g_mu;

}  // namespace

struct DeviceDb {
    std::shared_ptr<Database> db;
    int device = 0;
    std::vector<void*> allocs;
    GroupDev* d_groups = nullptr;
    int ngroups = 0;
    NfaView* d_nfas = nullptr;
    int nnfa = 0;
    bool simple = false;
    ~DeviceDb() { for (void* p : allocs) cudaFree(p); }
};

struct DevicePrefilter {
    int device = 0;
    uint32_t* d_table = nullptr;
    int table_words = 0;
    int stride = 4;
    bool fold = false;
    int mode = 0;        // 1 exact keys, 2 bloom bitmap
    int nodd = 0;        // register compares at offsets 2 mod 4 (mixed sampling; bloom mode only)
    ProbeParams pp{};
    uint32_t lookback = 0xffffffffu;
    uint32_t* d_confirm = nullptr;   // exact gram set for the verification kernel (Prefilter::confirm_keys), or null
    uint32_t* d_confirm_groups = nullptr;
    int confirm_log2 = 0;
    uint32_t confirm_mul = 0, confirm_mul2 = 0;
    double bloom_false_rate = 0;     // expected share of 16-byte chunks flagged by bloom collisions alone
    ~DevicePrefilter() { if (d_table) cudaFree(d_table); if (d_confirm) cudaFree(d_confirm); if (d_confirm_groups) cudaFree(d_confirm_groups); }
};

class ScanSlot {
public:
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // result D2H, so that it does not queue behind the next segment's kernels
    cudaEvent_t done = nullptr;           // all kernels of the segment + the totals copy have finished
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // 0/1: whole segment, 2/3: streaming kernel
    DevBuf d_input, d_meta, d_prefix, d_sums, d_cand, d_res, d_recoff, d_recs, d_totals;
    DevBuf d_nlpos, d_npl, d_ploff, d_plstart, d_pllen, d_flags, d_counts, d_events, d_gather, d_gidx;
    PinBuf h_totals, h_recs, h_stage, h_gather, h_probe;
    // state of the in-flight segment
    const DeviceDb* ddb = nullptr;
    const uint8_t* data = nullptr;   // device pointer of the segment
    size_t n = 0, nblk = 0, cand_cap = 0, rec_cap = 0;
    int buffer_size = 0;
    bool fast = false;
    bool want_records = true;   // false: the caller only counts matches (no callback, no limit): skip the record D2H
    SegmentStats stats;
    bool in_use = false;

    int run_general(SegmentResult& out, std::string& error);
};

int engine_select_device(int device, std::string& error) {
    std::lock_guard<std::mutex> lk(g_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        error = std::string("no CUDA device available: ") + cudaGetErrorString(e) + " (libgpugrep has no CPU fallback)";
        return 7;
    }
    if (device < 0 || device >= count) device = 0;
    CUDA_TRY(cudaSetDevice(device));
    if (g_device != device) {
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        g_num_sms = prop.multiProcessorCount;
        g_device = device;
    }
    return 0;
}

int engine_current_device() { return g_device; }

std::shared_ptr<DeviceDb> engine_upload(const std::shared_ptr<Database>& db, std::string& error) {
    static std::mutex mu;
    static std::vector<std::pair<std::weak_ptr<Database>, std::shared_ptr<DeviceDb>>> cache;
    std::lock_guard<std::mutex> lk(mu);
    int dev = 0;
    cudaGetDevice(&dev);
    for (auto it = cache.begin(); it != cache.end();) {
        auto sp = it->first.lock();
        if (!sp) { it = cache.erase(it); continue; }
        if (sp == db && it->second->device == dev) return it->second;
        ++it;
    }
    auto out = std::make_shared<DeviceDb>();
    out->db = db;
    out->device = dev;
    out->simple = db->simple;
    auto upload = [&](const void* src, size_t bytes) -> void* {
        void* p = nullptr;
        if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) return nullptr;
        out->allocs.push_back(p);
        if (bytes && cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
        return p;
    };
    std::vector<GroupDev> groups;
    for (size_t g = 0; g < db->groups.size(); g++) {
        const Dfa& d = db->groups[g].dfa;
        if (d.num_states > 65536) { error = "DFA group exceeds 65536 states"; return nullptr; }
        std::vector<uint16_t> t16(d.trans.size());
        for (size_t k = 0; k < d.trans.size(); k++) t16[k] = (uint16_t)d.trans[k];
        GroupDev G{};
        G.trans = (const uint16_t*)upload(t16.data(), t16.size() * sizeof(uint16_t));
        G.cls = (const uint8_t*)upload(d.byte_class, 256);
        G.accept_of = (const uint32_t*)upload(d.accept_of.data(), d.accept_of.size() * sizeof(uint32_t));
        if (!G.trans || !G.cls || !G.accept_of) { error = "cudaMalloc/cudaMemcpy failed while uploading DFA tables"; return nullptr; }
        if (db->simple && (size_t)d.num_states * 512 <= ((size_t)256 << 20)) {
            // Byte-indexed table for local verification with the line rules folded in (one load per byte, no special
            // cases in the walk): '\n' and NUL end the scanned block, so their columns hold either the absorbing
            // "matched" state (the block matched at its end) or state 0 (restart: next line / text after the NUL).
            std::vector<uint16_t> flat((size_t)d.num_states * 256), eod((size_t)d.num_states);
            // Every accepting state is replaced by ONE absorbing "matched" state that also survives '\n' and NUL: the
            // walk tests for it only where a line ends (see walk_local), never per byte.
            const uint16_t sink = (uint16_t)d.sink_match;
            for (int st = 0; st < d.num_states; st++) {
                const uint32_t at_eod = d.trans[(size_t)st * d.stride + d.num_classes];
                eod[st] = (uint16_t)at_eod;
                for (int b = 0; b < 256; b++) {
                    uint32_t nx = d.trans[(size_t)st * d.stride + d.byte_class[b]];
                    if (st >= d.first_accept) nx = sink;
                    else if (b == 0) nx = (int)at_eod >= d.first_accept ? sink : 0;
                    else if (b == '\n') nx = ((int)nx >= d.first_accept || (int)d.trans[(size_t)nx * d.stride + d.num_classes] >= d.first_accept) ? sink : 0;
                    else if ((int)nx >= d.first_accept) nx = sink;
                    flat[(size_t)st * 256 + b] = (uint16_t)nx;
                }
            }
            G.flat = (const uint16_t*)upload(flat.data(), flat.size() * sizeof(uint16_t));
            G.eod_next = (const uint16_t*)upload(eod.data(), eod.size() * sizeof(uint16_t));
            if (!G.flat || !G.eod_next) { error = "cudaMalloc/cudaMemcpy failed while uploading DFA tables"; return nullptr; }
        }
        G.stride = (uint32_t)d.stride;
        G.eod = (uint32_t)d.num_classes;
        G.first_accept = (uint32_t)d.first_accept;
        G.dead = d.dead >= 0 ? (uint32_t)d.dead : 0xffffffffu;
        G.idle_end = (uint32_t)d.idle_end;
        G.mid_other = (uint32_t)d.entry_mid_other;
        G.mid_word = (uint32_t)d.entry_mid_word;
        G.accept_base = db->report_begin.empty() ? 0u : 0u;
        groups.push_back(G);
    }
    // accept_base: flattened index of (group, accept set) = sum of accept-set counts of earlier groups
    uint32_t base = 0;
    for (size_t g = 0; g < groups.size(); g++) {
        groups[g].accept_base = base;
        base += (uint32_t)db->groups[g].dfa.accept_sets.size();
    }
    out->d_groups = (GroupDev*)upload(groups.data(), groups.size() * sizeof(GroupDev));
    out->ngroups = (int)groups.size();
    if (!out->d_groups) { error = "cudaMalloc failed for group table"; return nullptr; }
    std::vector<NfaView> nfas;
    for (size_t k = 0; k < db->nfas.size(); k++) {
        const NfaTables& t = db->nfas[k].tables;
        NfaView v;
        v.positions = t.positions;
        v.words = t.words;
        v.reach = (const uint32_t*)upload(t.reach.data(), t.reach.size() * 4);
        v.follow = (const uint32_t*)upload(t.follow.data(), t.follow.size() * 4);
        v.follow_match = (const uint32_t*)upload(t.follow_match.data(), t.follow_match.size() * 4);
        v.restart = (const uint32_t*)upload(t.restart.data(), t.restart.size() * 4);
        v.report = base + (uint32_t)k;   // flattened report index: after the accept sets of all DFA groups
        if (!v.reach || !v.follow || !v.follow_match || !v.restart) { error = "cudaMalloc failed for NFA tables"; return nullptr; }
        nfas.push_back(v);
    }
    out->d_nfas = (NfaView*)upload(nfas.data(), nfas.size() * sizeof(NfaView));
    out->nnfa = (int)nfas.size();
    cache.emplace_back(db, out);
    if (cache.size() > 8) cache.erase(cache.begin());
    return out;
}

std::shared_ptr<DevicePrefilter> engine_upload_prefilter(const Prefilter& pf, std::string& error) {
    if (!pf.enabled) return nullptr;
    auto out = std::make_shared<DevicePrefilter>();
    cudaGetDevice(&out->device);
    std::vector<uint32_t> replicated;
    const std::vector<uint32_t>* src = &pf.bitmap;
    const char* want = std::getenv("GPUGREP_FILTER");
    const bool use_exact = pf.exact && pf.odd.empty() && want && std::strcmp(want, "exact") == 0;
    out->stride = pf.stride;
    out->fold = pf.fold_case;
    out->mode = use_exact ? 1 : 2;
    out->pp.mul = use_exact ? pf.hash_mul : pf.bloom_mul;
    out->pp.mul2 = pf.hash_mul2;
    out->pp.shift = 32 - (pf.log2_bits - 3);   // bloom: product -> byte index
    out->lookback = pf.lookback;
    out->nodd = (int)std::min<size_t>(pf.odd.size(), 2);
    for (int k = 0; k < 2; k++) {
        const auto& c = pf.odd.empty() ? Prefilter::OddCompare{1u, 1u} : pf.odd[(size_t)k < pf.odd.size() ? (size_t)k : 0];
        out->pp.odd_mul[k] = c.mul;
        out->pp.odd_add[k] = c.add;
    }
    if (use_exact) {
        // replicate so that a lane reads copy (lane mod R): as many copies as fit ~160 KiB of shared memory, at most 32
        const size_t slots = (size_t)1 << pf.log2_slots;
        int rshift = 5;
        while (rshift > 0 && (3 * slots * 4) << rshift > 200 * 1024) rshift--;   // two halves + alignment slack of one half
        const size_t copies = (size_t)1 << rshift;
        replicated.resize(2 * slots * copies);
        for (size_t half = 0; half < 2; half++)
            for (size_t h = 0; h < slots; h++)
                for (size_t r = 0; r < copies; r++) replicated[half * slots * copies + h * copies + r] = pf.keys[half * slots + h];
        out->pp.rshift = rshift;
        out->pp.half_bytes = (uint32_t)(slots * copies * 4);
        // byte offset of slot h, copy r: ((h << rshift) | r) * 4.  h = product >> (32 - log2_slots), hence:
        out->pp.shift = 32 - pf.log2_slots - rshift - 2;
        out->pp.amask = (uint32_t)((slots - 1) << (rshift + 2));
        src = &replicated;
    }
    if (!pf.confirm_keys.empty() && pf.confirm_groups.size() == pf.confirm_keys.size()) {
        out->confirm_log2 = pf.confirm_log2;
        out->confirm_mul = pf.confirm_mul;
        out->confirm_mul2 = pf.confirm_mul2;
        if (cudaMalloc((void**)&out->d_confirm, pf.confirm_keys.size() * sizeof(uint32_t)) != cudaSuccess ||
            cudaMemcpy(out->d_confirm, pf.confirm_keys.data(), pf.confirm_keys.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMalloc((void**)&out->d_confirm_groups, pf.confirm_groups.size() * sizeof(uint32_t)) != cudaSuccess ||
            cudaMemcpy(out->d_confirm_groups, pf.confirm_groups.data(), pf.confirm_groups.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
            error = "cudaMalloc/cudaMemcpy failed for the gram confirmation table";
            return nullptr;
        }
    }
    out->bloom_false_rate = (double)pf.num_grams * (16.0 / pf.stride) / (double)((size_t)1 << pf.log2_bits);
    out->table_words = (int)src->size();
    if (cudaMalloc((void**)&out->d_table, src->size() * sizeof(uint32_t)) != cudaSuccess ||
        cudaMemcpy(out->d_table, src->data(), src->size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess) {
        error = "cudaMalloc/cudaMemcpy failed for the prefilter table";
        return nullptr;
    }
    return out;
}

namespace {
std::vector<ScanSlot*> g_free_slots;
}

ScanSlot* engine_acquire_slot(std::string& error) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { error = "cudaGetDevice failed"; return nullptr; }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        for (size_t i = 0; i < g_free_slots.size(); i++) {
            if (g_free_slots[i]->device == dev) {
                ScanSlot* s = g_free_slots[i];
                g_free_slots.erase(g_free_slots.begin() + i);
                s->in_use = true;
                return s;
            }
        }
    }
    ScanSlot* s = new ScanSlot();
    s->device = dev;
    if (cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking) != cudaSuccess) { error = "cudaStreamCreate failed"; delete s; return nullptr; }
    for (auto& e : s->ev) if (cudaEventCreate(&e) != cudaSuccess) { error = "cudaEventCreate failed"; delete s; return nullptr; }
    if (cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming) != cudaSuccess) {
        error = "cudaStreamCreate/cudaEventCreate failed"; delete s; return nullptr;
    }
    if (s->d_totals.reserve(sizeof(Totals)) != cudaSuccess || s->h_totals.reserve(sizeof(Totals)) != cudaSuccess) {
        error = "scratch allocation failed"; delete s; return nullptr;
    }
    s->in_use = true;
    return s;
}

void engine_release_slot(ScanSlot* slot) {
    if (!slot) return;
    slot->in_use = false;
    std::lock_guard<std::mutex> lk(g_mu);
    g_free_slots.push_back(slot);
}

uint8_t* slot_host_buffer(ScanSlot* slot, size_t capacity, std::string& error) {
    if (slot->h_stage.reserve(capacity + 64) != cudaSuccess) { error = "cudaHostAlloc failed for the staging buffer"; return nullptr; }
    return slot->h_stage.as<uint8_t>();
}

// Grid of a persistent (grid-stride) kernel: exactly as many blocks as can be resident at once.  A larger grid runs a
// second, partly empty wave in which the late blocks repeat the full per-block share of the work.
template <class Kernel>
static unsigned resident_grid(Kernel kernel, int block, size_t smem = 0) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem) != cudaSuccess || per_sm < 1) {
        (void)cudaGetLastError();
        per_sm = 1;
    }
    return (unsigned)(per_sm * g_num_sms);
}

template <class Load>
static void launch_scan(cudaStream_t st, Load load, size_t n, unsigned long long* out, unsigned long long* sums, unsigned long long* total,
                        SegmentStats& stats, const unsigned long long* limit = nullptr, int limit_shift = 32) {
    size_t nb = (n + kScanTile - 1) / kScanTile;
    if (nb == 0) nb = 1;
    // persistent grid: with a device-side `limit` most tiles are empty, and empty blocks are not free to schedule
    static const unsigned resident = resident_grid(k_scan_sums<Load>, kScanThreads);
    unsigned grid = (unsigned)std::min<size_t>(nb, resident);
    k_scan_sums<Load><<<grid, kScanThreads, 0, st>>>(load, n, sums, nb, limit, limit_shift);
    k_scan_top<<<1, 1024, 0, st>>>(sums, nb, total);
    k_scan_write<Load><<<grid, kScanThreads, 0, st>>>(load, n, sums, out, nb, limit, limit_shift);
    stats.launches += 3;
}

template <int STRIDE, bool FOLD, int MODE, int NODD = 0>
static cudaError_t launch_stream_t(cudaStream_t st, int grid, int block, size_t smem, const uint8_t* data, size_t n, unsigned long long* meta,
                                   const DevicePrefilter* pf) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_stream<STRIDE, FOLD, MODE, NODD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    k_stream<STRIDE, FOLD, MODE, NODD><<<grid, block, smem, st>>>(data, n, meta, pf ? pf->d_table : nullptr, pf ? pf->table_words : 0,
                                                            pf ? pf->pp : ProbeParams{});
    return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_stream_m(cudaStream_t st, int grid, int block, size_t smem, const uint8_t* data, size_t n, unsigned long long* meta,
                                   const DevicePrefilter* pf) {
    int key = pf->stride * 2 + (pf->fold ? 1 : 0);
    if (MODE == 2 && pf->stride == 4 && pf->nodd > 0)
        return pf->fold ? launch_stream_t<4, true, 2, 2>(st, grid, block, smem, data, n, meta, pf) : launch_stream_t<4, false, 2, 2>(st, grid, block, smem, data, n, meta, pf);
    switch (key) {
        case 8: return launch_stream_t<4, false, MODE>(st, grid, block, smem, data, n, meta, pf);
        case 9: return launch_stream_t<4, true, MODE>(st, grid, block, smem, data, n, meta, pf);
        case 4: return launch_stream_t<2, false, MODE>(st, grid, block, smem, data, n, meta, pf);
        case 5: return launch_stream_t<2, true, MODE>(st, grid, block, smem, data, n, meta, pf);
        case 2: return launch_stream_t<1, false, MODE>(st, grid, block, smem, data, n, meta, pf);
        default: return launch_stream_t<1, true, MODE>(st, grid, block, smem, data, n, meta, pf);
    }
}

int slot_submit(ScanSlot* s, const DeviceDb& ddb, const DevicePrefilter* pf, const uint8_t* host_data, const uint8_t* dev_data, size_t n,
                int buffer_size, void* user_stream, std::string& error) {
    if (n >= ((size_t)1 << 32) - 1024) { error = "segment too large"; return 7; }
    s->stream = user_stream ? (cudaStream_t)user_stream : s->own_stream;
    cudaStream_t st = s->stream;
    s->ddb = &ddb;
    s->n = n;
    s->buffer_size = buffer_size;
    s->stats = SegmentStats();
    s->nblk = (n + 511) / 512;
    // fast path: simple mode + prefilter + buffer large enough that "a super-block without newline" is a cheap
    // sufficient test for "no line needs gzgets splitting"
    size_t super_bytes = 0;
    if (buffer_size >= 4096) {
        super_bytes = 2048;
        while (super_bytes * 4 <= (size_t)buffer_size && super_bytes < 65536) super_bytes *= 2;   // 2*super-1 <= buffer_size-1
    }
    s->fast = ddb.simple && ddb.nnfa == 0 && pf != nullptr && super_bytes >= 2048 && std::getenv("GPUGREP_FORCE_GENERAL") == nullptr;

    if (host_data) {
        if (s->d_input.reserve(n + 1024) != cudaSuccess) { error = "cudaMalloc failed for the input segment"; return 3; }
        CUDA_TRY(cudaMemcpyAsync(s->d_input.p, host_data, n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaMemsetAsync(s->d_input.as<uint8_t>() + n, 0, 1024, st));
        s->data = s->d_input.as<uint8_t>();
        s->stats.h2d_bytes += n;
    } else {
        if (((uintptr_t)dev_data & 15) != 0) {
            // unaligned device segment (no aligned line end was available for the cut): stage it once, device to device
            if (s->d_input.reserve(n + 1024) != cudaSuccess) { error = "cudaMalloc failed for the input segment"; return 3; }
            CUDA_TRY(cudaMemcpyAsync(s->d_input.p, dev_data, n, cudaMemcpyDeviceToDevice, st));
            CUDA_TRY(cudaMemsetAsync(s->d_input.as<uint8_t>() + n, 0, 1024, st));
            s->data = s->d_input.as<uint8_t>();
        } else {
            s->data = dev_data;
        }
    }
    s->cand_cap = n / 64 + 4096;
    s->rec_cap = n / 48 + 4096;
    size_t nb_scan = (std::max(s->nblk, s->cand_cap) + kScanTile - 1) / kScanTile + 1;
    if (s->d_meta.reserve((s->nblk + 8) * 8) != cudaSuccess || s->d_prefix.reserve((s->nblk + 8) * 8) != cudaSuccess ||
        s->d_sums.reserve(nb_scan * 8) != cudaSuccess) { error = "cudaMalloc failed for scan scratch"; return 3; }
    if (s->fast) {
        if (s->d_cand.reserve(s->cand_cap * 4) != cudaSuccess || s->d_res.reserve(s->cand_cap * sizeof(uint32_t)) != cudaSuccess ||
            s->d_recoff.reserve((s->cand_cap / kEmitTile + 2) * sizeof(uint32_t)) != cudaSuccess || s->d_recs.reserve(s->rec_cap * sizeof(LineRec)) != cudaSuccess) {
            error = "cudaMalloc failed for candidate scratch"; return 3;
        }
    }
    Totals* dT = s->d_totals.as<Totals>();
    CUDA_TRY(cudaMemsetAsync(dT, 0, sizeof(Totals), st));
    CUDA_TRY(cudaEventRecord(s->ev[0], st));
    if (n == 0) {
        CUDA_TRY(cudaEventRecord(s->ev[1], st));
        CUDA_TRY(cudaEventRecord(s->done, st));
        return 0;
    }
    // ---- K1 ----
    // persistent grid: enough CTAs to fill every SM, each warp strides over groups of kStreamU blocks
    size_t smem = s->fast ? (size_t)pf->table_words * 4 + (pf->mode == 1 ? pf->pp.half_bytes : 0) : 0;
    int block = smem > 32 * 1024 ? 1024 : 256;
    int ctas_per_sm = smem > 100 * 1024 ? 1 : (smem > 32 * 1024 ? 2 : 6);
    int grid = g_num_sms * ctas_per_sm;
    size_t groups = (s->nblk + kStreamU - 1) / kStreamU;
    size_t max_grid = (groups + (block / 32) - 1) / (block / 32);
    if ((size_t)grid > max_grid) grid = (int)std::max<size_t>(1, max_grid);
    unsigned long long* meta = s->d_meta.as<unsigned long long>();
    CUDA_TRY(cudaEventRecord(s->ev[2], st));
    cudaError_t le;
    if (s->fast) le = pf->mode == 1 ? launch_stream_m<1>(st, grid, block, smem, s->data, n, meta, pf) : launch_stream_m<2>(st, grid, block, smem, s->data, n, meta, pf);
    else le = launch_stream_t<4, false, 0>(st, grid, block, 0, s->data, n, meta, nullptr);
    if (le != cudaSuccess) { error = std::string("k_stream launch: ") + cudaGetErrorString(le); return 7; }
    CUDA_TRY(cudaEventRecord(s->ev[3], st));
    s->stats.launches++;
    s->stats.stream_launches++;
    // ---- scan of (candidates, newlines) ----
    unsigned long long* prefix = s->d_prefix.as<unsigned long long>();
    launch_scan(st, LoadMetaGroup{meta, s->nblk}, (s->nblk + kGroupBlocks - 1) / kGroupBlocks, prefix, s->d_sums.as<unsigned long long>(),
                &dT->meta_total, s->stats);
    if (s->fast) {
        size_t bps = super_bytes / 512;
        size_t nsuper = s->nblk / bps;
        if (nsuper) {
            k_check_long<<<(unsigned)((nsuper + 255) / 256), 256, 0, st>>>(prefix, s->nblk, bps, &dT->meta_total, dT);
            s->stats.launches++;
        }
        static const unsigned list_resident = resident_grid(k_list_candidates, 256);
        k_list_candidates<<<(unsigned)std::min<size_t>((s->nblk + 255) / 256, list_resident), 256, 0, st>>>(meta, prefix, s->nblk, s->d_cand.as<uint32_t>(), s->cand_cap, dT);
        DbView view{ddb.d_groups, ddb.ngroups, ddb.d_nfas, ddb.nnfa};
        static const unsigned verify_resident = resident_grid(k_verify_local, 128);
        unsigned vgrid = (unsigned)std::min<size_t>((s->cand_cap + 127) / 128, verify_resident);
        // Finding the hits again costs two loads per sampled gram and candidate.  It pays when bloom collisions flag a
        // noticeable share of chunks (large gram sets: those candidates are dropped without a walk) and when several DFA
        // groups would each walk the whole chunk (measured with 32 patterns / 1 group / 415 grams: no gain, so not there).
        ReprobeParams rp{};
        const bool want_reprobe = ddb.ngroups >= 2 || pf->bloom_false_rate > 0.005;
        if (pf->mode == 2 && pf->d_confirm && want_reprobe && std::getenv("GPUGREP_NO_REPROBE") == nullptr) {
            rp.keys = pf->d_confirm;
            rp.groups = pf->d_confirm_groups;
            rp.mul = pf->confirm_mul; rp.mul2 = pf->confirm_mul2; rp.shift = 32 - pf->confirm_log2; rp.half = 1u << pf->confirm_log2;
            rp.stride = pf->stride; rp.fold = pf->fold ? 1 : 0;
            rp.nodd = pf->nodd;
            for (int k = 0; k < 2; k++) { rp.odd_mul[k] = pf->pp.odd_mul[k]; rp.odd_add[k] = pf->pp.odd_add[k]; }
        }
        // record offsets per emit tile of kEmitTile candidates: counted by the verification kernel, scanned by one block
        uint32_t* tile_records = s->d_recoff.as<uint32_t>();
        CUDA_TRY(cudaMemsetAsync(tile_records, 0, (s->cand_cap / kEmitTile + 2) * sizeof(uint32_t), st));
        k_verify_local<<<vgrid, 128, 0, st>>>(view, s->data, n, s->d_cand.as<uint32_t>(), &dT->meta_total, s->cand_cap, pf->lookback, rp,
                                              s->d_res.as<uint32_t>(), tile_records);
        k_tile_offsets<<<1, 1024, 0, st>>>(tile_records, &dT->meta_total, s->cand_cap, &dT->rec_total);
        static const unsigned emit_resident = resident_grid(k_emit_simple, kEmitThreads);
        k_emit_simple<<<(unsigned)std::min<size_t>((s->cand_cap + kEmitTile - 1) / kEmitTile, emit_resident), kEmitThreads, 0, st>>>(
            view, s->data, n, s->d_cand.as<uint32_t>(), s->d_res.as<uint32_t>(), tile_records, meta, prefix, &dT->meta_total,
            s->cand_cap, s->d_recs.as<LineRec>(), s->rec_cap, dT);
        s->stats.launches += 3;
        CUDA_TRY(cudaEventRecord(s->ev[1], st));
    }
    CUDA_TRY(cudaMemcpyAsync(&dT->last_byte, s->data + n - 1, 1, cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(s->h_totals.p, dT, sizeof(Totals), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaEventRecord(s->done, st));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int ScanSlot::run_general(SegmentResult& out, std::string& error) {
    cudaStream_t st = stream;
    Totals* dT = d_totals.as<Totals>();
    Totals* hT = h_totals.as<Totals>();
    stats.path |= 2;
    const size_t nl_total = (size_t)(uint32_t)hT->meta_total;
    const bool trailing = n > 0 && (hT->last_byte & 0xff) != '\n';
    const size_t nlines = nl_total + (trailing ? 1 : 0);
    unsigned long long* prefix = d_prefix.as<unsigned long long>();
    if (d_nlpos.reserve((nl_total + 1) * 4) != cudaSuccess) { error = "cudaMalloc failed for newline index"; return 3; }
    if (nblk) {
        k_newline_positions<<<(unsigned)((nblk * 32 + 255) / 256), 256, 0, st>>>(data, n, nblk, d_meta.as<unsigned long long>(), prefix, d_nlpos.as<uint32_t>());
        stats.launches++;
    }
    const uint32_t limit = (uint32_t)std::max(1, buffer_size - 1);
    size_t npl_total = nlines;
    bool split = false;
    if (nlines) {
        if (d_npl.reserve(nlines * 4) != cudaSuccess) { error = "cudaMalloc failed"; return 3; }
        k_count_pseudo_lines<<<(unsigned)((nlines + 255) / 256), 256, 0, st>>>(d_nlpos.as<uint32_t>(), nl_total, n, nlines, limit, d_npl.as<uint32_t>(), dT);
        stats.launches++;
        CUDA_TRY(cudaMemcpyAsync(hT, dT, sizeof(Totals), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        split = hT->max_line > limit;
        if (split) {
            size_t nb_scan = (nlines + kScanTile - 1) / kScanTile + 1;
            if (d_ploff.reserve(nlines * 8) != cudaSuccess || d_sums.reserve(nb_scan * 8) != cudaSuccess) { error = "cudaMalloc failed"; return 3; }
            launch_scan(st, LoadU32{d_npl.as<uint32_t>()}, nlines, d_ploff.as<unsigned long long>(), d_sums.as<unsigned long long>(), &dT->aux_total, stats);
            CUDA_TRY(cudaMemcpyAsync(hT, dT, sizeof(Totals), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            npl_total = (size_t)hT->aux_total;
        }
        if (d_plstart.reserve(npl_total * 4) != cudaSuccess || d_pllen.reserve(npl_total * 4) != cudaSuccess) { error = "cudaMalloc failed"; return 3; }
        k_build_pseudo_lines<<<(unsigned)((nlines + 255) / 256), 256, 0, st>>>(d_nlpos.as<uint32_t>(), nl_total, n, nlines, limit,
                                                                                 split ? d_ploff.as<unsigned long long>() : nullptr,
                                                                                 d_plstart.as<uint32_t>(), d_pllen.as<uint32_t>());
        stats.launches++;
    }
    out.num_lines = npl_total;
    out.lines = nullptr; out.num_line_recs = 0; out.events = nullptr; out.num_events = 0;
    if (npl_total == 0) {
        CUDA_TRY(cudaEventRecord(ev[1], st));
        CUDA_TRY(cudaStreamSynchronize(st));
        return 0;
    }
    DbView view{ddb->d_groups, ddb->ngroups, ddb->d_nfas, ddb->nnfa};
    size_t nb_scan = (npl_total + kScanTile - 1) / kScanTile + 1;
    if (d_sums.reserve(nb_scan * 8) != cudaSuccess || d_recoff.reserve(npl_total * 8) != cudaSuccess) { error = "cudaMalloc failed"; return 3; }
    unsigned mgrid = (unsigned)((npl_total + 127) / 128);
    if (ddb->simple) {
        if (d_flags.reserve(npl_total) != cudaSuccess) { error = "cudaMalloc failed"; return 3; }
        k_match_pl_simple<<<mgrid, 128, 0, st>>>(view, data, d_plstart.as<uint32_t>(), d_pllen.as<uint32_t>(), npl_total, d_flags.as<uint8_t>());
        stats.launches++;
        launch_scan(st, LoadU8{d_flags.as<uint8_t>()}, npl_total, d_recoff.as<unsigned long long>(), d_sums.as<unsigned long long>(), &dT->rec_total, stats);
        CUDA_TRY(cudaMemcpyAsync(hT, dT, sizeof(Totals), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        size_t nrec = (size_t)hT->rec_total;
        if (nrec) {
            if (d_recs.reserve(nrec * sizeof(LineRec)) != cudaSuccess) { error = "cudaMalloc failed"; return 3; }
            if (h_recs.reserve(nrec * sizeof(LineRec)) != cudaSuccess) { error = "cudaHostAlloc failed"; return 3; }
            k_emit_pl_simple<<<(unsigned)((npl_total + 255) / 256), 256, 0, st>>>(d_plstart.as<uint32_t>(), d_pllen.as<uint32_t>(), npl_total,
                                                                                    d_flags.as<uint8_t>(), d_recoff.as<unsigned long long>(), d_recs.as<LineRec>());
            stats.launches++;
            CUDA_TRY(cudaEventRecord(ev[1], st));
            CUDA_TRY(cudaMemcpyAsync(h_recs.p, d_recs.p, nrec * sizeof(LineRec), cudaMemcpyDeviceToHost, st));
            stats.d2h_bytes += nrec * sizeof(LineRec);
        } else {
            CUDA_TRY(cudaEventRecord(ev[1], st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));
        out.lines = h_recs.as<LineRec>();
        out.num_line_recs = nrec;
    } else {
        if (d_counts.reserve(npl_total * 4) != cudaSuccess) { error = "cudaMalloc failed"; return 3; }
        k_match_pl_events<<<mgrid, 128, 0, st>>>(view, data, d_plstart.as<uint32_t>(), d_pllen.as<uint32_t>(), npl_total, d_counts.as<uint32_t>(), nullptr, nullptr);
        stats.launches++;
        launch_scan(st, LoadU32{d_counts.as<uint32_t>()}, npl_total, d_recoff.as<unsigned long long>(), d_sums.as<unsigned long long>(), &dT->rec_total, stats);
        CUDA_TRY(cudaMemcpyAsync(hT, dT, sizeof(Totals), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        size_t nev = (size_t)hT->rec_total;
        if (nev > ((size_t)1 << 27)) { error = "too many match events in one segment (non-SINGLEMATCH pattern matching nearly every byte?)"; return 7; }
        if (nev) {
            if (d_events.reserve(nev * sizeof(EventRec)) != cudaSuccess) { error = "cudaMalloc failed"; return 3; }
            if (h_recs.reserve(nev * sizeof(EventRec)) != cudaSuccess) { error = "cudaHostAlloc failed"; return 3; }
            k_match_pl_events<<<mgrid, 128, 0, st>>>(view, data, d_plstart.as<uint32_t>(), d_pllen.as<uint32_t>(), npl_total, d_counts.as<uint32_t>(),
                                                     d_recoff.as<unsigned long long>(), d_events.as<EventRec>());
            stats.launches++;
            CUDA_TRY(cudaEventRecord(ev[1], st));
            CUDA_TRY(cudaMemcpyAsync(h_recs.p, d_events.p, nev * sizeof(EventRec), cudaMemcpyDeviceToHost, st));
            stats.d2h_bytes += nev * sizeof(EventRec);
        } else {
            CUDA_TRY(cudaEventRecord(ev[1], st));
        }
        CUDA_TRY(cudaStreamSynchronize(st));
        out.events = h_recs.as<EventRec>();
        out.num_events = nev;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int slot_collect(ScanSlot* s, SegmentResult& out, std::string& error) {
    cudaStream_t st = s->stream;
    out = SegmentResult();
    // wait for THIS segment only: the stream may already hold the next segment's kernels
    CUDA_TRY(cudaEventSynchronize(s->done));
    s->stats.d2h_bytes += sizeof(Totals);
    if (s->n == 0) { out.stats = s->stats; return 0; }
    Totals* hT = s->h_totals.as<Totals>();
    bool done = false;
    if (s->fast) {
        s->stats.candidates = hT->meta_total >> 32;
        if (hT->flags == 0) {
            s->stats.path |= 1;
            size_t nrec = (size_t)hT->rec_total;   // includes records marked kInvalidLen (repeats of a line, failed NUL re-checks)
            out.num_valid_recs = (size_t)hT->aux_total;
            const size_t nl_total = (size_t)(uint32_t)hT->meta_total;
            out.num_lines = nl_total + ((hT->last_byte & 0xff) != '\n' ? 1 : 0);
            if (nrec && s->want_records) {
                if (s->h_recs.reserve(nrec * sizeof(LineRec)) != cudaSuccess) { error = "cudaHostAlloc failed"; return 3; }
                CUDA_TRY(cudaMemcpyAsync(s->h_recs.p, s->d_recs.p, nrec * sizeof(LineRec), cudaMemcpyDeviceToHost, s->copy_stream));
                CUDA_TRY(cudaStreamSynchronize(s->copy_stream));
                s->stats.d2h_bytes += nrec * sizeof(LineRec);
            }
            out.lines = s->h_recs.as<LineRec>();
            out.num_line_recs = nrec;
            done = true;
        }
    }
    if (!done) {
        int rc = s->run_general(out, error);
        if (rc) return rc;
    }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]) == cudaSuccess) s->stats.gpu_ms = ms;
    if (cudaEventElapsedTime(&ms, s->ev[2], s->ev[3]) == cudaSuccess) s->stats.stream_ms = ms;
    out.stats = s->stats;
    return 0;
}

int slot_probe_input(ScanSlot* s, const uint8_t* dev_data, size_t size, size_t chunk, std::vector<size_t>& cuts, uint8_t* head, size_t head_len,
                     std::string& error) {
    cuts.clear();
    if (size == 0) return 0;
    const size_t ncuts = chunk && size > chunk ? (size + chunk - 1) / chunk : 0;
    head_len = std::min(head_len, size);
    if (s->h_probe.reserve(ncuts * 8 + head_len + 16) != cudaSuccess || (ncuts && s->d_sums.reserve(ncuts * 8) != cudaSuccess)) {
        error = "scratch allocation failed";
        return 3;
    }
    cudaStream_t st = s->own_stream;
    unsigned long long* h_cuts = s->h_probe.as<unsigned long long>();
    uint8_t* h_head = s->h_probe.as<uint8_t>() + ncuts * 8;
    if (ncuts) {
        k_find_cuts<<<(unsigned)((ncuts + 63) / 64), 64, 0, st>>>(dev_data, size, chunk, (size_t)4 << 20, ncuts, s->d_sums.as<unsigned long long>());
        CUDA_TRY(cudaMemcpyAsync(h_cuts, s->d_sums.p, ncuts * 8, cudaMemcpyDeviceToHost, st));
    }
    if (head_len) CUDA_TRY(cudaMemcpyAsync(h_head, dev_data, head_len, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (head_len) std::memcpy(head, h_head, head_len);
    size_t prev = 0;
    for (size_t j = 0; j < ncuts; j++) {
        if (h_cuts[j] == 0 || h_cuts[j] <= prev) { cuts.clear(); break; }   // no newline near a boundary: caller falls back
        cuts.push_back((size_t)h_cuts[j]);
        prev = (size_t)h_cuts[j];
    }
    return 0;
}

int slot_gather_lines(ScanSlot* s, const uint32_t* starts, const uint32_t* lens, size_t count, uint8_t* out, std::string& error) {
    if (count == 0) return 0;
    cudaStream_t st = s->stream;
    // offsets on the host (small), then one gather kernel and one D2H
    std::vector<unsigned long long> off(count);
    unsigned long long total = 0;
    for (size_t i = 0; i < count; i++) { off[i] = total; total += (unsigned long long)lens[i] + 1; }
    if (s->d_gidx.reserve(count * 16) != cudaSuccess || s->d_gather.reserve(total) != cudaSuccess || s->h_gather.reserve(total) != cudaSuccess) {
        error = "allocation failed in gather"; return 3;
    }
    uint32_t* d_starts = s->d_gidx.as<uint32_t>();
    uint32_t* d_lens = d_starts + count;
    unsigned long long* d_off = reinterpret_cast<unsigned long long*>(d_starts + 2 * count);
    CUDA_TRY(cudaMemcpyAsync(d_starts, starts, count * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_lens, lens, count * 4, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_off, off.data(), count * 8, cudaMemcpyHostToDevice, st));
    k_gather_lines<<<(unsigned)((count * 32 + 255) / 256), 256, 0, st>>>(s->data, d_starts, d_lens, d_off, count, s->d_gather.as<uint8_t>());
    CUDA_TRY(cudaMemcpyAsync(s->h_gather.p, s->d_gather.p, total, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    std::memcpy(out, s->h_gather.p, total);
    s->stats.launches++;
    s->stats.d2h_bytes += total;
    return 0;
}

}  // namespace gpugrep
