// SWAR byte tests on 32-bit words / 16-byte chunks, line-extent searches (per thread and warp-cooperative), newline counting.
// Part of the CUDA engine (engine.cu includes these files in this order; they form one translation unit).
#pragma once

namespace gpugrep {

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

// The line-extent searches below run one thread per line, 32 different lines per warp (verification with an unbounded
// look-back, NFA checks; the emit kernel uses the newline-chunk masks instead).  They are written as "a tight loop that
// only skips chunks without a hit, then the exact (expensive) look at the chunk that stopped the loop": the threads of a
// warp leave the loop at different iterations but meet again behind it, so the expensive part runs once per warp with
// every lane active instead of once per iteration with one or two lanes.
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

}  // namespace gpugrep
