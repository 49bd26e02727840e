// Exact DFA / NFA walk over one scanned block (pseudo-line): used by the general path and by the NUL re-check.
// Part of the CUDA engine (engine.cu includes these files in this order; they form one translation unit).
#pragma once

namespace gpugrep {

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
        if (G.flat) {
            // byte-indexed table with an absorbing "matched" state and the end-of-line rule folded into the '\n' column
            // (engine_upload): one load per byte, and aligned words without '\n' / NUL take four chained loads and one test
            const uint16_t* __restrict__ flat = G.flat;
            const uint32_t fa = G.first_accept;
            size_t pos = p0;
            bool open = true;   // no '\n' / NUL seen yet: the block ends at lim and needs the end-of-data transition
            while (pos < lim) {
                if ((pos & 3) == 0 && pos + 4 <= lim) {
                    const uint32_t word = *reinterpret_cast<const uint32_t*>(data + pos);
                    if ((haszero4(word) | haszero4(word ^ 0x0a0a0a0au)) == 0) {
                        s = flat[(s << 8) | (word & 0xffu)];
                        s = flat[(s << 8) | ((word >> 8) & 0xffu)];
                        s = flat[(s << 8) | ((word >> 16) & 0xffu)];
                        s = flat[(s << 8) | (word >> 24)];
                        pos += 4;
                        if (s >= fa) return true;
                        if (s == G.dead) { open = false; break; }
                        continue;
                    }
                }
                const uint32_t b = data[pos];
                if (b == 0) break;   // end of the block: end-of-data transition below
                s = flat[(s << 8) | b];
                pos++;
                if (s >= fa) return true;
                if (b == '\n' || s == G.dead) { open = false; break; }
            }
            if (open && G.eod_next[s] >= fa) return true;
            continue;
        }
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

// simple mode, NFA-fallback patterns only: does one of them match the block [start, lim)?
__device__ bool block_matches_nfa(const DbView& db, const uint8_t* data, size_t start, size_t lim) {
    const size_t p0 = skip_leading_nuls(data, start, lim);
    const size_t e = scanned_block_end(data, p0, lim);
    for (int k = 0; k < db.nnfa; k++)
        if (nfa_scan_block(db.nfas[k], data + p0, e - p0, [](size_t) { return true; })) return true;
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

}  // namespace gpugrep
