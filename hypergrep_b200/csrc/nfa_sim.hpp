// Bit-parallel position-automaton (Glushkov-style) simulation: the fallback for a pattern whose own DFA exceeds the
// state budget (e.g. `a.{200}b`), where Hyperscan would use its NFA engines.  One bit per byte-consuming position;
// per input byte:  active = (current | restarts) & reach[byte];  next = OR of follow[p] over active positions.
// Assertions (^ $ \b \B \A \z) are folded into the tables: follow/restart sets are stored per combination of
// (class of the byte just consumed) x (class of the next byte), which is all an assertion can observe.
// The same inline function runs on the device (general path kernels) and in the CPU test build (mock engine).
#pragma once
#include <cstddef>
#include <cstdint>

#ifdef __CUDACC__
#define GPUGREP_HD __host__ __device__
#else
#define GPUGREP_HD
#endif

namespace gpugrep {

constexpr int kNfaMaxWords = 128;   // up to 4096 positions per NFA pattern

// byte kinds an assertion can tell apart
enum : int { kKindWord = 0, kKindOther = 1, kKindNewline = 2, kKindEod = 3, kKindStart = 3 };

struct NfaView {
    int positions = 0;
    int words = 0;                    // ceil(positions / 32)
    const uint32_t* reach = nullptr;  // [256][words]: positions that consume byte b
    // [positions][3 prev kinds (the byte just consumed: word/other/newline)][4 next kinds][words]
    const uint32_t* follow = nullptr;
    // [positions][3][4] (one word each): non-zero if the pattern can end right after this position
    const uint32_t* follow_match = nullptr;
    // [4 prev kinds (word/other/newline/start-of-block)][4 kinds of the current byte][words]: positions where a match may begin
    const uint32_t* restart = nullptr;
    uint32_t report = 0;              // flattened report index (general mode)
};

GPUGREP_HD inline int nfa_byte_kind(uint32_t b) {
    if (b == '\n') return kKindNewline;
    bool word = (b - '0' < 10u) || ((b | 0x20u) - 'a' < 26u) || b == '_';
    return word ? kKindWord : kKindOther;
}

// Scans text[0, len) (already stripped of leading NULs and cut at the first NUL; the trailing '\n' included).
// Calls on_end(end_offset) for every offset at which a match ends; stops early if on_end returns true.
// Returns true if on_end asked to stop.
template <class OnEnd>
GPUGREP_HD inline bool nfa_scan_block(const NfaView& n, const uint8_t* text, size_t len, OnEnd on_end) {
    uint32_t cur[kNfaMaxWords], nxt[kNfaMaxWords];
    const int W = n.words;
    for (int w = 0; w < W; w++) cur[w] = 0;
    int prev_kind = kKindStart;
    for (size_t i = 0; i < len; i++) {
        const uint32_t b = text[i];
        const int kind = nfa_byte_kind(b);
        const int next_kind = i + 1 < len ? nfa_byte_kind(text[i + 1]) : kKindEod;
        const uint32_t* reach = n.reach + (size_t)b * W;
        const uint32_t* restart = n.restart + ((size_t)prev_kind * 4 + kind) * W;
        for (int w = 0; w < W; w++) nxt[w] = 0;
        bool matched = false;
        const size_t combo = (size_t)kind * 4 + next_kind;
        for (int w = 0; w < W; w++) {
            uint32_t active = (cur[w] | restart[w]) & reach[w];
            while (active) {
#ifdef __CUDA_ARCH__
                const int bit = __ffs(active) - 1;
#else
                const int bit = __builtin_ctz(active);
#endif
                active &= active - 1;
                const size_t p = (size_t)w * 32 + bit;
                const uint32_t* f = n.follow + (p * 12 + combo) * W;
                for (int v = 0; v < W; v++) nxt[v] |= f[v];
                matched |= n.follow_match[p * 12 + combo] != 0;
            }
        }
        for (int w = 0; w < W; w++) cur[w] = nxt[w];
        if (matched && on_end(i + 1)) return true;
        prev_kind = kind;
    }
    return false;
}

}  // namespace gpugrep
