// Literal-factor prefilter: the role Hyperscan's FDR/Teddy literal matchers play in front of its automata,
// re-designed for a SIMT machine with no byte shuffle: every pattern contributes a REQUIRED FACTOR (a set of
// short class-strings one of which occurs in every match); all 4-byte windows of those factors that can land
// on a sampled text position are inserted into a hashed bitmap that lives in shared memory.  The streaming
// kernel hashes one 4-byte gram per sampled position (every 4th, 2nd or every byte, depending on the shortest
// factor) and flags 16-byte chunks whose grams hit the bitmap; only lines touching flagged chunks are walked
// by the DFA.  The filter is a superset filter: it may flag lines that do not match, never the reverse.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "regex.hpp"

namespace gpugrep {

using ClassString = std::vector<ByteSet>;

struct Prefilter {
    bool enabled = false;
    int stride = 4;             // text positions sampled: multiples of `stride` (4, 2 or 1)
    bool fold_case = false;     // text bytes are OR-ed with 0x20 before hashing (grams stored folded)
    uint32_t hash_mul = 0x9E3779B1u;
    int log2_bits = 13;         // bitmap holds 1 << log2_bits bits
    std::vector<uint32_t> bitmap;
    size_t num_grams = 0;
    int min_factor_len = 0;
    std::string note;           // why it is disabled, or a one-line summary
};

// Required-factor analysis of one pattern.  Returns false if no usable factor exists.
bool extract_factor(const Node& ast, std::vector<ClassString>& alternatives);

// Build the shared prefilter of a pattern set (one entry per pattern, in order).
void build_prefilter(const std::vector<const Node*>& asts, const std::vector<unsigned>& flags, Prefilter& out);

inline uint32_t prefilter_hash(uint32_t gram, uint32_t mul, int log2_bits) { return (gram * mul) >> (32 - log2_bits); }

}  // namespace gpugrep
