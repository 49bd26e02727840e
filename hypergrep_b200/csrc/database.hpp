// Compiled pattern database: what hs_compile_multi() produces in the reference (hyperscanner.c:126-142),
// rebuilt as DFA groups + literal prefilter tables laid out for the CUDA engine.
#pragma once
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "automata.hpp"
#include "prefilter.hpp"

namespace gpugrep {

struct PatternInfo {
    std::string source;
    unsigned flags = 0;
    unsigned id = 0;
};

struct DfaGroup {
    Dfa dfa;
    std::vector<int> members;   // group-local pattern index -> database pattern index
};

// One (id, singlematch) report descriptor; accept sets of every group are flattened into lists of these.
struct ReportDesc {
    unsigned id;
    unsigned singlematch;
};

// A pattern whose own DFA exceeds the state budget: simulated as a bit-parallel NFA (general path only).
struct NfaPattern {
    int pattern = 0;          // index into Database::patterns
    NfaTables tables;
    uint32_t report_begin = 0;   // its single report: Database::reports[report_begin]
};

struct Database {
    std::vector<PatternInfo> patterns;
    std::vector<NfaPattern> nfas;
    // simple: every pattern has SINGLEMATCH and all share one id -> "does any pattern match this line?"
    bool simple = false;
    unsigned simple_id = 0;
    std::vector<DfaGroup> groups;
    // event mode: accept set k of group g -> reports[report_begin[g][k] .. report_begin[g][k+1])
    std::vector<std::vector<uint32_t>> report_begin;
    std::vector<ReportDesc> reports;
    FactorSet factors;       // required factors per pattern (input of the prefilter builder)
    Prefilter prefilter;     // statically chosen windows (used when no input sample is available)
    std::string key;   // cache key: patterns + flags + ids
};

// Return codes follow the reference's enum (hyperscanner.c:25-33): 0 ok, 4 = HYPERSCANNER_DB for any rejection.
// `error` receives a human-readable reason (printed to stderr by the boundary, never returned).
int compile_database(const char* const* patterns, const unsigned* flags, const unsigned* ids, unsigned n,
                     std::shared_ptr<Database>& out, std::string& error);

// Process-wide cache (the reference recompiles per call, hyperscanner.c:296).
std::shared_ptr<Database> cached_database(const char* const* patterns, const unsigned* flags, const unsigned* ids,
                                          unsigned n, int& rc, std::string& error);

}  // namespace gpugrep
