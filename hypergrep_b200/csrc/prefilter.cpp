// Required-factor analysis and gram-bitmap construction.  See prefilter.hpp.
#include "prefilter.hpp"

#include <algorithm>
#include <cmath>
#include <set>

namespace gpugrep {
namespace {

constexpr size_t kMaxSet = 64;     // alternatives tracked per node
constexpr size_t kMaxLen = 24;     // positions tracked per class-string
constexpr double kMaxGramsPerWindow = 4096.0;
constexpr size_t kMaxGramsTotal = 131072;

using Alts = std::vector<ClassString>;

struct Lits {
    bool exact_ok = false;
    Alts exact;    // the node matches exactly one of these (valid iff exact_ok)
    Alts prefix;   // every match starts with one of these ({""} = nothing known)
    Alts suffix;   // every match ends with one of these
    Alts best;     // required factor: every match contains one of these (empty = none found)
};

const Alts kEpsilon = {ClassString{}};

size_t max_len(const Alts& a) { size_t m = 0; for (auto& s : a) m = std::max(m, s.size()); return m; }
size_t min_len(const Alts& a) { size_t m = a.empty() ? 0 : SIZE_MAX; for (auto& s : a) m = std::min(m, s.size()); return m; }

void dedupe(Alts& a) {
    std::sort(a.begin(), a.end());
    a.erase(std::unique(a.begin(), a.end()), a.end());
}

bool fits(const Alts& a, const Alts& b) { return a.size() * b.size() <= kMaxSet && max_len(a) + max_len(b) <= kMaxLen; }

Alts cross(const Alts& a, const Alts& b) {
    Alts r;
    for (auto& x : a) for (auto& y : b) { ClassString s = x; s.insert(s.end(), y.begin(), y.end()); r.push_back(std::move(s)); }
    dedupe(r);
    return r;
}
// products that may be truncated: prefixes keep their head, suffixes keep their tail
Alts cross_head(const Alts& a, const Alts& b) {
    if (a.size() * b.size() > kMaxSet) return a;
    Alts r = cross(a, b);
    for (auto& s : r) if (s.size() > kMaxLen) s.resize(kMaxLen);
    dedupe(r);
    return r;
}
Alts cross_tail(const Alts& a, const Alts& b) {
    if (a.size() * b.size() > kMaxSet) return b;
    Alts r = cross(a, b);
    for (auto& s : r) if (s.size() > kMaxLen) s.erase(s.begin(), s.end() - kMaxLen);
    dedupe(r);
    return r;
}

// Probability that a random 4-gram falls in the most selective window of `s` (uniform byte model).
double window_prob(const ClassString& s, size_t w) {
    if (s.empty()) return 1.0;
    w = std::min(w, s.size());
    double best = 1.0;
    for (size_t t = 0; t + w <= s.size(); t++) {
        double p = 1.0;
        for (size_t i = 0; i < w; i++) p *= s[t + i].count() / 256.0;
        best = std::min(best, p);
    }
    return best;
}

// Lower is better; infinity = unusable.
double cost(const Alts& a) {
    if (a.empty()) return INFINITY;
    size_t ml = min_len(a);
    if (ml < 4) return INFINITY;   // the filter hashes 4-byte grams
    double tier = ml >= 7 ? 1.0 : (ml >= 5 ? 2.0 : 4.0);   // stride 4 / 2 / 1 in the streaming kernel
    double p = 0;
    for (auto& s : a) p += window_prob(s, 4);
    return p * tier;
}

void consider(Alts& best, const Alts& cand) {
    if (cost(cand) < cost(best)) best = cand;
}

Lits analyse(const Node& n) {
    Lits r;
    switch (n.kind) {
        case NodeKind::Empty:
        case NodeKind::Assert:
            r.exact_ok = true; r.exact = kEpsilon; r.prefix = kEpsilon; r.suffix = kEpsilon;
            return r;
        case NodeKind::Set: {
            r.exact_ok = true;
            r.exact = {ClassString{n.set}};
            r.prefix = r.suffix = r.exact;
            return r;
        }
        case NodeKind::Concat: {
            Alts run = kEpsilon;
            bool whole = true;       // everything so far is inside `run`
            r.prefix = kEpsilon;
            for (auto& kid : n.kids) {
                Lits k = analyse(*kid);
                if (k.exact_ok && fits(run, k.exact)) { run = cross(run, k.exact); continue; }
                Alts closed = cross_head(run, k.exact_ok ? k.exact : k.prefix);
                consider(r.best, closed);
                consider(r.best, k.best);
                if (k.exact_ok) consider(r.best, k.exact);
                if (whole) r.prefix = closed;
                whole = false;
                run = k.exact_ok ? k.exact : k.suffix;
                if (k.exact_ok) for (auto& s : run) if (s.size() > kMaxLen) s.erase(s.begin(), s.end() - kMaxLen);
            }
            if (whole) {
                r.exact_ok = true; r.exact = run; r.prefix = run; r.suffix = run;
                for (auto& s : r.prefix) if (s.size() > kMaxLen) s.resize(kMaxLen);
            } else {
                consider(r.best, run);
                r.suffix = run;
            }
            return r;
        }
        case NodeKind::Alt: {
            r.exact_ok = true;
            bool all_factor = true;
            Alts pre, suf, fac;
            bool pre_ok = true, suf_ok = true;
            for (auto& kid : n.kids) {
                Lits k = analyse(*kid);
                if (k.exact_ok && r.exact_ok && r.exact.size() + k.exact.size() <= kMaxSet) r.exact.insert(r.exact.end(), k.exact.begin(), k.exact.end());
                else r.exact_ok = false;
                const Alts& f = (k.exact_ok && cost(k.exact) <= cost(k.best)) ? k.exact : k.best;
                if (std::isinf(cost(f))) all_factor = false; else fac.insert(fac.end(), f.begin(), f.end());
                pre.insert(pre.end(), k.prefix.begin(), k.prefix.end());
                suf.insert(suf.end(), k.suffix.begin(), k.suffix.end());
            }
            dedupe(pre); dedupe(suf); dedupe(fac);
            if (pre.size() > kMaxSet) pre_ok = false;
            if (suf.size() > kMaxSet) suf_ok = false;
            r.prefix = pre_ok ? pre : kEpsilon;
            r.suffix = suf_ok ? suf : kEpsilon;
            if (r.exact_ok) dedupe(r.exact); else r.exact.clear();
            if (all_factor && fac.size() <= 4 * kMaxSet) r.best = fac;
            return r;
        }
        case NodeKind::Repeat: {
            Lits k = analyse(*n.kids[0]);
            if (n.min == 0) {
                r.prefix = r.suffix = kEpsilon;
                if (n.max == 1 && k.exact_ok && k.exact.size() + 1 <= kMaxSet) {
                    r.exact_ok = true; r.exact = k.exact; r.exact.push_back(ClassString{}); dedupe(r.exact);
                }
                return r;
            }
            // min >= 1: the child occurs at least `min` times in a row
            if (k.exact_ok) {
                Alts pow = k.exact;
                int reps = 1;
                while (reps < n.min && fits(pow, k.exact)) { pow = cross(pow, k.exact); reps++; }
                if (reps == n.min && n.max == n.min) { r.exact_ok = true; r.exact = pow; r.prefix = r.suffix = pow; return r; }
                r.prefix = r.suffix = pow;
                r.best = pow;
                if (std::isinf(cost(r.best))) r.best.clear();
                return r;
            }
            r.prefix = k.prefix; r.suffix = k.suffix; r.best = k.best;
            return r;
        }
    }
    return r;
}

struct Window { const ClassString* s; size_t start; };

double grams_in_window(const ClassString& s, size_t t, int stride, bool fold) {
    double total = 0;
    for (int j = 0; j < stride; j++) {
        double p = 1;
        for (int i = 0; i < 4; i++) {
            const ByteSet& b = s[t + j + i];
            int c = 0;
            if (fold) { ByteSet f; for (unsigned v = 0; v < 256; v++) if (b.test(v)) f.set(v | 0x20); c = f.count(); }
            else c = b.count();
            p *= c;
        }
        total += p;
    }
    return total;
}

}  // namespace

bool extract_factor(const Node& ast, std::vector<ClassString>& alternatives) {
    Lits l = analyse(ast);
    Alts best = l.best;
    if (l.exact_ok) consider(best, l.exact);
    consider(best, l.prefix);
    consider(best, l.suffix);
    if (std::isinf(cost(best))) return false;
    alternatives = best;
    return true;
}

void build_prefilter(const std::vector<const Node*>& asts, const std::vector<unsigned>& flags, Prefilter& out) {
    out = Prefilter();
    std::vector<Alts> factors(asts.size());
    size_t ml = SIZE_MAX;
    for (size_t i = 0; i < asts.size(); i++) {
        if (!extract_factor(*asts[i], factors[i])) {
            out.note = "pattern " + std::to_string(i) + " has no required factor of >= 4 bytes";
            return;
        }
        ml = std::min(ml, min_len(factors[i]));
    }
    (void)flags;
    out.min_factor_len = (int)ml;
    int first_stride = ml >= 7 ? 4 : (ml >= 5 ? 2 : 1);
    for (int stride = first_stride; stride >= 1; stride /= 2) {
        size_t w = 3 + (size_t)stride;
        // pick, per alternative, the window with the fewest grams; decide on folding from the totals
        for (int pass = 0; pass < 2; pass++) {
            bool fold = pass == 1;
            std::vector<Window> wins;
            double total = 0;
            bool ok = true;
            for (auto& alts : factors) {
                for (auto& s : alts) {
                    double best = INFINITY; size_t bt = 0;
                    for (size_t t = 0; t + w <= s.size(); t++) {
                        double g = grams_in_window(s, t, stride, fold);
                        if (g < best) { best = g; bt = t; }
                    }
                    if (best > kMaxGramsPerWindow * stride) { ok = false; break; }
                    total += best;
                    wins.push_back(Window{&s, bt});
                }
                if (!ok) break;
            }
            if (!ok || total > (double)kMaxGramsTotal) continue;
            if (!fold) {
                // prefer exact-case grams unless folding shrinks the table a lot (caseless sets)
                double folded_total = 0;
                for (auto& wn : wins) folded_total += grams_in_window(*wn.s, wn.start, stride, true);
                if (total > 3.0 * folded_total && total > 2048) continue;
            }
            // materialise the grams
            std::set<uint32_t> grams;
            for (auto& wn : wins) {
                for (int j = 0; j < stride; j++) {
                    std::vector<uint32_t> cur = {0};
                    for (int i = 0; i < 4; i++) {
                        const ByteSet& b = (*wn.s)[wn.start + j + i];
                        ByteSet eff;
                        if (fold) { for (unsigned v = 0; v < 256; v++) if (b.test(v)) eff.set(v | 0x20); } else eff = b;
                        std::vector<uint32_t> next;
                        next.reserve(cur.size() * eff.count());
                        for (uint32_t g : cur) for (unsigned v = 0; v < 256; v++) if (eff.test(v)) next.push_back(g | (v << (8 * i)));
                        cur.swap(next);
                    }
                    grams.insert(cur.begin(), cur.end());
                }
            }
            size_t need = grams.size() * 64;
            int lb = 13;
            while (lb < 20 && (1ull << lb) < need) lb++;
            out.enabled = true;
            out.stride = stride;
            out.fold_case = fold;
            out.log2_bits = lb;
            out.bitmap.assign((1u << lb) / 32, 0);
            for (uint32_t g : grams) {
                uint32_t h = prefilter_hash(g, out.hash_mul, lb);
                out.bitmap[h >> 5] |= 1u << (h & 31);
            }
            out.num_grams = grams.size();
            out.note = "stride " + std::to_string(stride) + (fold ? ", folded" : "") + ", " + std::to_string(grams.size()) + " grams, " +
                       std::to_string(1u << lb) + " bits";
            return;
        }
    }
    out.note = "gram expansion too large";
}

}  // namespace gpugrep
