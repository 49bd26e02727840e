// TEST-ONLY stand-in for hypergrep_b200/csrc/engine.cu.
//
// Implements the host-side engine interface (engine.hpp) with plain C++ loops so that the HOST logic of the
// boundary (capi.cpp: segment cutting, double buffering, batching, max_match_count, return codes, ingest) can be
// exercised by `pytest -m "not gpu"` on a machine without a GPU.  It is built into tests/_build/ only, is never
// part of libgpugrep.so, and nothing in hypergrep_b200/ references it.  The CUDA kernels themselves are checked
// on the GPU by the `-m gpu` tests.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "engine.hpp"
#include "nfa_sim.hpp"

namespace gpugrep {

struct DeviceDb {
    std::shared_ptr<Database> db;
};
struct DevicePrefilter {
    int unused = 0;
};

class ScanSlot {
public:
    std::vector<uint8_t> stage;
    const DeviceDb* ddb = nullptr;
    const uint8_t* data = nullptr;
    size_t n = 0;
    int buffer_size = 0;
    std::vector<LineRec> recs;
    std::vector<EventRec> events;
};

int engine_select_device(int, std::string&) { return 0; }
int engine_current_device() { return 0; }
int engine_device_count() {
    // $GPUGREP_MOCK_DEVICES pretends several devices so that the shard / merge logic of the boundary runs on CPU
    const char* e = std::getenv("GPUGREP_MOCK_DEVICES");
    return e && *e ? std::max(1, std::atoi(e)) : 1;
}
void slot_set_want_records(ScanSlot*, bool) {}

std::shared_ptr<DeviceDb> engine_upload(const std::shared_ptr<Database>& db, std::string&) {
    auto d = std::make_shared<DeviceDb>();
    d->db = db;
    return d;
}

std::shared_ptr<DevicePrefilter> engine_upload_prefilter(const Prefilter& pf, std::string&) {
    return pf.enabled ? std::make_shared<DevicePrefilter>() : nullptr;
}

double prefilter_expected_hits(const DevicePrefilter*) { return -1.0; }

ScanSlot* engine_acquire_slot(std::string&, bool) { return new ScanSlot(); }
void engine_release_slot(ScanSlot* s) { delete s; }

uint8_t* slot_host_buffer(ScanSlot* s, size_t capacity, std::string&) {
    if (s->stage.size() < capacity + 64) s->stage.resize(capacity + 64);
    return s->stage.data();
}

int slot_submit(ScanSlot* s, const DeviceDb& ddb, const DevicePrefilter*, const uint8_t* host_data, const uint8_t* dev_data, size_t n,
                int buffer_size, void*, std::string& error) {
    if (!host_data) { (void)dev_data; error = "mock engine: device-resident input is not supported"; return 7; }
    s->ddb = &ddb; s->data = host_data; s->n = n; s->buffer_size = buffer_size;
    return 0;
}

static void walk(const Database& db, const uint8_t* p, size_t len, uint32_t line, uint32_t start, bool simple, bool& hit, std::vector<EventRec>& ev) {
    size_t a = 0;
    while (a < len && p[a] == 0) a++;
    uint32_t base = 0;
    for (size_t g = 0; g < db.groups.size(); g++) {
        const Dfa& d = db.groups[g].dfa;
        uint32_t s = 0;
        bool dead = false;
        size_t q = a;
        while (q < len) {
            uint8_t b = p[q];
            if (b == 0) break;
            s = d.trans[(size_t)s * d.stride + d.byte_class[b]];
            if ((int)s >= d.first_accept) {
                if (simple) { hit = true; return; }
                ev.push_back(EventRec{line, start, (uint32_t)len, (uint32_t)(q - a), base + d.accept_of[s]});
            }
            if ((int)s == d.dead) { dead = true; break; }
            q++;
            if (b == '\n') break;
        }
        if (!dead) {
            s = d.trans[(size_t)s * d.stride + d.num_classes];
            if ((int)s >= d.first_accept) {
                if (simple) { hit = true; return; }
                ev.push_back(EventRec{line, start, (uint32_t)len, (uint32_t)(q - a), base + d.accept_of[s]});
            }
        }
        base += (uint32_t)d.accept_sets.size();
    }
    if (!db.nfas.empty()) {
        size_t e = a;
        while (e < len) { uint8_t b = p[e]; if (b == 0) break; e++; if (b == '\n') break; }
        for (size_t k = 0; k < db.nfas.size(); k++) {
            const NfaTables& t = db.nfas[k].tables;
            NfaView v;
            v.positions = t.positions; v.words = t.words; v.reach = t.reach.data(); v.follow = t.follow.data();
            v.follow_match = t.follow_match.data(); v.restart = t.restart.data(); v.report = base + (uint32_t)k;
            bool stop = nfa_scan_block(v, p + a, e - a, [&](size_t end) {
                if (simple) { hit = true; return true; }
                ev.push_back(EventRec{line, start, (uint32_t)len, (uint32_t)end, v.report});
                return false;
            });
            if (stop) return;
        }
    }
}

int slot_collect(ScanSlot* s, SegmentResult& out, std::string&, size_t) {
    out = SegmentResult();
    s->recs.clear(); s->events.clear();
    const Database& db = *s->ddb->db;
    size_t limit = (size_t)std::max(1, s->buffer_size - 1);
    size_t pos = 0;
    uint32_t line = 0;
    while (pos < s->n) {
        size_t avail = std::min(limit, s->n - pos);
        const void* nl = std::memchr(s->data + pos, '\n', avail);
        size_t len = nl ? (size_t)((const uint8_t*)nl - (s->data + pos)) + 1 : avail;
        bool hit = false;
        walk(db, s->data + pos, len, line, (uint32_t)pos, db.simple, hit, s->events);
        if (hit) s->recs.push_back(LineRec{line, (uint32_t)pos, (uint32_t)len});
        pos += len;
        line++;
    }
    out.num_lines = line;
    out.lines = s->recs.data(); out.num_line_recs = s->recs.size();
    out.events = s->events.data(); out.num_events = s->events.size();
    out.stats.path = 2;
    return 0;
}

int slot_probe_input(ScanSlot*, const uint8_t*, size_t, size_t, std::vector<size_t>& cuts, uint8_t*, size_t, void*, std::string& error) {
    cuts.clear();
    error = "mock engine: device-resident input is not supported";
    return 7;
}

int slot_gather_lines(ScanSlot*, const uint32_t*, const uint32_t*, size_t, uint8_t*, std::string& error) {
    error = "mock engine: gather is not supported";
    return 7;
}

}  // namespace gpugrep
