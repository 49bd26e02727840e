// Seeded synthetic syslog-shaped text: bench and parity inputs.  Host only, no CUDA.
//   Mon DD HH:MM:SS hostNNN proc[pid]: LEVEL message...\n      80-250 bytes, mean ~145
//   LEVEL: INFO 90 %, WARN 7 %, ERROR 2 %, DEBUG 1 %
// Rare message templates (ssh failures, OOM kills, segfaults, ...) give the BASELINE pattern sets something to
// find; `plants` lets a caller inject its own indicator strings (config 3's IOC set) at a chosen rate.
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/gpugrep_synth.h"

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull) { next(); next(); }
    uint64_t next() {
        s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
        return s * 0x2545F4914F6CDD1Dull;
    }
    uint32_t below(uint32_t n) { return (uint32_t)((next() >> 32) * (uint64_t)n >> 32); }
};

const char* const kMonths[] = {"Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec"};
const char* const kProcs[] = {"sshd", "kernel", "systemd", "nginx", "postfix/smtpd", "cron", "dockerd", "kubelet", "haproxy", "postgres",
                              "redis-server", "auditd", "NetworkManager", "rsyslogd", "containerd", "etcd"};
const char* const kWords[] = {
    "request", "completed", "started", "worker", "queue", "flush", "buffer", "timeout", "retry", "client", "server", "handler", "thread",
    "pool", "resource", "update", "status", "service", "module", "loaded", "config", "reload", "signal", "received", "process", "exited",
    "checkpoint", "snapshot", "volume", "mounted", "device", "link", "state", "changed", "lease", "renewed", "route", "added", "packet",
    "forwarded", "cache", "evicted", "index", "rebuilt", "shard", "replica", "synced", "leader", "elected", "heartbeat", "latency", "bytes",
    "written", "read", "transaction", "committed", "rollback", "scheduled", "job", "finished", "backup", "rotated", "compressed", "archive",
    "upload", "download", "stream", "closed", "opened", "listener", "bound", "socket", "accepted", "throttled", "quota", "exceeded", "limit",
    "warning", "notice", "metric", "sampled", "trace", "span", "exported", "batch", "processed", "pipeline", "stage", "ok", "done", "pending",
    "waiting", "lock", "acquired", "released", "session", "token", "refreshed", "certificate", "verified", "handshake", "negotiated", "cipher",
    "policy", "applied", "rule", "matched", "allowed", "node", "ready", "pod", "container", "image", "pulled", "layer", "extracted", "health",
    "probe", "passed", "endpoint", "registered", "dns", "resolved", "upstream", "backend", "balanced", "weight", "adjusted"};
const char* const kUsers[] = {"root", "admin", "deploy", "ubuntu", "postgres", "git", "jenkins", "backup", "oracle", "www-data", "svc-build", "alice", "bob"};
constexpr int kNumWords = sizeof(kWords) / sizeof(kWords[0]);

struct Out {
    char* p;
    char* end;
    void put(const char* s) { size_t n = std::strlen(s); std::memcpy(p, s, n); p += n; }
    void putc(char c) { *p++ = c; }
    void num(uint32_t v, int width = 0) {
        char tmp[16]; int n = 0;
        do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (n < width) tmp[n++] = '0';
        while (n) *p++ = tmp[--n];
    }
    void hex(uint64_t v, int digits) {
        for (int i = digits - 1; i >= 0; i--) *p++ = "0123456789abcdef"[(v >> (4 * i)) & 15];
    }
};

void put_ip(Out& o, Rng& r) {
    o.num(10 + r.below(200)); o.putc('.'); o.num(r.below(256)); o.putc('.'); o.num(r.below(256)); o.putc('.'); o.num(1 + r.below(254));
}

void put_words(Out& o, Rng& r, int count) {
    for (int i = 0; i < count; i++) { if (i) o.putc(' '); o.put(kWords[r.below(kNumWords)]); }
}

}  // namespace

extern "C" {

// Fills out[0, size) with complete lines (the last byte written is '\n').  Returns the number of lines.
// plants: optional indicator strings; a line receives one with probability plant_ppm / 1e6.
GPUGREP_SYNTH_API size_t gpugrep_synth_syslog(unsigned long long seed, char* out, size_t size, const char* const* plants, unsigned int nplants,
                                        unsigned int plant_ppm) {
    Rng r(seed);
    Out o{out, out + size};
    size_t lines = 0;
    uint32_t sec = r.below(86400), day = 1 + r.below(28), mon = r.below(12);
    while ((size_t)(o.end - o.p) >= 640) {
        sec += r.below(3);
        if (sec >= 86400) { sec -= 86400; day = day % 28 + 1; }
        o.put(kMonths[mon]); o.putc(' '); o.num(day, 2); o.putc(' ');
        o.num(sec / 3600, 2); o.putc(':'); o.num(sec / 60 % 60, 2); o.putc(':'); o.num(sec % 60, 2);
        o.put(" host"); o.num(r.below(1000), 3); o.putc(' ');
        o.put(kProcs[r.below(16)]); o.putc('['); o.num(100 + r.below(64000)); o.put("]: ");
        uint32_t lv = r.below(100);
        o.put(lv < 90 ? "INFO " : lv < 97 ? "WARN " : lv < 99 ? "ERROR " : "DEBUG ");
        uint32_t t = r.below(10000);
        if (t < 5600) {
            put_words(o, r, 5 + (int)r.below(14)); o.put(" req="); o.hex(r.next(), 8); o.put(" dur="); o.num(r.below(5000)); o.put("ms");
        } else if (t < 7100) {
            o.put("session opened for user "); o.put(kUsers[r.below(13)]); o.put(" by (uid="); o.num(r.below(2000)); o.putc(')');
            o.putc(' '); put_words(o, r, 2 + (int)r.below(6));
        } else if (t < 8100) {
            o.put("GET /api/v1/"); o.put(kWords[r.below(kNumWords)]); o.putc('/'); o.num(r.below(100000)); o.put(" HTTP/1.1 ");
            o.num(r.below(50) ? 200 : 404); o.putc(' '); o.num(r.below(90000)); o.put("B "); o.num(r.below(900)); o.put("ms ua=");
            put_words(o, r, 1 + (int)r.below(3));
        } else if (t < 8700) {
            o.put("connection from "); put_ip(o, r); o.put(" sport="); o.num(1024 + r.below(64000)); o.put(" proto=tcp ");
            put_words(o, r, 2 + (int)r.below(8));
        } else if (t < 9400) {
            o.put("cache "); o.put(kWords[r.below(kNumWords)]); o.put(" hit ratio=0."); o.num(r.below(100), 2); o.put(" keys="); o.num(r.below(1000000));
            o.putc(' '); put_words(o, r, 3 + (int)r.below(10));
        } else if (t < 9500) {
            o.put("txn id="); o.hex(r.next(), 12); o.put(" state="); o.put(kWords[r.below(kNumWords)]); o.putc(' '); put_words(o, r, 3 + (int)r.below(8));
        } else if (t < 9550) {
            o.put("Accepted publickey for "); o.put(kUsers[r.below(13)]); o.put(" from "); put_ip(o, r); o.put(" port "); o.num(1024 + r.below(64000)); o.put(" ssh2");
        } else if (t < 9610) {
            o.put(r.below(4) ? "Failed password for " : "failed password for invalid user "); o.put(kUsers[r.below(13)]); o.put(" from "); put_ip(o, r);
            o.put(" port "); o.num(1024 + r.below(64000)); o.put(" ssh2");
        } else if (t < 9640) {
            o.put("segfault at "); o.hex(r.next(), 12); o.put(" ip "); o.hex(r.next(), 16); o.put(" sp "); o.hex(r.next(), 16); o.put(" error "); o.num(r.below(16));
            o.put(" in lib"); o.put(kWords[r.below(kNumWords)]); o.put(".so");
        } else if (t < 9670) {
            o.put("Out of memory: Kill process "); o.num(r.below(65000)); o.put(" ("); o.put(kWords[r.below(kNumWords)]); o.put(") score "); o.num(r.below(1000));
            o.put(" or sacrifice child");
        } else if (t < 9710) {
            o.put("nf_conntrack: table full, dropping packet");
        } else if (t < 9740) {
            o.put("authentication failure; logname= uid=0 euid=0 tty=ssh ruser= rhost="); put_ip(o, r); o.put(" user="); o.put(kUsers[r.below(13)]);
        } else if (t < 9765) {
            o.put("TLS handshake error from "); put_ip(o, r); o.putc(':'); o.num(1024 + r.below(64000)); o.put(": remote error: tls: bad certificate");
        } else if (t < 9785) {
            o.put("disk quota exceeded on /dev/sd"); o.putc((char)('a' + r.below(6))); o.num(1 + r.below(4)); o.put(" inode="); o.num(r.below(9000000));
        } else if (t < 9800) {
            o.put("possible SYN flooding on port "); o.num(r.below(4) ? 443 : 8080); o.put(". Sending cookies.");
        } else if (t < 9810) {
            o.put("I/O error, dev nvme"); o.num(r.below(4)); o.put("n1, sector "); o.num(r.below(2000000000u)); o.put(" op 0x1:(WRITE)");
        } else if (t < 9820) {
            o.put("panic: runtime error: invalid memory address or nil pointer dereference goroutine "); o.num(r.below(5000));
        } else {
            put_words(o, r, 8 + (int)r.below(12));
        }
        if (nplants && r.below(1000000) < plant_ppm) { o.putc(' '); o.put(plants[r.below(nplants)]); }
        if (r.below(3) == 0) { o.put(" trace="); o.hex(r.next(), 16); }
        o.putc('\n');
        lines++;
    }
    // final filler line so that the buffer is exactly full and newline-terminated
    size_t rest = (size_t)(o.end - o.p);
    if (rest > 0) {
        for (size_t i = 0; i + 1 < rest; i++) o.putc(i == 0 ? '#' : (i % 7 == 6 ? ' ' : 'x'));
        o.putc('\n');
        lines++;
    }
    return lines;
}

}  // extern "C"
