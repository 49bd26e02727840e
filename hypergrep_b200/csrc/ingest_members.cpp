// Host ingest: one compressed file decoded by several threads.
//
// The reference reads a .gz / .zst file through ONE gzgets() stream (hyperscanner.c:189-199): 0.2 GB/s (gzip) and
// 0.6 GB/s (zstd) of text per core, two orders of magnitude below what one GPU scans.  Files written by bgzip, pzstd,
// `zstd -B`, log rotation with `cat a.gz b.gz`, ... hold many independent gzip members / zstd frames.  This source
// decodes them ahead of the reader on helper threads and hands the text over in file order.
//
// Member boundaries are not indexed in either format, so the helpers SPECULATE: about every MiB of the compressed file
// the next byte sequence that looks like a member header becomes a start, and a helper decodes members from there
// until it reaches another start (or has 16 MiB of text, then it registers the boundary it stopped at as a start).
// The reader walks the real chain of members from offset 0: at every boundary it either finds a finished (or running)
// speculative decode that begins exactly there and takes its text, or decodes the member itself, streaming, as the
// sequential source does.  A start inside a member (a false header) is never reached at a boundary and its output is
// dropped, so the bytes delivered are exactly those of the sequential decode - including where it stops: trailing
// garbage, a corrupt or truncated member (text up to the error is kept), a skippable zstd frame (ingest_codecs.hpp).
// A file that is one large member is decoded by the reader alone, streaming; a speculative decode that grows past
// 128 MiB is dropped and left to the reader, so memory stays bounded (window x (16 MiB + one member)).
#include "ingest.hpp"
#include "ingest_codecs.hpp"

#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

namespace gpugrep {
namespace {

// helper threads of all parallel decodes of the process (multiscanner runs one scan per host thread)
std::atomic<int> g_decode_helpers{0};

// (GPUGREP_DECODE_MIN_BYTES / _SPACING / _CHAIN override the first three: the tests run small files through many starts)
constexpr size_t kMinParallelBytes = (size_t)4 << 20;   // smaller files: one thread
constexpr size_t kStartSpacing = (size_t)1 << 20;       // compressed bytes between speculative starts
constexpr size_t kChainTarget = (size_t)16 << 20;       // text after which a helper stops at the next member boundary
constexpr size_t kMaxTaskBytes = (size_t)128 << 20;     // a speculative decode larger than this is left to the reader
constexpr size_t kScanPerCall = (size_t)64 << 20;       // compressed bytes searched for starts per scheduling call
constexpr int kMaxHelpers = 16;

size_t env_bytes(const char* name, size_t fallback) {
    const char* e = std::getenv(name);
    return e && *e ? (size_t)std::strtoull(e, nullptr, 10) : fallback;
}

// Could a member begin here?  (Only a hint: what counts is whether the reader arrives here at a member boundary.)
bool start_candidate(Packing kind, const uint8_t* p, size_t remaining) {
    if (kind == Packing::Gzip) return remaining >= 18 && p[0] == 0x1f && p[1] == 0x8b && p[2] == 8 && (p[3] & 0xe0) == 0;
    return remaining >= 9 && member_continues(kind, p) && (p[4] & 0x08) == 0;   // frame header descriptor: reserved bit
}

class ParallelMemberSource : public ByteSource {
    struct Task {
        enum State { Queued, Running, Done, Dropped };
        explicit Task(size_t s) : start(s) {}
        size_t start, end = 0;
        std::vector<uint8_t> text;
        State state = Queued;
        bool last = false;   // the data ends with this text (garbage, a corrupt or a truncated member follows)
        std::atomic<bool> cancel{false};
    };

public:
    ParallelMemberSource(Packing kind, const uint8_t* map, size_t size, int helpers)
        : kind_(kind), map_(map), size_(size), helpers_(helpers), spacing_(std::max<size_t>(1, env_bytes("GPUGREP_DECODE_SPACING", kStartSpacing))),
          chain_target_(env_bytes("GPUGREP_DECODE_CHAIN", kChainTarget)), codec_(make_member_codec(kind)) {
        max_tasks_ = (size_t)helpers * 2 + 2;
        for (int i = 0; i < helpers; i++) pool_.emplace_back([this] { helper(); });
    }
    ~ParallelMemberSource() override {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            for (auto& kv : tasks_) kv.second->cancel = true;
        }
        cv_work_.notify_all();
        for (auto& th : pool_) th.join();
        ::munmap(const_cast<uint8_t*>(map_), size_);
        g_decode_helpers.fetch_sub(helpers_);
        if (std::getenv("GPUGREP_DECODE_TRACE"))
            std::fprintf(stderr, "[gpugrep] %s decode: %d helpers, %zu texts taken from helpers (%zu bytes), %zu members decoded by the reader, %zu starts dropped\n",
                         kind(), helpers_, served_tasks_, served_bytes_, own_members_, dropped_);
    }

    size_t read(uint8_t* dst, size_t cap) override {
        if (!codec_ || !codec_->ok()) return 0;
        size_t produced = 0;
        while (produced < cap && !done_) {
            if (serving_) {
                const size_t n = std::min(cap - produced, serving_->text.size() - served_);
                std::memcpy(dst + produced, serving_->text.data() + served_, n);
                served_ += n;
                produced += n;
                if (served_ == serving_->text.size()) {
                    pos_ = serving_->end;
                    if (serving_->last) done_ = true;
                    serving_.reset();
                    between_ = true;
                }
                continue;
            }
            if (between_) {
                if (size_ - pos_ < member_header_bytes(kind_) || !member_continues(kind_, map_ + pos_)) { done_ = true; break; }
                if (take_task_at_boundary()) continue;
                codec_->reset();
                between_ = false;
                own_members_++;
            }
            size_t used = 0, made = 0;
            const MemberCodec::Step st = codec_->step(map_ + pos_, size_ - pos_, used, dst + produced, cap - produced, made);
            pos_ += used;
            produced += made;
            if (st == MemberCodec::MemberEnd) { between_ = true; continue; }
            if (st == MemberCodec::Failed) { done_ = true; break; }
            if (pos_ >= size_ && used == 0 && made == 0) { done_ = true; break; }   // truncated member
        }
        return produced;
    }
    const char* kind() const override { return kind_ == Packing::Gzip ? "gzip" : "zstd"; }

private:
    // The reader stands at the member boundary pos_.  True: a speculative decode that starts here is (or will be) served.
    bool take_task_at_boundary() {
        std::unique_lock<std::mutex> lk(mu_);
        // starts the reader has passed were inside a member, or are consumed
        while (!tasks_.empty() && tasks_.begin()->first < pos_) {
            tasks_.begin()->second->cancel = true;
            tasks_.erase(tasks_.begin());
            dropped_++;
        }
        schedule_locked();
        auto it = tasks_.find(pos_);
        if (it == tasks_.end()) return false;
        std::shared_ptr<Task> t = it->second;
        if (t->state == Task::Queued) {   // nobody has started it: the reader is a decoder too
            tasks_.erase(it);
            return false;
        }
        cv_done_.wait(lk, [&] { return t->state != Task::Running; });
        tasks_.erase(pos_);
        if (t->state != Task::Done) return false;
        serving_ = t;
        served_ = 0;
        served_tasks_++;
        served_bytes_ += t->text.size();
        if (t->text.empty()) {   // (a member without text)
            pos_ = t->end;
            if (t->last) done_ = true;
            serving_.reset();
        }
        return true;
    }

    void add_task_locked(size_t start) {
        if (tasks_.count(start)) return;
        tasks_.emplace(start, std::make_shared<Task>(start));
        cv_work_.notify_one();
    }

    void schedule_locked() {
        size_t searched = 0;
        while (tasks_.size() < max_tasks_ && searched < kScanPerCall) {
            size_t from = std::max(scan_from_, pos_ + 1);
            if (from >= size_) break;
            const size_t limit = std::min(size_, from + (kScanPerCall - searched));
            size_t found = SIZE_MAX;
            const uint8_t first = kind_ == Packing::Gzip ? 0x1f : 0x28;
            for (size_t at = from; at < limit;) {
                const void* hit = std::memchr(map_ + at, first, limit - at);
                if (!hit) break;
                at = (size_t)((const uint8_t*)hit - map_);
                if (start_candidate(kind_, map_ + at, size_ - at)) { found = at; break; }
                at++;
            }
            if (found == SIZE_MAX) {
                searched += limit - from;
                scan_from_ = limit;
                continue;
            }
            searched += found - from;
            add_task_locked(found);
            scan_from_ = std::max(found + 1, (found / spacing_ + 1) * spacing_);
        }
    }

    void helper() {
        std::unique_ptr<MemberCodec> codec = make_member_codec(kind_);
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            std::shared_ptr<Task> t;
            cv_work_.wait(lk, [&] {
                if (stop_) return true;
                for (auto& kv : tasks_)
                    if (kv.second->state == Task::Queued) { t = kv.second; return true; }
                return false;
            });
            if (stop_) return;
            t->state = Task::Running;
            lk.unlock();
            const Task::State result = codec && codec->ok() ? decode_chain(*t, *codec) : Task::Dropped;
            lk.lock();
            t->state = result;
            if (result != Task::Done) std::vector<uint8_t>().swap(t->text);
            cv_done_.notify_all();
        }
    }

    // Members from t.start on, until a boundary that is another start, 16 MiB of text, or the end of the data.
    Task::State decode_chain(Task& t, MemberCodec& codec) {
        size_t cur = t.start, filled = 0;
        codec.reset();
        for (;;) {
            if (t.cancel.load(std::memory_order_relaxed)) return Task::Dropped;
            if (filled > kMaxTaskBytes) return Task::Dropped;
            if (t.text.size() - filled < ((size_t)256 << 10)) t.text.resize(std::max(t.text.size() * 2, filled + ((size_t)4 << 20)));
            size_t used = 0, made = 0;
            // (bounded input per call, so that a cancelled decode of a long member stops soon)
            const size_t in_len = std::min(size_ - cur, (size_t)4 << 20);
            const MemberCodec::Step st = codec.step(map_ + cur, in_len, used, t.text.data() + filled, t.text.size() - filled, made);
            cur += used;
            filled += made;
            if (st == MemberCodec::Failed) { t.last = true; break; }
            if (st == MemberCodec::MemberEnd) {
                if (size_ - cur < member_header_bytes(kind_) || !member_continues(kind_, map_ + cur)) { t.last = true; break; }
                bool hand_over;
                {
                    std::lock_guard<std::mutex> lk(mu_);
                    hand_over = tasks_.count(cur) != 0;
                    if (!hand_over && filled >= chain_target_ && !stop_) {
                        add_task_locked(cur);
                        hand_over = true;
                    }
                }
                if (hand_over) break;
                codec.reset();
                continue;
            }
            if (cur >= size_ && used == 0 && made == 0) { t.last = true; break; }   // truncated member
        }
        t.end = cur;
        t.text.resize(filled);
        return Task::Done;
    }

    const Packing kind_;
    const uint8_t* const map_;
    const size_t size_;
    const int helpers_;
    const size_t spacing_, chain_target_;
    std::unique_ptr<MemberCodec> codec_;   // the reader's own decoder
    std::vector<std::thread> pool_;

    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    std::map<size_t, std::shared_ptr<Task>> tasks_;   // by start offset
    size_t max_tasks_ = 0;
    size_t scan_from_ = 1;   // compressed offset from which the next start is searched
    bool stop_ = false;

    // reader state
    size_t pos_ = 0;          // compressed offset (a member boundary when between_)
    bool between_ = true, done_ = false;
    std::shared_ptr<Task> serving_;
    size_t served_ = 0;
    size_t served_tasks_ = 0, served_bytes_ = 0, own_members_ = 0, dropped_ = 0;   // GPUGREP_DECODE_TRACE
};

}  // namespace

std::unique_ptr<ByteSource> open_parallel_members(Packing kind, int fd) {
    struct stat sb;
    if (::fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode) || (size_t)sb.st_size < env_bytes("GPUGREP_DECODE_MIN_BYTES", kMinParallelBytes)) return nullptr;
    const int hw = (int)std::thread::hardware_concurrency();
    int take = 0;
    if (const char* e = std::getenv("GPUGREP_DECODE_THREADS")) {
        take = std::max(0, std::min(64, std::atoi(e) - 1));   // the count includes the reader
        if (take <= 0) return nullptr;
        g_decode_helpers.fetch_add(take);
    } else {
        // up to sixteen helpers out of what is left of the process-wide budget (cores - 2).  The reader mostly copies
        // finished text, so the decode rate is about `helpers` times that of one thread.
        const int budget = std::max(0, hw - 2), want = std::min(kMaxHelpers, budget);
        int busy = g_decode_helpers.load();
        do {
            take = std::min(want, std::max(0, budget - busy));
            if (take <= 0) return nullptr;
        } while (!g_decode_helpers.compare_exchange_weak(busy, busy + take));
    }
    void* map = ::mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map == MAP_FAILED) {
        g_decode_helpers.fetch_sub(take);
        return nullptr;
    }
    ::madvise(map, (size_t)sb.st_size, MADV_SEQUENTIAL);
    return std::make_unique<ParallelMemberSource>(kind, (const uint8_t*)map, (size_t)sb.st_size, take);
}

}  // namespace gpugrep
