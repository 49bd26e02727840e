// Host ingest implementations.  See ingest.hpp.
#include "ingest.hpp"
#include "ingest_codecs.hpp"

#include <dlfcn.h>
#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <sys/stat.h>

#include <atomic>
#include <cerrno>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace gpugrep {
namespace {

std::mutex g_zstd_mu;
std::string g_zstd_path = "libzstd.so.1";

// Helper threads of all plain-file reads of the process (multiscanner runs one scan per host thread): together they
// never exceed the cores of the machine, so fifteen concurrent scans read with one thread each instead of 15 x 14.
std::atomic<int> g_read_helpers{0};

// raw file with a small look-ahead buffer
class RawFile {
public:
    explicit RawFile(int fd) : fd_(fd) {}
    ~RawFile() { if (fd_ >= 0) ::close(fd_); }
    size_t read(uint8_t* dst, size_t cap) {
        size_t got = 0;
        while (got < cap) {
            ssize_t r = ::read(fd_, dst + got, cap - got);
            if (r < 0) { if (errno == EINTR) continue; break; }   // EISDIR etc: end of data (zlib's gzgets returns NULL)
            if (r == 0) break;
            got += (size_t)r;
        }
        return got;
    }
private:
    int fd_;
};

// Plain files.  Large reads of regular files are split over a few threads (pread into disjoint slices of the pinned
// destination): one thread copying out of the page cache tops out far below what the PCIe link can take.
class PlainSource : public ByteSource {
public:
    PlainSource(std::unique_ptr<RawFile> f, const uint8_t* head, size_t head_len, int fd, bool regular)
        : f_(std::move(f)), head_(head, head + head_len), fd_(fd), regular_(regular), offset_(head_len) {}
    // byte range [begin, end) of a regular file
    PlainSource(std::unique_ptr<RawFile> f, int fd, size_t begin, size_t end)
        : f_(std::move(f)), fd_(fd), regular_(true), offset_(begin), limit_(end) {}
    size_t read(uint8_t* dst, size_t cap) override {
        if (limit_ != SIZE_MAX) cap = offset_ < limit_ ? std::min(cap, limit_ - offset_) : 0;
        if (cap == 0) return 0;
        size_t got = 0;
        if (head_pos_ < head_.size()) {
            got = std::min(cap, head_.size() - head_pos_);
            std::memcpy(dst, head_.data() + head_pos_, got);
            head_pos_ += got;
        }
        if (got == cap) return got;
        if (!regular_) return got + f_->read(dst + got, cap - got);
        size_t want = cap - got;
        // threads for this read: what is left of the process-wide budget (cores - 2, at most 16)
        const int hw = (int)std::thread::hardware_concurrency();
        const int budget = std::max(1, std::min(16, hw > 2 ? hw - 2 : 1));
        int taken = 0;
        if (want >= ((size_t)8 << 20)) {
            int busy = g_read_helpers.load();
            while (busy < budget && !g_read_helpers.compare_exchange_weak(busy, budget)) {}
            taken = busy < budget ? budget - busy : 0;
        }
        struct GiveBack { int n; ~GiveBack() { if (n) g_read_helpers.fetch_sub(n); } } give_back{taken};
        const size_t nthreads = (size_t)std::max(1, taken);
        if (nthreads <= 1) {
            size_t r = pread_all(dst + got, want, offset_);
            offset_ += r;
            return got + r;
        }
        // The slices go to helper threads that live as long as this source (the calling thread takes slice 0): starting
        // fourteen threads for every 32 MiB read cost about as much as reading the last slice (hyperscan(path) on a tmpfs
        // file: 43 GB/s, 48 with this; the PCIe link takes 55).  Fresh threads started as a tree - every thread starts the
        // upper half of its range - were measured too: 30 GB/s.
        size_t slice = ((want / nthreads) + 4095) & ~(size_t)4095;
        const size_t nslices = (want + slice - 1) / slice;
        std::vector<size_t> done(nslices, 0);
        {
            std::unique_lock<std::mutex> lk(mu_);
            while (pool_.size() + 1 < nslices) {
                const size_t index = pool_.size();
                pool_.emplace_back([this, index] { helper(index); });
            }
            job_dst_ = dst + got;
            job_off_ = offset_;
            job_want_ = want;
            job_slice_ = slice;
            job_done_ = done.data();
            job_slices_ = nslices;
            pending_ = nslices - 1;
            generation_++;
        }
        cv_go_.notify_all();
        done[0] = pread_all(dst + got, std::min(slice, want), offset_);
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_done_.wait(lk, [this] { return pending_ == 0; });
        }
        size_t total = 0;
        for (size_t t = 0; t < nslices; t++) {
            total += done[t];
            if (done[t] < std::min(slice, want - t * slice)) break;   // end of file inside this slice
        }
        offset_ += total;
        return got + total;
    }
    const char* kind() const override { return "plain"; }
    ~PlainSource() override {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_go_.notify_all();
        for (auto& th : pool_) th.join();
    }
private:
    // helper `index` reads slice index + 1 of every job that has that many slices
    void helper(size_t index) {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> lk(mu_);
        for (;;) {
            cv_go_.wait(lk, [&] { return stop_ || generation_ != seen; });
            if (stop_) return;
            seen = generation_;
            const size_t k = index + 1;
            if (k >= job_slices_) continue;
            uint8_t* dst = job_dst_ + k * job_slice_;
            const size_t len = std::min(job_slice_, job_want_ - k * job_slice_), off = job_off_ + k * job_slice_;
            size_t* out = job_done_ + k;
            lk.unlock();
            const size_t got = pread_all(dst, len, off);
            lk.lock();
            *out = got;
            if (--pending_ == 0) cv_done_.notify_one();
        }
    }
    size_t pread_all(uint8_t* dst, size_t len, size_t off) const {
        size_t got = 0;
        while (got < len) {
            ssize_t r = ::pread(fd_, dst + got, len - got, (off_t)(off + got));
            if (r < 0) { if (errno == EINTR) continue; break; }
            if (r == 0) break;
            got += (size_t)r;
        }
        return got;
    }
    std::unique_ptr<RawFile> f_;
    std::vector<uint8_t> head_;
    size_t head_pos_ = 0;
    int fd_;
    bool regular_;
    size_t offset_;
    size_t limit_ = SIZE_MAX;
    // helper threads of the split reads
    std::vector<std::thread> pool_;
    std::mutex mu_;
    std::condition_variable cv_go_, cv_done_;
    unsigned long long generation_ = 0;
    size_t pending_ = 0, job_slices_ = 0, job_want_ = 0, job_slice_ = 0, job_off_ = 0;
    uint8_t* job_dst_ = nullptr;
    size_t* job_done_ = nullptr;
    bool stop_ = false;
};

// zstd through dlopen (the image ships libzstd.so.1 without headers)
struct ZstdApi {
    void* handle = nullptr;
    void* (*create)() = nullptr;
    size_t (*free_ds)(void*) = nullptr;
    size_t (*init_ds)(void*) = nullptr;
    size_t (*decompress)(void*, void*, void*) = nullptr;
    unsigned (*is_error)(size_t) = nullptr;
    bool load() {
        std::lock_guard<std::mutex> lk(g_zstd_mu);
        if (handle) return true;
        void* h = dlopen(g_zstd_path.c_str(), RTLD_NOW);
        if (!h) h = dlopen("libzstd.so.1", RTLD_NOW);
        if (!h) return false;
        create = (void* (*)())dlsym(h, "ZSTD_createDStream");
        free_ds = (size_t (*)(void*))dlsym(h, "ZSTD_freeDStream");
        init_ds = (size_t (*)(void*))dlsym(h, "ZSTD_initDStream");
        decompress = (size_t (*)(void*, void*, void*))dlsym(h, "ZSTD_decompressStream");
        is_error = (unsigned (*)(size_t))dlsym(h, "ZSTD_isError");
        if (!create || !free_ds || !init_ds || !decompress || !is_error) return false;
        handle = h;
        return true;
    }
};
ZstdApi g_zstd;

class GzipCodec : public MemberCodec {
public:
    GzipCodec() {
        std::memset(&zs_, 0, sizeof(zs_));
        ok_ = inflateInit2(&zs_, 15 + 16) == Z_OK;   // gzip wrapper
    }
    ~GzipCodec() override { if (ok_) inflateEnd(&zs_); }
    bool ok() const override { return ok_; }
    void reset() override { inflateReset(&zs_); }
    Step step(const uint8_t* in, size_t in_len, size_t& in_pos, uint8_t* out, size_t out_cap, size_t& out_pos) override {
        zs_.next_in = const_cast<Bytef*>(in + in_pos);
        zs_.avail_in = (uInt)std::min<size_t>(in_len - in_pos, 1u << 30);
        zs_.next_out = out + out_pos;
        zs_.avail_out = (uInt)std::min<size_t>(out_cap - out_pos, 1u << 30);
        const uInt in_before = zs_.avail_in, out_before = zs_.avail_out;
        const int rc = inflate(&zs_, Z_NO_FLUSH);
        const size_t used = in_before - zs_.avail_in, made = out_before - zs_.avail_out;
        in_pos += used;
        out_pos += made;
        if (rc == Z_STREAM_END) return MemberEnd;
        if (rc != Z_OK && rc != Z_BUF_ERROR) return Failed;   // corrupt data: stop, keep what was decoded
        if (rc == Z_BUF_ERROR && used == 0 && made == 0 && in_pos < in_len) return Failed;
        return More;
    }
private:
    z_stream zs_;
    bool ok_ = false;
};

class ZstdCodec : public MemberCodec {
    struct InBuf { const void* src; size_t size; size_t pos; };
    struct OutBuf { void* dst; size_t size; size_t pos; };
public:
    ZstdCodec() : ds_(g_zstd.create()) {}
    ~ZstdCodec() override { if (ds_) g_zstd.free_ds(ds_); }
    bool ok() const override { return ds_ != nullptr; }
    void reset() override { g_zstd.init_ds(ds_); }
    Step step(const uint8_t* in, size_t in_len, size_t& in_pos, uint8_t* out, size_t out_cap, size_t& out_pos) override {
        InBuf ib{in + in_pos, in_len - in_pos, 0};
        OutBuf ob{out + out_pos, out_cap - out_pos, 0};
        const size_t rc = g_zstd.decompress(ds_, &ob, &ib);
        in_pos += ib.pos;
        out_pos += ob.pos;
        if (g_zstd.is_error(rc)) return Failed;
        if (rc == 0) return MemberEnd;
        if (ib.pos == 0 && ob.pos == 0 && in_pos < in_len && out_pos < out_cap) return Failed;
        return More;
    }
private:
    void* ds_;
};

// Members / frames back to back, decoded by the calling thread.
class SequentialMemberSource : public ByteSource {
public:
    SequentialMemberSource(Packing kind, std::unique_ptr<RawFile> f, const uint8_t* head, size_t head_len)
        : kind_(kind), f_(std::move(f)), in_(1 << 20), codec_(make_member_codec(kind)) {
        std::memcpy(in_.data(), head, head_len);
        avail_ = head_len;
    }
    size_t read(uint8_t* dst, size_t cap) override {
        if (!codec_ || !codec_->ok() || done_) return 0;
        size_t produced = 0;
        while (produced < cap) {
            if (avail_ == 0 && !eof_) {
                pos_ = 0;
                avail_ = f_->read(in_.data(), in_.size());
                eof_ = avail_ == 0;
            }
            if (between_) {
                const size_t need = member_header_bytes(kind_);
                if (avail_ < need && !eof_) {
                    // pull more so that the header bytes can be inspected
                    std::memmove(in_.data(), in_.data() + pos_, avail_);
                    pos_ = 0;
                    avail_ += f_->read(in_.data() + avail_, in_.size() - avail_);
                }
                if (avail_ < need || !member_continues(kind_, in_.data() + pos_)) { done_ = true; break; }
                codec_->reset();
                between_ = false;
            }
            // (at the end of the file the decoder may still hold text that did not fit the previous destination: it is
            // called with no input until nothing comes out any more)
            size_t used = 0, made = 0;
            const MemberCodec::Step st = codec_->step(in_.data() + pos_, avail_, used, dst + produced, cap - produced, made);
            pos_ += used; avail_ -= used;
            produced += made;
            if (st == MemberCodec::MemberEnd) { between_ = true; continue; }
            if (st == MemberCodec::Failed) { done_ = true; break; }
            if (eof_ && avail_ == 0 && made == 0) { done_ = true; break; }   // truncated member
        }
        return produced;
    }
    const char* kind() const override { return kind_ == Packing::Gzip ? "gzip" : "zstd"; }
private:
    Packing kind_;
    std::unique_ptr<RawFile> f_;
    std::vector<uint8_t> in_;
    size_t pos_ = 0, avail_ = 0;
    std::unique_ptr<MemberCodec> codec_;
    bool done_ = false, between_ = false, eof_ = false;
};

}  // namespace

bool zstd_available() { return g_zstd.load(); }

std::unique_ptr<MemberCodec> make_member_codec(Packing kind) {
    if (kind == Packing::Gzip) return std::make_unique<GzipCodec>();
    if (!g_zstd.load()) return nullptr;
    return std::make_unique<ZstdCodec>();
}

void set_zstd_library_path(const std::string& path) {
    std::lock_guard<std::mutex> lk(g_zstd_mu);
    g_zstd_path = path;
}

std::unique_ptr<ByteSource> open_byte_source(const char* path, std::string& error) {
    int fd = ::open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) {
        error = std::string("cannot open ") + path + ": " + std::strerror(errno);
        return nullptr;
    }
    auto raw = std::make_unique<RawFile>(fd);
    uint8_t head[4];
    size_t got = raw->read(head, sizeof(head));
    const bool gz = got >= 2 && head[0] == 0x1f && head[1] == 0x8b;
    const bool zst = got == 4 && head[0] == 0x28 && head[1] == 0xb5 && head[2] == 0x2f && head[3] == 0xfd;
    if (zst && !g_zstd.load()) {
        error = "zstd input but libzstd could not be loaded";
        return nullptr;
    }
    if (gz || zst) {
        const Packing kind = gz ? Packing::Gzip : Packing::Zstd;
        // regular files with several members / frames: decoded by several threads (ingest_members.cpp)
        if (auto parallel = open_parallel_members(kind, fd)) return parallel;
        return std::make_unique<SequentialMemberSource>(kind, std::move(raw), head, got);
    }
    struct stat sb;
    bool regular = ::fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode);
    // one front-to-back pass: ask the kernel for aggressive read-ahead (a hint; ignored by tmpfs)
    if (regular) ::posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
    return std::make_unique<PlainSource>(std::move(raw), head, got, fd, regular);
}

std::unique_ptr<ByteSource> open_plain_range(const char* path, size_t begin, size_t end, std::string& error) {
    int fd = ::open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) {
        error = std::string("cannot open ") + path + ": " + std::strerror(errno);
        return nullptr;
    }
    ::posix_fadvise(fd, (off_t)begin, (off_t)(end - begin), POSIX_FADV_SEQUENTIAL);
    return std::make_unique<PlainSource>(std::make_unique<RawFile>(fd), fd, begin, end);
}

size_t plain_regular_file_size(const char* path) {
    int fd = ::open(path, O_RDONLY | O_CLOEXEC);
    if (fd < 0) return 0;
    struct stat sb;
    uint8_t head[4] = {0, 0, 0, 0};
    size_t size = 0;
    if (::fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode) && ::pread(fd, head, 4, 0) >= 2) {
        const bool gz = head[0] == 0x1f && head[1] == 0x8b;
        const bool zst = head[0] == 0x28 && head[1] == 0xb5 && head[2] == 0x2f && head[3] == 0xfd;
        if (!gz && !zst) size = (size_t)sb.st_size;
    }
    ::close(fd);
    return size;
}

}  // namespace gpugrep
