// Host ingest: sequential byte source over a plain, gzip or zstd file.
// Replaces gzopen()/gzgets() of zstd's zlibWrapper in the reference (hyperscanner.c:189-199, built per
// utils/build_hyperscanner.sh:76-89).  Decompression stays on host threads (BASELINE.json north_star:
// "decompressed on host threads as stated ingest, not the hot path").
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>

namespace gpugrep {

class ByteSource {
public:
    virtual ~ByteSource() = default;
    // Fill dst with up to cap bytes; returns 0 at end of data.  Decode errors end the data (like gzgets() == NULL).
    virtual size_t read(uint8_t* dst, size_t cap) = 0;
    virtual const char* kind() const = 0;
};

// nullptr when the file cannot be opened (-> HYPERSCANNER_GZ_OPEN = 6, hyperscanner.c:192-195).
std::unique_ptr<ByteSource> open_byte_source(const char* path, std::string& error);

// Byte range [begin, end) of a plain regular file (one shard of a scan that is split over several GPUs).
std::unique_ptr<ByteSource> open_plain_range(const char* path, size_t begin, size_t end, std::string& error);

// Size of `path` if it is a plain (not gzip / zstd) regular file, else 0: only such files can be split by byte ranges.
size_t plain_regular_file_size(const char* path);

void set_zstd_library_path(const std::string& path);

}  // namespace gpugrep
