// Internal to the host ingest (ingest.cpp, ingest_members.cpp): one streaming decoder interface for gzip members and
// zstd frames, and the rules that decide what follows a member.  The reference reads both through gzopen()/gzgets() of
// zstd's zlibWrapper (hyperscanner.c:189-199); the rules are that wrapper's (gz_look() in gzread.c).
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>

namespace gpugrep {

enum class Packing { Gzip, Zstd };

// Streaming decoder of ONE member (gzip) / frame (zstd) at a time.
class MemberCodec {
public:
    enum Step { More = 0, MemberEnd = 1, Failed = -1 };
    virtual ~MemberCodec() = default;
    virtual bool ok() const = 0;
    virtual void reset() = 0;   // before the first byte of the next member
    // Consumes in[in_pos..in_len), produces out[out_pos..out_cap).  MemberEnd: the member ended exactly at in_pos.
    // Failed: corrupt data, or no progress although input and room were there (what was decoded so far stays valid).
    virtual Step step(const uint8_t* in, size_t in_len, size_t& in_pos, uint8_t* out, size_t out_cap, size_t& out_pos) = 0;
};

std::unique_ptr<MemberCodec> make_member_codec(Packing kind);   // nullptr: zstd library not loadable
bool zstd_available();

// Bytes the rule below needs to see.
inline size_t member_header_bytes(Packing kind) { return kind == Packing::Gzip ? 2 : 4; }

// After a member: does the data continue with another member?  gzip: the two magic bytes (anything else is trailing
// garbage that zlib ignores).  zstd: the frame magic; ANYTHING else - a skippable frame included - ends the data (the
// wrapper reports Z_STREAM_END at the end of every frame, so libzstd never gets to skip a skippable frame behind one).
inline bool member_continues(Packing kind, const uint8_t* p) {
    if (kind == Packing::Gzip) return p[0] == 0x1f && p[1] == 0x8b;
    return p[0] == 0x28 && p[1] == 0xb5 && p[2] == 0x2f && p[3] == 0xfd;
}

class ByteSource;
// Several threads decoding one regular file that holds many members / frames (ingest_members.cpp); nullptr when that
// does not apply (not a regular file, small, no helper threads to be had, GPUGREP_DECODE_THREADS=0).
std::unique_ptr<ByteSource> open_parallel_members(Packing kind, int fd);

}  // namespace gpugrep
