// nk_host.h — host-side helpers behind the C ABI (no device code).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

struct nk_counter;

namespace nk {

// pack_kmer — reference src/utils.rs:26-39 (non-ACGT bytes skipped, no mask)
uint64_t host_pack_kmer(const uint8_t* kmer, uint64_t len);

// Incremental FASTA/FASTQ record reader with needletail's record rules as the
// reference uses them (src/utils.rs:9-24):
//   * gzip (zlib), bzip2, xz and zstd input is decompressed transparently, sniffed from the magic
//     bytes like needletail's reader;
//   * format by first byte: '>' FASTA, '@' FASTQ; anything else / empty file is an error at open;
//   * a record's sequence = its sequence line(s) with '\n' and '\r' removed, bytes otherwise untouched;
//   * FASTQ records are 4 lines; '+' separator and |qual| == |seq| are checked;
//   * iteration ends at EOF or at the first malformed record (not an error to the caller).
// bzip2 / xz / zstd stream decoder over a file descriptor (nk_decomp.cpp: the libraries are resolved
// with dlopen at first use).  sniff() -> 0 none, 1 bzip2, 2 xz, 3 zstd.
class StreamDecoder {
public:
    StreamDecoder();
    ~StreamDecoder();
    StreamDecoder(const StreamDecoder&) = delete;
    StreamDecoder& operator=(const StreamDecoder&) = delete;
    static int sniff(const unsigned char* magic, size_t n);
    bool start(int kind, int fd, std::string* err);
    size_t read(uint8_t* dst, size_t cap);

private:
    struct Impl;
    Impl* p_ = nullptr;
};

class FastxReader {
public:
    FastxReader() = default;
    ~FastxReader();
    FastxReader(const FastxReader&) = delete;
    FastxReader& operator=(const FastxReader&) = delete;

    // 0 on success; on failure returns non-zero and fills *err
    int open(const char* path, std::string* err);
    bool is_fastq() const { return fastq_; }
    bool is_gzip() const { return gz_ != nullptr; }
    bool is_compressed() const { return gz_ != nullptr || dec_ != nullptr; }

    // Advance to the next record. false: EOF or malformed record (iteration over).
    bool next_record();
    // Copy up to `cap` sequence bytes of the current record to dst; returns bytes written.
    // *done = true once the record's sequence is exhausted.
    size_t read_seq(uint8_t* dst, size_t cap, bool* done);
    // FASTQ only: after the sequence is exhausted, validate '+' line and quality length.
    // false: malformed (the record must be discarded and iteration ends).
    bool finish_record(uint64_t seq_len);

private:
    bool fill();
    int peek();  // next byte or -1
    void skip_line();

    int fd_ = -1;
    void* gz_ = nullptr;  // gzFile when the input is gzip-compressed
    StreamDecoder* dec_ = nullptr;  // bzip2 / xz / zstd
    std::vector<uint8_t> buf_;            // read buffer of compressed / non-regular input
    const uint8_t* base_ = nullptr;       // the current buffer: buf_.data(), or the mapping of a plain file
    void* map_ = nullptr;
    size_t map_size_ = 0;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false, fastq_ = false, at_line_start_ = true, started_ = false;
};

// FASTA sequence data: dst <- src[0, n) without '\n' / '\r', stopping at the first '>' that starts a line
// (src[0] itself is data).  *consumed = input bytes taken; returns bytes written.  One fused pass
// (AVX-512 VBMI2 compress where the CPU has it).
size_t strip_until_header(uint8_t* dst, const uint8_t* src, size_t n, size_t* consumed);

// ---- parallel FASTA ingest (plain files): the file is cut into windows that several host threads
// strip (headers, '\n', '\r') concurrently; see process_file in nk_api.cu -------------------------
constexpr unsigned kFastaSlack = 64u << 10;  // a window ends at most this far past its nominal size
enum FastaStart { FA_LINE_START = 0, FA_MID_HEADER = 1, FA_MID_SEQ = 2 };
struct FastaWindowPlan {
    size_t ws = 0, we = 0;   // file byte range [ws, we)
    int start_state = FA_LINE_START;
};
// Next window after `ws` of nominally `window` bytes: ends at a line start when a '\n' is found within
// `window` bytes past the nominal end, else mid-line.  `state_in` is the start state of THIS window;
// the start state of the following window is returned in *state_next.
FastaWindowPlan fasta_plan_window(const uint8_t* file, size_t size, size_t ws, size_t window, int state_in, int* state_next);
// Strip one window into `data` (capacity >= we - ws): sequence bytes in file order; rec_starts gets the
// offset (into data) of every record that STARTS in this window (a header line was seen).
void fasta_parse_window(const uint8_t* file, const FastaWindowPlan& w, uint8_t* data, size_t* fill,
                        std::vector<uint64_t>* rec_starts);

// ---- parallel FASTQ ingest (plain files) ------------------------------------------------------------
// A record is exactly 4 lines, so the line number modulo 4 identifies record starts: newlines are
// counted per fixed byte range in parallel, prefix-summed, and every range's first record start is
// the first line start whose global line number is a multiple of 4.  (If an earlier record is
// malformed the phase of later windows is meaningless — but iteration ends at the first malformed
// record, so they are discarded anyway.)
size_t fastq_count_newlines(const uint8_t* file, size_t a, size_t b);
// first record start at or after byte `s`, given the number of '\n' before `s`
size_t fastq_first_record_start(const uint8_t* file, size_t size, size_t s, uint64_t newlines_before);
// Parse the records of [a, b) (a is a record start, b a record start or EOF): sequence bytes are
// appended to data, offsets gets one entry per record end.  Returns false at the first malformed
// record (everything before it is kept).
bool fastq_parse_window(const uint8_t* file, size_t a, size_t b, uint8_t* data, size_t cap, size_t* fill,
                        std::vector<uint64_t>* offsets);

// ---- 2-bit packer of the pre-packed input path (nk_pack.cpp; layout in include/neurokmer.h) ---------
// bases [p0, p1) -> codes / other (p0 a multiple of 64; whole words are written).  Returns the number
// of non-ACGT bytes.  body: 0 best available, 1 portable, 2 AVX2, 3 AVX-512BW.
uint64_t host_pack_range(const uint8_t* bases, uint64_t p0, uint64_t p1, uint32_t* codes, uint32_t* other, int body);
// whole array on `threads` host threads (<= 0: all hardware threads)
uint64_t host_pack_bases(const uint8_t* bases, uint64_t n, uint32_t* codes, uint32_t* other, int threads, int body);
int host_pack_body_available(int which);

// whole-file driver behind nk_process_file (defined in nk_api.cu)
int process_file(nk_counter* h, const char* path, bool streaming, std::string* err, bool ingest_only);

}  // namespace nk
