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
//   * gzip input is decompressed transparently (zlib), like needletail's sniffing reader;
//   * format by first byte: '>' FASTA, '@' FASTQ; anything else / empty file is an error at open;
//   * a record's sequence = its sequence line(s) with '\n' and '\r' removed, bytes otherwise untouched;
//   * FASTQ records are 4 lines; '+' separator and |qual| == |seq| are checked;
//   * iteration ends at EOF or at the first malformed record (not an error to the caller).
class FastxReader {
public:
    FastxReader() = default;
    ~FastxReader();
    FastxReader(const FastxReader&) = delete;
    FastxReader& operator=(const FastxReader&) = delete;

    // 0 on success; on failure returns non-zero and fills *err
    int open(const char* path, std::string* err);
    bool is_fastq() const { return fastq_; }

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
    std::vector<uint8_t> buf_;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false, fastq_ = false, at_line_start_ = true, started_ = false;
};

// whole-file driver behind nk_process_file (defined in nk_api.cu)
int process_file(nk_counter* h, const char* path, bool streaming, std::string* err);

}  // namespace nk
