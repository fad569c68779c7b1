// nk_fastx.cpp — FASTA/FASTQ record reader (host).  Stands in for needletail 0.6.3's
// parse_fastx_file as the reference drives it (src/utils.rs:9-24): see nk_host.h.
#include "nk_host.h"

#include <algorithm>
#include <cerrno>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <immintrin.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <zlib.h>

#include <cstdlib>

namespace nk {

uint64_t host_pack_kmer(const uint8_t* kmer, uint64_t len) {
    uint64_t packed = 0;
    for (uint64_t i = 0; i < len; ++i) {
        uint64_t bits;
        switch (kmer[i]) {
            case 'A': case 'a': bits = 0; break;
            case 'C': case 'c': bits = 1; break;
            case 'G': case 'g': bits = 2; break;
            case 'T': case 't': bits = 3; break;
            default: continue;
        }
        packed = (packed << 2) | bits;
    }
    return packed;
}

// ---- FASTA sequence data: dst <- src without '\n' / '\r', up to the next header -------------------------
// Stripping the terminators of 60-column lines is what bounds FASTA ingest: one memchr + one memcpy per line
// ran at ~1.5 GB/s, a separate search for the next header was a second pass over the data.  One fused pass:
// copy src[0, n) without '\n' and '\r' ('\r' never occurs inside FASTA text) and stop at the first '>' that
// starts a line (previous byte '\n'; src[0] itself is data by contract).  With AVX-512 VBMI2 a 64-byte block
// is compared against the three bytes, compressed in a register and stored with a length mask; otherwise a
// per-line loop runs.  *consumed = input bytes taken (n, or the index of the header); returns bytes written.
namespace {

size_t strip_generic(uint8_t* dst, const uint8_t* src, size_t n, size_t* consumed) {
    size_t w = 0, p = 0;
    while (p < n) {
        if (p > 0 && src[p] == '>') break;  // p > 0 here means a line start: the previous byte was '\n'
        const uint8_t* nl = (const uint8_t*)memchr(src + p, '\n', n - p);
        const size_t len = nl ? (size_t)(nl - (src + p)) : n - p;
        const uint8_t* s = src + p;
        if (len > 0 && s[len - 1] != '\r' && !memchr(s, '\r', len)) {
            memcpy(dst + w, s, len);
            w += len;
        } else {
            for (size_t i = 0; i < len; ++i)
                if (s[i] != '\r') dst[w++] = s[i];
        }
        p += len + (nl ? 1 : 0);
    }
    *consumed = p;
    return w;
}

__attribute__((target("avx512f,avx512bw,avx512vbmi2"))) size_t strip_avx512(uint8_t* dst, const uint8_t* src, size_t n,
                                                                             size_t* consumed) {
    const __m512i lf = _mm512_set1_epi8('\n'), cr = _mm512_set1_epi8('\r'), gt = _mm512_set1_epi8('>');
    size_t w = 0, p = 0;
    unsigned long long prev_lf = 0;  // bit 0: the byte before this block is '\n' (src[0] is data: starts clear)
    for (; p + 64 <= n; p += 64) {
        const __m512i v = _mm512_loadu_si512(src + p);
        const __mmask64 m_lf = _mm512_cmpeq_epi8_mask(v, lf);
        __mmask64 keep = ~(m_lf | _mm512_cmpeq_epi8_mask(v, cr));
        const __mmask64 hdr = _mm512_cmpeq_epi8_mask(v, gt) & ((m_lf << 1) | prev_lf);
        if (hdr) {  // a header starts inside this block: take the bytes before it and stop
            const unsigned first = (unsigned)__builtin_ctzll(hdr);
            keep &= (1ull << first) - 1;  // first < 64
            const unsigned cnt = (unsigned)__builtin_popcountll(keep);
            _mm512_mask_storeu_epi8(dst + w, (1ull << cnt) - 1, _mm512_maskz_compress_epi8(keep, v));  // cnt < 64
            *consumed = p + first;
            return w + cnt;
        }
        const unsigned cnt = (unsigned)__builtin_popcountll(keep);
        _mm512_mask_storeu_epi8(dst + w, cnt == 64 ? ~0ull : ((1ull << cnt) - 1), _mm512_maskz_compress_epi8(keep, v));
        w += cnt;
        prev_lf = m_lf >> 63;
    }
    for (; p < n; ++p) {
        const uint8_t b = src[p];
        if (b == '>' && p > 0 && src[p - 1] == '\n') break;
        if (b != '\n' && b != '\r') dst[w++] = b;
    }
    *consumed = p;
    return w;
}

using StripFn = size_t (*)(uint8_t*, const uint8_t*, size_t, size_t*);
StripFn pick_strip() {
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi2") &&
        !getenv("NK_NO_SIMD_STRIP"))
        return strip_avx512;
    return strip_generic;
}

}  // namespace

size_t strip_until_header(uint8_t* dst, const uint8_t* src, size_t n, size_t* consumed) {
    static const StripFn fn = pick_strip();
    return fn(dst, src, n, consumed);
}

FastxReader::~FastxReader() {
    delete dec_;
    if (map_) munmap(map_, map_size_);
    if (gz_) gzclose((gzFile)gz_);  // also closes the descriptor
    else if (fd_ >= 0) ::close(fd_);
}

bool FastxReader::fill() {
    if (eof_) return false;
    if (map_) { eof_ = true; return false; }  // a mapped file is one buffer: [0, size) was handed out at open()
    pos_ = 0;
    end_ = 0;
    for (;;) {
        ssize_t n = dec_ ? (ssize_t)dec_->read(buf_.data(), buf_.size())
                  : gz_  ? (ssize_t)gzread((gzFile)gz_, buf_.data(), (unsigned)buf_.size())
                         : ::read(fd_, buf_.data(), buf_.size());
        if (n < 0 && !gz_ && !dec_ && errno == EINTR) continue;
        if (n <= 0) { eof_ = true; return false; }  // a truncated / corrupt gzip stream ends the iteration
        end_ = (size_t)n;
        return true;
    }
}

int FastxReader::peek() {
    if (pos_ == end_ && !fill()) return -1;
    return base_[pos_];
}

void FastxReader::skip_line() {
    for (;;) {
        if (pos_ == end_ && !fill()) return;
        const uint8_t* p = (const uint8_t*)memchr(base_ + pos_, '\n', end_ - pos_);
        if (p) { pos_ = (size_t)(p - base_) + 1; at_line_start_ = true; return; }
        pos_ = end_;
    }
}

int FastxReader::open(const char* path, std::string* err) {
    fd_ = ::open(path, O_RDONLY);
    if (fd_ < 0) {
        if (err) *err = std::string("cannot open ") + path + ": " + strerror(errno);
        return 1;
    }
    struct stat st;
    const bool regular = fstat(fd_, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0;
    // needletail sniffs the compression format from the magic bytes: gzip through zlib here,
    // bzip2 / xz / zstd through nk_decomp.cpp
    unsigned char magic[4] = {0, 0, 0, 0};
    const ssize_t got = ::pread(fd_, magic, 4, 0);
    if (got >= 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
        gz_ = gzdopen(fd_, "rb");
        if (!gz_) { if (err) *err = std::string(path) + ": cannot start gzip decompression"; return 1; }
        gzbuffer((gzFile)gz_, 1u << 20);
    } else if (const int kind = StreamDecoder::sniff(magic, got > 0 ? (size_t)got : 0)) {
        dec_ = new StreamDecoder;
        std::string why;
        if (!dec_->start(kind, fd_, &why)) {
            if (err) *err = std::string(path) + ": " + why;
            return 1;
        }
    }
    if (!gz_ && !dec_ && regular) {
        // plain file: map it (no read() copy); the whole file is the reader's one buffer
        // small files are pre-faulted in one go (demand faults cost as much as the read() copy they replace)
        const int populate = st.st_size <= (1ll << 30) ? MAP_POPULATE : 0;
        void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE | populate, fd_, 0);
        if (m != MAP_FAILED) {
            map_ = m;
            map_size_ = (size_t)st.st_size;
            madvise(m, map_size_, MADV_SEQUENTIAL);
            base_ = static_cast<const uint8_t*>(m);
            pos_ = 0;
            end_ = map_size_;
        }
    }
    if (!map_) {
        buf_.resize(4u << 20);
        base_ = buf_.data();
    }
    const int c = peek();
    if (c < 0) { if (err) *err = std::string(path) + ": empty file"; return 1; }
    if (c == '>') fastq_ = false;
    else if (c == '@') fastq_ = true;
    else { if (err) *err = std::string(path) + ": not FASTA/FASTQ (first byte must be '>' or '@')"; return 1; }
    at_line_start_ = true;
    return 0;
}

bool FastxReader::next_record() {
    const int c = peek();
    if (c < 0) return false;
    if (fastq_) {
        if (c != '@') return false;  // malformed: iteration ends (src/utils.rs:17-20)
    } else {
        if (c != '>') {
            // only reachable right after open() (first byte checked) or after read_seq stopped at '>'
            return false;
        }
    }
    skip_line();  // header
    started_ = true;
    return true;
}

size_t FastxReader::read_seq(uint8_t* dst, size_t cap, bool* done) {
    size_t w = 0;
    *done = false;
    // FASTA: whole regions between headers are stripped at once (a line-start '>' ends the record)
    while (!fastq_ && w < cap) {
        if (pos_ == end_ && !fill()) { *done = true; return w; }
        if (at_line_start_ && base_[pos_] == '>') { *done = true; return w; }
        const uint8_t* s = base_ + pos_;
        const size_t n = std::min(end_ - pos_, cap - w);
        size_t q = 0;  // s[0] is data: not '>' or not at a line start
        w += strip_until_header(dst + w, s, n, &q);
        at_line_start_ = s[q - 1] == '\n';
        pos_ += q;
        if (q < n) { *done = true; return w; }
    }
    while (fastq_ && w < cap) {
        if (pos_ == end_ && !fill()) { *done = true; return w; }
        if (at_line_start_) {
            at_line_start_ = false;
        }
        // copy up to end of line
        const uint8_t* s = base_ + pos_;
        size_t avail = end_ - pos_;
        if (avail > cap - w) avail = cap - w;
        const uint8_t* nl = (const uint8_t*)memchr(s, '\n', avail);
        const size_t n = nl ? (size_t)(nl - s) : avail;
        // strip '\r' (only ever a line terminator in text FASTA/FASTQ); the common line has none: memcpy
        if (n > 0 && s[n - 1] != '\r' && !memchr(s, '\r', n)) {
            memcpy(dst + w, s, n);
            w += n;
        } else {
            for (size_t i = 0; i < n; ++i) {
                const uint8_t b = s[i];
                if (b != '\r') dst[w++] = b;
            }
        }
        pos_ += n;
        if (nl) {
            pos_ += 1;
            at_line_start_ = true;
            if (fastq_) { *done = true; return w; }  // FASTQ sequence is one line
        }
    }
    // cap reached: the record may or may not be finished
    if (!fastq_) {
        if (pos_ == end_ && !fill()) { *done = true; return w; }
        if (at_line_start_ && base_[pos_] == '>') *done = true;
    }
    return w;
}

bool FastxReader::finish_record(uint64_t seq_len) {
    if (!fastq_) return true;
    if (peek() != '+') return false;
    skip_line();
    // quality line: count bytes up to '\n' (excluding '\r')
    uint64_t q = 0;
    bool saw_nl = false;
    for (;;) {
        if (pos_ == end_ && !fill()) break;
        const uint8_t* s = base_ + pos_;
        const size_t avail = end_ - pos_;
        const uint8_t* nl = (const uint8_t*)memchr(s, '\n', avail);
        const size_t n = nl ? (size_t)(nl - s) : avail;
        q += n;
        if (n > 0 && s[n - 1] == '\r' && nl) q -= 1;
        pos_ += n;
        if (nl) { pos_ += 1; saw_nl = true; break; }
    }
    (void)saw_nl;
    at_line_start_ = true;
    return q == seq_len;
}

FastaWindowPlan fasta_plan_window(const uint8_t* file, size_t size, size_t ws, size_t window, int state_in, int* state_next) {
    FastaWindowPlan w;
    w.ws = ws;
    w.start_state = state_in;
    size_t e = ws + window;
    if (e >= size) {
        w.we = size;
        *state_next = FA_LINE_START;
        return w;
    }
    // prefer to end at a line start: look for the next '\n' at or after the nominal end (bounded, so
    // that a window never exceeds window + kFastaSlack bytes)
    const size_t span = std::min(std::min(window, (size_t)kFastaSlack), size - e);
    const uint8_t* nl = (const uint8_t*)memchr(file + e - 1, '\n', span + 1);
    if (nl) {
        w.we = (size_t)(nl - file) + 1;
        *state_next = FA_LINE_START;
        return w;
    }
    // a line longer than the window: cut mid-line; is that line a header?
    w.we = e;
    const uint8_t* prev = (const uint8_t*)memrchr(file + ws, '\n', e - ws);
    if (prev) {
        *state_next = (prev + 1 < file + e && prev[1] == '>') ? FA_MID_HEADER : FA_MID_SEQ;
    } else if (state_in == FA_LINE_START) {
        *state_next = file[ws] == '>' ? FA_MID_HEADER : FA_MID_SEQ;
    } else {
        *state_next = state_in;  // still inside the same overlong line
    }
    return w;
}

void fasta_parse_window(const uint8_t* file, const FastaWindowPlan& w, uint8_t* data, size_t* fill,
                        std::vector<uint64_t>* rec_starts) {
    size_t p = w.ws, out = 0;
    rec_starts->clear();
    int state = w.start_state;
    while (p < w.we) {
        if (state == FA_MID_HEADER || (state == FA_LINE_START && file[p] == '>')) {
            // a header line (or the rest of one that started in an earlier window): skip it
            if (state == FA_LINE_START) rec_starts->push_back(out);
            const uint8_t* nl = (const uint8_t*)memchr(file + p, '\n', w.we - p);
            p = nl ? (size_t)(nl - file) + 1 : w.we;
            state = FA_LINE_START;
            continue;
        }
        // sequence data up to the next line-start '>' (file[p] itself is data: mid-line, or not '>')
        size_t took = 0;
        out += strip_until_header(data + out, file + p, w.we - p, &took);
        p += took;
        state = FA_LINE_START;  // either at a header or at the end of the window
    }
    *fill = out;
}

size_t fastq_count_newlines(const uint8_t* file, size_t a, size_t b) {
    size_t n = 0;
    const uint8_t* p = file + a;
    const uint8_t* end = file + b;
    while (p < end) {
        const uint8_t* q = (const uint8_t*)memchr(p, '\n', (size_t)(end - p));
        if (!q) break;
        ++n;
        p = q + 1;
    }
    return n;
}

size_t fastq_first_record_start(const uint8_t* file, size_t size, size_t s, uint64_t newlines_before) {
    if (s >= size) return size;
    const bool at_line_start = s == 0 || file[s - 1] == '\n';
    // index of the line that contains byte s == number of newlines before s
    uint64_t line = newlines_before;
    size_t p = s;
    if (!at_line_start) {  // move to the start of the next line
        const uint8_t* q = (const uint8_t*)memchr(file + p, '\n', size - p);
        if (!q) return size;
        p = (size_t)(q - file) + 1;
        ++line;
    }
    while (line % 4 != 0) {
        if (p >= size) return size;
        const uint8_t* q = (const uint8_t*)memchr(file + p, '\n', size - p);
        if (!q) return size;
        p = (size_t)(q - file) + 1;
        ++line;
    }
    return p;
}

bool fastq_parse_window(const uint8_t* file, size_t a, size_t b, uint8_t* data, size_t cap, size_t* fill,
                        std::vector<uint64_t>* offsets) {
    size_t p = a, out = 0;
    offsets->clear();
    offsets->push_back(0);
    auto line = [&](size_t from, size_t* end, size_t* next) -> bool {  // [from, *end) without the terminator
        if (from >= b) return false;
        const uint8_t* q = (const uint8_t*)memchr(file + from, '\n', b - from);
        const size_t e = q ? (size_t)(q - file) : b;
        *next = q ? e + 1 : b;
        *end = (e > from && file[e - 1] == '\r') ? e - 1 : e;
        return true;
    };
    while (p < b) {
        size_t e0, e1, e2, e3, n1, n2, n3, n4;
        if (file[p] != '@' || !line(p, &e0, &n1)) break;                         // header
        if (!line(n1, &e1, &n2)) break;                                          // sequence
        if (!line(n2, &e2, &n3) || e2 == n2 || file[n2] != '+') break;           // separator
        if (!line(n3, &e3, &n4)) break;                                          // quality
        const size_t len = e1 - n1;
        if (out + len > cap) break;
        // like the serial reader: every '\r' of the sequence line is dropped, and the QUALITY length
        // (trailing '\r' trimmed) must equal the number of bytes kept
        const size_t before = out;
        if (len && memchr(file + n1, '\r', len)) {
            for (size_t i = 0; i < len; ++i)
                if (file[n1 + i] != '\r') data[out++] = file[n1 + i];
        } else {
            memcpy(data + out, file + n1, len);
            out += len;
        }
        if (e3 - n3 != out - before) { out = before; break; }                    // |qual| must equal |seq|
        offsets->push_back(out);
        p = n4;
    }
    *fill = out;
    return p >= b;  // false: stopped at a malformed (or truncated) record
}

}  // namespace nk
