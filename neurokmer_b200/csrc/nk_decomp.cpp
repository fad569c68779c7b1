// nk_decomp.cpp — bzip2 / xz / zstd input for the FASTA/FASTQ reader (host only).
//
// needletail 0.6.3 (the reference's parser, Cargo.lock:981-993; default feature "compression")
// sniffs the first bytes of the input and wraps it in a decoder: gzip -> flate2 MultiGzDecoder,
// "BZh" -> bzip2 BzDecoder, FD 37 7A 58 -> XzDecoder, 28 B5 2F FD -> zstd Decoder.  gzip is handled
// in nk_fastx.cpp through zlib.  This image ships the RUNTIME libraries of the other three
// (libbz2.so.1.0, liblzma.so.5, libzstd.so.1) but not their headers, so the few entry points needed
// are declared here from the libraries' stable public ABIs and resolved with dlopen at first use;
// if a library is missing the open fails with a message that says so (no silent misparse).
// A corrupt or truncated stream ends the iteration, like a parse error does (src/utils.rs:17-20).
// Stream rules mirror the Rust wrappers: bzip2 and xz decode ONE stream (BzDecoder, XzDecoder with
// flags 0), zstd decodes every concatenated frame.
#include <dlfcn.h>
#include <unistd.h>

#include <cerrno>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "nk_host.h"

namespace nk {

namespace {

// ---- libbz2 (bzlib.h, ABI 1.0) ------------------------------------------------------------------
struct bz_stream {
    char* next_in;
    unsigned int avail_in, total_in_lo32, total_in_hi32;
    char* next_out;
    unsigned int avail_out, total_out_lo32, total_out_hi32;
    void* state;
    void* (*bzalloc)(void*, int, int);
    void (*bzfree)(void*, void*);
    void* opaque;
};
constexpr int BZ_OK = 0, BZ_STREAM_END = 4;

// ---- liblzma (lzma/base.h, ABI 5) ---------------------------------------------------------------
struct lzma_stream {
    const uint8_t* next_in;
    size_t avail_in;
    uint64_t total_in;
    uint8_t* next_out;
    size_t avail_out;
    uint64_t total_out;
    const void* allocator;
    void* internal;
    void* reserved_ptr[4];
    uint64_t reserved_int1, reserved_int2;
    size_t reserved_int3, reserved_int4;
    int reserved_enum1, reserved_enum2;
    uint64_t pad_[8];  // head-room should a later 5.x grow the struct (it must stay zero-initialised)
};
constexpr int LZMA_OK = 0, LZMA_STREAM_END = 1, LZMA_RUN = 0, LZMA_FINISH = 3;

// ---- libzstd (zstd.h, ABI 1) ----------------------------------------------------------------------
struct ZSTD_inBuffer { const void* src; size_t size, pos; };
struct ZSTD_outBuffer { void* dst; size_t size, pos; };

struct Libs {
    void* bz = nullptr; void* xz = nullptr; void* zs = nullptr;
    int (*bz_init)(bz_stream*, int, int) = nullptr;
    int (*bz_run)(bz_stream*) = nullptr;
    int (*bz_end)(bz_stream*) = nullptr;
    int (*xz_decoder)(lzma_stream*, uint64_t, uint32_t) = nullptr;
    int (*xz_code)(lzma_stream*, int) = nullptr;
    void (*xz_end)(lzma_stream*) = nullptr;
    void* (*zs_create)() = nullptr;
    size_t (*zs_init)(void*) = nullptr;
    size_t (*zs_run)(void*, ZSTD_outBuffer*, ZSTD_inBuffer*) = nullptr;
    unsigned (*zs_is_error)(size_t) = nullptr;
    size_t (*zs_free)(void*) = nullptr;
};
Libs g_libs;
std::mutex g_libs_mu;  // distinct handles may open compressed files from different threads

template <typename F>
bool sym(void* lib, const char* name, F* out) {
    *out = reinterpret_cast<F>(dlsym(lib, name));
    return *out != nullptr;
}

bool load_bz(std::string* err) {
    std::lock_guard<std::mutex> lock(g_libs_mu);
    Libs& L = g_libs;
    if (L.bz_run) return true;
    L.bz = dlopen("libbz2.so.1.0", RTLD_NOW | RTLD_LOCAL);
    if (!L.bz) L.bz = dlopen("libbz2.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!L.bz || !sym(L.bz, "BZ2_bzDecompressInit", &L.bz_init) || !sym(L.bz, "BZ2_bzDecompress", &L.bz_run) ||
        !sym(L.bz, "BZ2_bzDecompressEnd", &L.bz_end)) {
        L.bz_run = nullptr;
        if (err) *err = "bzip2 input needs libbz2.so.1.0 at run time (not found)";
        return false;
    }
    return true;
}
bool load_xz(std::string* err) {
    std::lock_guard<std::mutex> lock(g_libs_mu);
    Libs& L = g_libs;
    if (L.xz_code) return true;
    L.xz = dlopen("liblzma.so.5", RTLD_NOW | RTLD_LOCAL);
    if (!L.xz || !sym(L.xz, "lzma_stream_decoder", &L.xz_decoder) || !sym(L.xz, "lzma_code", &L.xz_code) ||
        !sym(L.xz, "lzma_end", &L.xz_end)) {
        L.xz_code = nullptr;
        if (err) *err = "xz input needs liblzma.so.5 at run time (not found)";
        return false;
    }
    return true;
}
bool load_zs(std::string* err) {
    std::lock_guard<std::mutex> lock(g_libs_mu);
    Libs& L = g_libs;
    if (L.zs_run) return true;
    L.zs = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL);
    if (!L.zs || !sym(L.zs, "ZSTD_createDStream", &L.zs_create) || !sym(L.zs, "ZSTD_initDStream", &L.zs_init) ||
        !sym(L.zs, "ZSTD_decompressStream", &L.zs_run) || !sym(L.zs, "ZSTD_isError", &L.zs_is_error) ||
        !sym(L.zs, "ZSTD_freeDStream", &L.zs_free)) {
        L.zs_run = nullptr;
        if (err) *err = "zstd input needs libzstd.so.1 at run time (not found)";
        return false;
    }
    return true;
}

}  // namespace

struct StreamDecoder::Impl {
    int kind = 0;  // 1 bzip2, 2 xz, 3 zstd
    int fd = -1;
    std::vector<uint8_t> in;
    size_t in_pos = 0, in_end = 0;
    bool in_eof = false, done = false;
    bz_stream bz{};
    lzma_stream xz{};
    void* zs = nullptr;
    bool zs_mid_frame = false;

    bool refill() {
        if (in_eof) return false;
        for (;;) {
            const ssize_t n = ::read(fd, in.data(), in.size());
            if (n < 0 && errno == EINTR) continue;
            if (n <= 0) { in_eof = true; return false; }
            in_pos = 0;
            in_end = (size_t)n;
            return true;
        }
    }
};

StreamDecoder::StreamDecoder() = default;

StreamDecoder::~StreamDecoder() {
    if (!p_) return;
    if (p_->kind == 1 && g_libs.bz_end) g_libs.bz_end(&p_->bz);
    if (p_->kind == 2 && g_libs.xz_end) g_libs.xz_end(&p_->xz);
    if (p_->kind == 3 && p_->zs && g_libs.zs_free) g_libs.zs_free(p_->zs);
    delete p_;
}

int StreamDecoder::sniff(const unsigned char* magic, size_t n) {
    if (n >= 3 && magic[0] == 'B' && magic[1] == 'Z' && magic[2] == 'h') return 1;
    if (n >= 4 && magic[0] == 0xfd && magic[1] == '7' && magic[2] == 'z' && magic[3] == 'X') return 2;
    if (n >= 4 && magic[0] == 0x28 && magic[1] == 0xb5 && magic[2] == 0x2f && magic[3] == 0xfd) return 3;
    return 0;
}

bool StreamDecoder::start(int kind, int fd, std::string* err) {
    if (kind == 1 && !load_bz(err)) return false;
    if (kind == 2 && !load_xz(err)) return false;
    if (kind == 3 && !load_zs(err)) return false;
    p_ = new Impl;
    p_->kind = kind;
    p_->fd = fd;
    p_->in.resize(1u << 20);
    bool ok = true;
    if (kind == 1) ok = g_libs.bz_init(&p_->bz, 0, 0) == BZ_OK;
    if (kind == 2) ok = g_libs.xz_decoder(&p_->xz, ~0ull, 0) == LZMA_OK;  // XzDecoder::new: no memory limit, flags 0
    if (kind == 3) {
        p_->zs = g_libs.zs_create();
        ok = p_->zs && !g_libs.zs_is_error(g_libs.zs_init(p_->zs));
    }
    if (!ok) {
        if (err) *err = "cannot start the decompressor";
        p_->kind = 0;
        delete p_;
        p_ = nullptr;
        return false;
    }
    return true;
}

// decompressed bytes into dst (up to cap); 0 = end of data (or a corrupt stream: the iteration ends)
size_t StreamDecoder::read(uint8_t* dst, size_t cap) {
    Impl& s = *p_;
    size_t w = 0;
    while (w == 0 && !s.done) {
        if (s.in_pos == s.in_end && !s.in_eof) s.refill();
        const size_t avail = s.in_end - s.in_pos;
        if (s.kind == 1) {
            if (avail == 0) { s.done = true; break; }  // truncated
            s.bz.next_in = reinterpret_cast<char*>(s.in.data() + s.in_pos);
            s.bz.avail_in = (unsigned)avail;
            s.bz.next_out = reinterpret_cast<char*>(dst);
            s.bz.avail_out = (unsigned)std::min<size_t>(cap, 1u << 30);
            const unsigned out0 = s.bz.avail_out;
            const int rc = g_libs.bz_run(&s.bz);
            s.in_pos += avail - s.bz.avail_in;
            w = out0 - s.bz.avail_out;
            if (rc == BZ_STREAM_END || rc != BZ_OK) s.done = true;  // one stream (bzip2::read::BzDecoder)
        } else if (s.kind == 2) {
            s.xz.next_in = s.in.data() + s.in_pos;
            s.xz.avail_in = avail;
            s.xz.next_out = dst;
            s.xz.avail_out = cap;
            const int rc = g_libs.xz_code(&s.xz, avail == 0 ? LZMA_FINISH : LZMA_RUN);
            s.in_pos += avail - s.xz.avail_in;
            w = cap - s.xz.avail_out;
            if (rc != LZMA_OK) s.done = true;  // LZMA_STREAM_END, or an error / truncated input
        } else {
            if (avail == 0) { s.done = true; break; }  // end of input (inside a frame: truncated)
            ZSTD_inBuffer ib{s.in.data() + s.in_pos, avail, 0};
            ZSTD_outBuffer ob{dst, cap, 0};
            const size_t rc = g_libs.zs_run(s.zs, &ob, &ib);
            s.in_pos += ib.pos;
            w = ob.pos;
            if (g_libs.zs_is_error(rc)) s.done = true;
            // rc == 0: a frame ended; the next call starts the following frame, if any
        }
    }
    return w;
}

}  // namespace nk
