// nk_pack.cpp — host-side 2-bit packer for the pre-packed input path (no device code).
//
// The packed form is what the count kernel's windowing stage builds from ASCII on the fly
// (nk_device.cuh: convert16), moved to the producer so that 4x fewer bytes cross PCIe:
//   codes[i] (u32)  bases 16i .. 16i+15, base 16i in bits 31:30, codes A,a=0 C,c=1 G,g=2 T,t=3
//                   (base_to_bits, reference src/models.rs:231-239); any other byte -> 0
//   other[w] (u32)  bit (p & 31) of word p >> 5 is set iff byte p is NOT one of ACGTacgt.  Such a
//                   base is code 0 on BOTH strands in canonical mode (src/models.rs:237,249) and
//                   is skipped by pack_kmer (src/utils.rs:35); the kernels need the bit for both.
// Bits past the last base of the last word are zero.
//
// Three bodies (AVX-512BW, AVX2, portable), chosen once at run time; identical output
// (tests/test_abi.py compares all available bodies against a numpy restatement).
#include <immintrin.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "nk_host.h"

namespace nk {

namespace {

// ---- portable body ---------------------------------------------------------------------------------
struct PackLut {
    uint8_t code[256];
    uint8_t other[256];
    PackLut() {
        for (int b = 0; b < 256; ++b) {
            code[b] = 0;
            other[b] = 1;
        }
        const char* up = "ACGT";
        const char* lo = "acgt";
        for (int c = 0; c < 4; ++c) {
            code[(uint8_t)up[c]] = (uint8_t)c;
            code[(uint8_t)lo[c]] = (uint8_t)c;
            other[(uint8_t)up[c]] = 0;
            other[(uint8_t)lo[c]] = 0;
        }
    }
};
const PackLut kLut;

// bases [p0, p1) of `bases`, p0 a multiple of 32; writes whole words (the last one zero-padded)
uint64_t pack_scalar(const uint8_t* bases, uint64_t p0, uint64_t p1, uint32_t* codes, uint32_t* other) {
    uint64_t n_other = 0;
    for (uint64_t p = p0; p < p1; p += 16) {
        const uint64_t e = std::min<uint64_t>(p + 16, p1);
        uint32_t w = 0, x = 0;
        for (uint64_t q = p; q < e; ++q) {
            const uint8_t b = bases[q];
            w |= (uint32_t)kLut.code[b] << (30u - 2u * (unsigned)(q - p));
            x |= (uint32_t)kLut.other[b] << (unsigned)(q - p);
        }
        codes[p >> 4] = w;
        n_other += (uint64_t)__builtin_popcount(x);
        if (other) {
            if ((p & 16) == 0) other[p >> 5] = x;
            else other[p >> 5] |= x << 16;
        }
    }
    return n_other;
}

// ---- AVX2 body: 32 bases per iteration -------------------------------------------------------------
__attribute__((target("avx2"))) uint64_t pack_avx2(const uint8_t* bases, uint64_t p0, uint64_t p1, uint32_t* codes,
                                                  uint32_t* other) {
    // expected upper-case letter per low nibble: 1->'A' 3->'C' 7->'G' 4->'T', anything else never matches
    const __m256i tbl = _mm256_setr_epi8(0x7F, 0x41, 0x7F, 0x43, 0x54, 0x7F, 0x7F, 0x47, 0x7F, 0x7F, 0x7F, 0x7F, 0x7F, 0x7F,
                                         0x7F, 0x7F, 0x7F, 0x41, 0x7F, 0x43, 0x54, 0x7F, 0x7F, 0x47, 0x7F, 0x7F, 0x7F, 0x7F,
                                         0x7F, 0x7F, 0x7F, 0x7F);
    const __m256i m0f = _mm256_set1_epi8(0x0F), mdf = _mm256_set1_epi8((char)0xDF), m03 = _mm256_set1_epi8(3);
    const __m256i mul_b = _mm256_set1_epi16(0x0104);  // bytes (4, 1): first base of a pair is the high one
    const __m256i mul_w = _mm256_set1_epi32(0x00010010);  // words (16, 1)
    // byte 0 of dwords 3,2,1,0 -> one little-endian u32 whose top byte is the first four bases
    const __m256i gather = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 12, 8, 4, 0, -1, -1,
                                            -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    uint64_t n_other = 0;
    uint64_t p = p0;
    for (; p + 32 <= p1; p += 32) {
        const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(bases + p));
        const __m256i valid = _mm256_cmpeq_epi8(_mm256_shuffle_epi8(tbl, _mm256_and_si256(b, m0f)), _mm256_and_si256(b, mdf));
        const __m256i h1 = _mm256_srli_epi16(b, 1), h2 = _mm256_srli_epi16(b, 2);
        const __m256i code = _mm256_and_si256(_mm256_and_si256(_mm256_xor_si256(h1, h2), m03), valid);
        const __m256i nib = _mm256_maddubs_epi16(code, mul_b);
        const __m256i byt = _mm256_madd_epi16(nib, mul_w);
        const __m256i g = _mm256_shuffle_epi8(byt, gather);
        codes[p >> 4] = (uint32_t)_mm256_extract_epi32(g, 0);
        codes[(p >> 4) + 1] = (uint32_t)_mm256_extract_epi32(g, 4);
        const uint32_t x = ~(uint32_t)_mm256_movemask_epi8(valid);
        n_other += (uint64_t)__builtin_popcount(x);
        if (other) other[p >> 5] = x;
    }
    if (p < p1) n_other += pack_scalar(bases, p, p1, codes, other);
    return n_other;
}

// ---- AVX-512BW body: 64 bases per iteration --------------------------------------------------------
__attribute__((target("avx512f,avx512bw"))) uint64_t pack_avx512(const uint8_t* bases, uint64_t p0, uint64_t p1,
                                                                 uint32_t* codes, uint32_t* other) {
    const __m512i tbl = _mm512_broadcast_i32x4(
        _mm_setr_epi8(0x7F, 0x41, 0x7F, 0x43, 0x54, 0x7F, 0x7F, 0x47, 0x7F, 0x7F, 0x7F, 0x7F, 0x7F, 0x7F, 0x7F, 0x7F));
    const __m512i m0f = _mm512_set1_epi8(0x0F), mdf = _mm512_set1_epi8((char)0xDF), m03 = _mm512_set1_epi8(3);
    const __m512i mul_b = _mm512_set1_epi16(0x0104);
    const __m512i mul_w = _mm512_set1_epi32(0x00010010);
    const __m512i gather = _mm512_broadcast_i32x4(_mm_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1));
    const __m512i lanes = _mm512_setr_epi32(0, 4, 8, 12, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    uint64_t n_other = 0;
    uint64_t p = p0;
    for (; p + 64 <= p1; p += 64) {
        // one core streams ~5.6 GB/s of a large array through this loop on its own; asking for the line 2 KiB ahead:
        // 8.3 GB/s (512 B / 1 KiB / 2 KiB / 4 KiB ahead: 6.4 / 7.4 / 8.3 / 8.5 GB/s, single thread, 96 MB)
        _mm_prefetch(reinterpret_cast<const char*>(bases + p + 2048), _MM_HINT_T0);
        const __m512i b = _mm512_loadu_si512(bases + p);
        const __mmask64 valid = _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(tbl, _mm512_and_si512(b, m0f)), _mm512_and_si512(b, mdf));
        const __m512i h1 = _mm512_srli_epi16(b, 1), h2 = _mm512_srli_epi16(b, 2);
        const __m512i code = _mm512_maskz_mov_epi8(valid, _mm512_and_si512(_mm512_xor_si512(h1, h2), m03));
        const __m512i nib = _mm512_maddubs_epi16(code, mul_b);
        const __m512i byt = _mm512_madd_epi16(nib, mul_w);
        const __m512i g = _mm512_permutexvar_epi32(lanes, _mm512_shuffle_epi8(byt, gather));
        _mm_storeu_si128(reinterpret_cast<__m128i*>(codes + (p >> 4)), _mm512_castsi512_si128(g));
        const uint64_t x = ~(uint64_t)valid;
        n_other += (uint64_t)__builtin_popcountll(x);
        if (other) {
            other[p >> 5] = (uint32_t)x;
            other[(p >> 5) + 1] = (uint32_t)(x >> 32);
        }
    }
    if (p < p1) n_other += pack_scalar(bases, p, p1, codes, other);
    return n_other;
}

using PackFn = uint64_t (*)(const uint8_t*, uint64_t, uint64_t, uint32_t*, uint32_t*);

PackFn pick_body(int which) {
    __builtin_cpu_init();
    const bool has512 = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
    const bool has2 = __builtin_cpu_supports("avx2");
    if (which == 3) return has512 ? pack_avx512 : nullptr;
    if (which == 2) return has2 ? pack_avx2 : nullptr;
    if (which == 1) return pack_scalar;
    return has512 ? pack_avx512 : (has2 ? pack_avx2 : pack_scalar);
}

}  // namespace

int host_pack_body_available(int which) { return pick_body(which) != nullptr; }

uint64_t host_pack_range(const uint8_t* bases, uint64_t p0, uint64_t p1, uint32_t* codes, uint32_t* other, int body) {
    static const PackFn best = pick_body(0);
    PackFn fn = body == 0 ? best : pick_body(body);
    if (!fn) fn = pack_scalar;
    return fn(bases, p0, p1, codes, other);
}

uint64_t host_pack_bases(const uint8_t* bases, uint64_t n, uint32_t* codes, uint32_t* other, int threads, int body) {
    if (n == 0) return 0;
    constexpr uint64_t kGrain = 1ull << 16;  // multiple of 64: every range starts on a word of both arrays
    uint64_t want = threads <= 0 ? std::max(1u, std::thread::hardware_concurrency()) : (uint64_t)threads;
    want = std::min<uint64_t>(want, (n + kGrain - 1) / kGrain);
    if (want <= 1) return host_pack_range(bases, 0, n, codes, other, body);
    const uint64_t per = ((n + want - 1) / want + kGrain - 1) / kGrain * kGrain;
    std::vector<uint64_t> part(want, 0);
    std::vector<std::thread> pool;
    pool.reserve(want - 1);
    auto run = [&](uint64_t t) {
        const uint64_t a = std::min(n, t * per), b = std::min(n, a + per);
        if (a < b) part[t] = host_pack_range(bases, a, b, codes, other, body);
    };
    for (uint64_t t = 1; t < want; ++t) pool.emplace_back(run, t);
    run(0);
    uint64_t total = 0;
    for (auto& th : pool) th.join();
    for (uint64_t v : part) total += v;
    return total;
}

}  // namespace nk
