// nk_misc.cu — parity tap (hash of a word array) and the synthetic base generator.
#include "nk_kernels.cuh"

namespace nk {

namespace {

// SipHash-1-3{0,0}(LE64(word)) and % pool_size for an array of words: the values
// `hasher.finish()` and `idx` take at src/spiking_hash.rs:79-81.
template <bool POW2>
__global__ void hash_words_kernel(const unsigned long long* __restrict__ words, unsigned long long n, FastMod fm,
                                  unsigned long long* hashes, unsigned long long* idx) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long w = words[i];
    const U64 h = siphash13_dev((unsigned)w, (unsigned)(w >> 32));
    if (hashes) hashes[i] = ((unsigned long long)h.hi << 32) | h.lo;
    if (idx) idx[i] = fastmod_dev<POW2>(h, fm);
}

// Position-addressable synthetic stream: 32 bases per splitmix64 draw.
//   flags bit0: one run of 100..10000 'N' per 2^20-base block (≈0.5 % of bases)
//   flags bit1: 1 % of the 4096-base blocks are lower-case (soft-masked)
constexpr unsigned long long kGold = 0x9E3779B97F4A7C15ULL;

__device__ __forceinline__ bool in_n_run(unsigned long long seed, unsigned long long p) {
    const unsigned long long blk = p >> 20;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
        if (d == 1 && blk == 0) break;
        const unsigned long long b = blk - d;
        const unsigned long long h = splitmix64((seed ^ 0x4E52554EULL) * kGold + b);
        const unsigned long long start = (b << 20) + (h & 0xFFFFFULL);
        const unsigned long long len = 100ULL + ((h >> 20) % 9901ULL);
        if (p >= start && p < start + len) return true;
    }
    return false;
}

__global__ void synth_kernel(unsigned char* out, unsigned long long seed, unsigned long long start,
                             unsigned long long n, unsigned flags) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long p = start + i;
        const unsigned long long h = splitmix64(seed * kGold + (p >> 5));
        unsigned char c = "ACGT"[(h >> (2 * (p & 31))) & 3];
        if ((flags & 2u) && (splitmix64((seed ^ 0x6C6F7765ULL) * kGold + (p >> 12)) % 100ULL) == 0) c |= 0x20;
        if ((flags & 1u) && in_n_run(seed, p)) c = 'N';
        out[i] = c;
    }
}

}  // namespace

cudaError_t launch_hash_words(const unsigned long long* words, unsigned long long n, FastMod fm,
                              unsigned long long* hashes, unsigned long long* idx, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (fm.is_pow2)
        hash_words_kernel<true><<<blocks, 256, 0, s>>>(words, n, fm, hashes, idx);
    else
        hash_words_kernel<false><<<blocks, 256, 0, s>>>(words, n, fm, hashes, idx);
    return cudaGetLastError();
}

cudaError_t launch_synth(unsigned char* out, unsigned long long seed, unsigned long long start,
                         unsigned long long n, unsigned flags, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    unsigned long long blocks = (n + 255) / 256;
    if (blocks > 148ull * 32) blocks = 148ull * 32;
    synth_kernel<<<(unsigned)blocks, 256, 0, s>>>(out, seed, start, n, flags);
    return cudaGetLastError();
}

}  // namespace nk
