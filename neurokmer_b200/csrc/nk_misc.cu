// nk_misc.cu — parity tap (hash of a word array) and the synthetic base generator.
#include "nk_kernels.cuh"

namespace nk {

namespace {

// SipHash-1-3{0,0}(LE64(word)) and % pool_size for an array of words: the values
// `hasher.finish()` and `idx` take at src/spiking_hash.rs:79-81.
template <bool POW2>
__global__ void hash_words_kernel(const unsigned long long* __restrict__ words, unsigned long long n, FastMod fm,
                                  RotMul rm, unsigned long long* hashes, unsigned long long* idx) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long w = words[i];
    const U64 h = siphash13_dev((unsigned)w, (unsigned)(w >> 32), rm);
    if (hashes) hashes[i] = ((unsigned long long)h.hi << 32) | h.lo;
    if (idx) idx[i] = fastmod_dev<POW2>(h, fm);
}

// h % pool for an array of raw 64-bit values (parity tap for the exact-modulo routine alone)
template <bool POW2>
__global__ void mod_words_kernel(const unsigned long long* __restrict__ h, unsigned long long n, FastMod fm,
                                 unsigned long long* out, int which) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const U64 x{(unsigned)h[i], (unsigned)(h[i] >> 32)};
    if (which == 2 && !POW2) {  // the routine the count kernel picks for this pool size (fm.kind)
        out[i] = fm.kind == 3u ? fastmod_single_dev<3>(x, fm) : (fm.kind == 2u ? fastmod_single_dev<2>(x, fm) : fastmod_dev<false>(x, fm));
        return;
    }
    out[i] = which == 1 ? fastmod_mg_dev<POW2>(x, fm) : fastmod_dev<POW2>(x, fm);
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

// ---- peak calibration micro-kernels (roofline denominators, measured live) -----------------
// MODE 0: independent LOP3 + SHF chains (the ALU-pipe-only part of SipHash: xor, rotate)
// MODE 1: independent SipRound chains (add : xor : rotate = 1 : 1 : 1 on 64-bit words,
//         24 32-bit ops per round) — the integer-pipe ceiling for SipHash's own mix
template <int MODE>
__global__ void __launch_bounds__(256) int_peak_kernel(unsigned int* out, unsigned iters, unsigned seed, RotMul rm) {
    constexpr int PLANROW = 2;  // a middle round of the production plan (same pipe mix as the count kernel)
    constexpr int CH = 4;
    if (MODE == 0) {
        unsigned a[2 * CH];
#pragma unroll
        for (int i = 0; i < 2 * CH; ++i) a[i] = seed + threadIdx.x * 31u + i * 0x9E3779B9u + blockIdx.x;
        const unsigned b = seed ^ 0x5bd1e995u;
        for (unsigned it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
#pragma unroll
                for (int i = 0; i < 2 * CH; ++i) {
                    a[i] = __funnelshift_l(a[i], a[(i + 1) % (2 * CH)], 13);  // SHF
                    a[i] ^= b;                                                  // LOP3
                }
            }
        }
        unsigned x = 0;
#pragma unroll
        for (int i = 0; i < 2 * CH; ++i) x ^= a[i];
        if (x == 0x12345u) out[0] = x;
    } else {
        U64 v0[CH], v1[CH], v2[CH], v3[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            v0[i] = U64{seed + i, threadIdx.x};
            v1[i] = U64{seed * 3u + i, blockIdx.x};
            v2[i] = U64{seed ^ 0xabcdefu, i * 77u + threadIdx.x};
            v3[i] = U64{threadIdx.x * 2654435761u, seed + 5u * i};
        }
        for (unsigned it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
#pragma unroll
                for (int i = 0; i < CH; ++i) NK_SIPROUND32(v0[i], v1[i], v2[i], v3[i], PLANROW, rm);
            }
        }
        unsigned x = 0;
#pragma unroll
        for (int i = 0; i < CH; ++i) x ^= v0[i].lo ^ v1[i].hi ^ v2[i].lo ^ v3[i].hi;
        if (x == 0x12345u) out[0] = x;
    }
}

// The pool-update limit: RED.ADD.U32 at the addresses real k-mer traffic produces — SipHash-1-3 of consecutive
// words modulo the pool, precomputed into `idx` so that the timed kernel does nothing but stream 4-byte indices
// (evict-first loads) and fire reductions.  (Round 1 generated addresses with an LCG inside the kernel; at
// 64 K neurons the count kernel itself beat that "ceiling", i.e. the address stream, not the L2, was the limit.)
template <bool POW2>
__global__ void hashed_idx_kernel(unsigned int* __restrict__ idx, unsigned long long n, unsigned long long seed, FastMod fm, RotMul rm) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long w = seed + i;
        const U64 h = siphash13_dev((unsigned)w, (unsigned)(w >> 32), rm);
        idx[i] = fastmod_dev<POW2>(h, fm);
    }
}

__global__ void __launch_bounds__(256) red_peak_kernel(unsigned int* acc, const unsigned int* __restrict__ idx, unsigned long long n) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        const unsigned a = __ldcs(idx + i), b = __ldcs(idx + i + stride), c = __ldcs(idx + i + 2 * stride),
                       d = __ldcs(idx + i + 3 * stride);
        atomicAdd(acc + a, 1u);
        atomicAdd(acc + b, 1u);
        atomicAdd(acc + c, 1u);
        atomicAdd(acc + d, 1u);
    }
    for (; i < n; i += stride) atomicAdd(acc + __ldcs(idx + i), 1u);
}

}  // namespace

cudaError_t launch_int_peak(int mode, unsigned int* out, int blocks, unsigned iters, cudaStream_t s) {
    if (mode == 0) int_peak_kernel<0><<<blocks, 256, 0, s>>>(out, iters, 17u, make_rotmul());
    else int_peak_kernel<1><<<blocks, 256, 0, s>>>(out, iters, 17u, make_rotmul());
    return cudaGetLastError();
}
// 32-bit integer ops one thread executes per `iters` unit in launch_int_peak
unsigned long long int_peak_ops_per_iter(int mode) { return mode == 0 ? 8ull * 8 * 2 : 2ull * 4 * 24; }

cudaError_t launch_hashed_idx(unsigned int* idx, unsigned long long n, unsigned long long seed, FastMod fm, cudaStream_t s) {
    if (fm.is_pow2) hashed_idx_kernel<true><<<148 * 8, 256, 0, s>>>(idx, n, seed, fm, make_rotmul());
    else hashed_idx_kernel<false><<<148 * 8, 256, 0, s>>>(idx, n, seed, fm, make_rotmul());
    return cudaGetLastError();
}

cudaError_t launch_red_peak(unsigned int* acc, const unsigned int* idx, unsigned long long n, int blocks, cudaStream_t s) {
    red_peak_kernel<<<blocks, 256, 0, s>>>(acc, idx, n);
    return cudaGetLastError();
}

cudaError_t launch_hash_words(const unsigned long long* words, unsigned long long n, FastMod fm,
                              unsigned long long* hashes, unsigned long long* idx, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (fm.is_pow2)
        hash_words_kernel<true><<<blocks, 256, 0, s>>>(words, n, fm, make_rotmul(), hashes, idx);
    else
        hash_words_kernel<false><<<blocks, 256, 0, s>>>(words, n, fm, make_rotmul(), hashes, idx);
    return cudaGetLastError();
}

cudaError_t launch_mod_words(const unsigned long long* h, unsigned long long n, FastMod fm, unsigned long long* out,
                             int which, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (fm.is_pow2) mod_words_kernel<true><<<blocks, 256, 0, s>>>(h, n, fm, out, which);
    else mod_words_kernel<false><<<blocks, 256, 0, s>>>(h, n, fm, out, which);
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
