// nk_topn.cu — top-N neurons by spike count (sm_100a).
//
// Replaces top_abundant_neurons (src/spiking_hash.rs:661-673): the reference
// materialises pool_size tuples and stable-sorts them descending by spike_count,
// so ties keep ascending neuron index.  Here:
//   1. MSB-first 8-bit radix SELECT on the u64 spike counts finds the exact
//      threshold value T of the N-th largest count, the number of counts > T and
//      how many counts == T are still needed (block histograms in shared memory);
//   2. a three-kernel ordered gather writes every neuron with count > T and the
//      `need` LOWEST-INDEX neurons with count == T (block counts -> scan -> scatter,
//      index order preserved inside a block by a thread-ordered scan);
//   3. a bitonic sort of the <= N candidates orders them (spikes desc, idx asc).
// Only the digits below the top set bit of max_spikes are visited (2 passes for the
// reference's 334-spike ceiling).
#include "nk_kernels.cuh"

namespace nk {

namespace {

constexpr int TN_THREADS = 256;
constexpr int TN_ITEMS = TOPN_BLOCK_ITEMS / TN_THREADS;  // 16 consecutive neurons per thread
// ctrl layout
enum { C_PREFIX = 0, C_RANK = 1, C_CURSOR = 2, C_GT = 3, C_NEED = 4 };

__global__ void topn_init_kernel(unsigned int* hist, unsigned long long* ctrl, unsigned long long n) {
    hist[threadIdx.x] = 0;
    if (threadIdx.x == 0) {
        ctrl[C_PREFIX] = 0;
        ctrl[C_RANK] = n;
        ctrl[C_CURSOR] = 0;
        ctrl[C_GT] = 0;
        ctrl[C_NEED] = 0;
    }
}

// histogram of digit `d` over the elements whose higher digits equal the prefix
__global__ void __launch_bounds__(TN_THREADS) topn_hist_kernel(const unsigned long long* __restrict__ spikes,
                                                               unsigned long long pool, int d, int top,
                                                               unsigned int* hist,
                                                               const unsigned long long* __restrict__ ctrl) {
    __shared__ unsigned int sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long prefix = ctrl[C_PREFIX];
    const int hs = 8 * (d + 1);
    const unsigned long long stride = (unsigned long long)gridDim.x * TN_THREADS;
    for (unsigned long long i = blockIdx.x * (unsigned long long)TN_THREADS + threadIdx.x; i < pool; i += stride) {
        const unsigned long long v = spikes[i];
        const unsigned long long high = (d == top) ? 0ull : (v >> hs);
        if (high == prefix) atomicAdd(&sh[(v >> (8 * d)) & 255u], 1u);
    }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// pick the bin that contains the rank-th largest element; one block of 256 threads
__global__ void topn_pick_kernel(unsigned int* hist, unsigned long long* ctrl, int last) {
    __shared__ unsigned long long above[256];  // elements in bins > b
    const unsigned b = threadIdx.x;
    __shared__ unsigned int h[256];
    h[b] = hist[b];
    __syncthreads();
    if (b == 0) {
        unsigned long long run = 0;
        for (int x = 255; x >= 0; --x) { above[x] = run; run += h[x]; }
    }
    __syncthreads();
    const unsigned long long rank = ctrl[C_RANK];
    const bool mine = above[b] < rank && rank <= above[b] + h[b];
    __syncthreads();
    if (mine) {
        ctrl[C_PREFIX] = (ctrl[C_PREFIX] << 8) | b;
        ctrl[C_RANK] = rank - above[b];
        ctrl[C_GT] += above[b];
        if (last) ctrl[C_NEED] = rank - above[b];
    }
    hist[b] = 0;
}

__global__ void __launch_bounds__(TN_THREADS) topn_count_eq_kernel(const unsigned long long* __restrict__ spikes,
                                                                   unsigned long long pool,
                                                                   const unsigned long long* __restrict__ ctrl,
                                                                   unsigned int* block_counts) {
    const unsigned long long T = ctrl[C_PREFIX];
    const unsigned long long base = blockIdx.x * (unsigned long long)TOPN_BLOCK_ITEMS;
    unsigned c = 0;
    for (int it = 0; it < TN_ITEMS; ++it) {  // coalesced; order is irrelevant for a count
        const unsigned long long i = base + it * TN_THREADS + threadIdx.x;
        if (i < pool && spikes[i] == T) ++c;
    }
    c = __reduce_add_sync(0xFFFFFFFFu, c);
    __shared__ unsigned int s[TN_THREADS / 32];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < TN_THREADS / 32; ++w) t += s[w];
        block_counts[blockIdx.x] = t;
    }
}

// in-place exclusive scan of block_counts (saturating at 2^32-1 is impossible: pool < 2^32)
__global__ void __launch_bounds__(1024) topn_scan_kernel(unsigned int* block_counts, unsigned long long nblocks) {
    __shared__ unsigned int s[1024];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (unsigned long long base = 0; base < nblocks; base += 1024) {
        const unsigned long long i = base + threadIdx.x;
        const unsigned v = i < nblocks ? block_counts[i] : 0u;
        s[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            const unsigned a = threadIdx.x >= (unsigned)o ? s[threadIdx.x - o] : 0u;
            __syncthreads();
            s[threadIdx.x] += a;
            __syncthreads();
        }
        const unsigned incl = s[threadIdx.x];
        if (i < nblocks) block_counts[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(TN_THREADS) topn_gather_kernel(const unsigned long long* __restrict__ spikes,
                                                                 unsigned long long pool, unsigned long long* ctrl,
                                                                 const unsigned int* __restrict__ block_counts,
                                                                 unsigned long long* out_idx,
                                                                 unsigned long long* out_spikes) {
    const unsigned long long T = ctrl[C_PREFIX], gt = ctrl[C_GT], need = ctrl[C_NEED];
    // thread t owns the TN_ITEMS consecutive neurons base + t*TN_ITEMS ... (index order)
    const unsigned long long base = blockIdx.x * (unsigned long long)TOPN_BLOCK_ITEMS + threadIdx.x * TN_ITEMS;
    unsigned long long v[TN_ITEMS];
    unsigned eq = 0;
#pragma unroll
    for (int it = 0; it < TN_ITEMS; ++it) {
        const unsigned long long i = base + it;
        v[it] = i < pool ? spikes[i] : 0ull;
        if (i < pool && v[it] == T) ++eq;
        if (i < pool && v[it] > T) {
            const unsigned long long slot = atomicAdd(&ctrl[C_CURSOR], 1ull);
            out_idx[slot] = i;
            out_spikes[slot] = v[it];
        }
    }
    // exclusive scan of eq over the block's threads
    __shared__ unsigned int s[TN_THREADS];
    s[threadIdx.x] = eq;
    __syncthreads();
    for (int o = 1; o < TN_THREADS; o <<= 1) {
        const unsigned a = threadIdx.x >= (unsigned)o ? s[threadIdx.x - o] : 0u;
        __syncthreads();
        s[threadIdx.x] += a;
        __syncthreads();
    }
    unsigned long long rank = (unsigned long long)block_counts[blockIdx.x] + (s[threadIdx.x] - eq);
    if (eq == 0 || rank >= need) return;
#pragma unroll
    for (int it = 0; it < TN_ITEMS; ++it) {
        const unsigned long long i = base + it;
        if (i < pool && v[it] == T) {
            if (rank < need) {
                out_idx[gt + rank] = i;
                out_spikes[gt + rank] = T;
            }
            ++rank;
        }
    }
}

__device__ __forceinline__ bool before(unsigned long long sa, unsigned long long ia, unsigned long long sb,
                                       unsigned long long ib) {
    return sa > sb || (sa == sb && ia < ib);
}

// single-block bitonic sort of n <= 2048 candidates (padded with (0, ~0) sentinels)
__global__ void __launch_bounds__(1024) topn_sort_small_kernel(unsigned long long* idx, unsigned long long* spk,
                                                               unsigned n) {
    __shared__ unsigned long long si[2048], ss[2048];
    unsigned N = 1;
    while (N < n) N <<= 1;
    for (unsigned i = threadIdx.x; i < N; i += 1024) {
        si[i] = i < n ? idx[i] : ~0ull;
        ss[i] = i < n ? spk[i] : 0ull;
    }
    __syncthreads();
    for (unsigned k = 2; k <= N; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = threadIdx.x; i < N; i += 1024) {
                const unsigned l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const bool swap = up ? before(ss[l], si[l], ss[i], si[i]) : before(ss[i], si[i], ss[l], si[l]);
                    if (swap) {
                        const unsigned long long a = si[i], b = ss[i];
                        si[i] = si[l]; ss[i] = ss[l];
                        si[l] = a; ss[l] = b;
                    }
                }
            }
            __syncthreads();
        }
    }
    for (unsigned i = threadIdx.x; i < n; i += 1024) {
        idx[i] = si[i];
        spk[i] = ss[i];
    }
}

__global__ void topn_pad_kernel(unsigned long long* idx, unsigned long long* spk, unsigned long long n,
                                unsigned long long N) {
    const unsigned long long i = n + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < N) { idx[i] = ~0ull; spk[i] = 0ull; }
}

__global__ void topn_bitonic_step_kernel(unsigned long long* idx, unsigned long long* spk, unsigned long long N,
                                         unsigned long long k, unsigned long long j) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= N) return;
    const unsigned long long l = i ^ j;
    if (l > i) {
        const bool up = (i & k) == 0;
        const unsigned long long si = spk[i], ii = idx[i], sl = spk[l], il = idx[l];
        const bool swap = up ? before(sl, il, si, ii) : before(si, ii, sl, il);
        if (swap) { idx[i] = il; spk[i] = sl; idx[l] = ii; spk[l] = si; }
    }
}

}  // namespace

cudaError_t launch_topn(const unsigned long long* spikes, unsigned long long pool, unsigned long long n,
                        unsigned long long max_spikes, const TopNScratch& sc, cudaStream_t s,
                        uint64_t* launches) {
    if (n > pool) n = pool;
    if (n == 0) return cudaSuccess;
    uint64_t L = 0;
    int bits = 0;
    while (bits < 64 && (max_spikes >> bits)) ++bits;
    int passes = (bits + 7) / 8;
    if (passes < 1) passes = 1;
    const int top = passes - 1;

    topn_init_kernel<<<1, 256, 0, s>>>(sc.hist, sc.ctrl, n); ++L;
    unsigned long long hb = (pool + TN_THREADS * 8 - 1) / (TN_THREADS * 8);
    if (hb > 148ull * 8) hb = 148ull * 8;
    if (hb == 0) hb = 1;
    for (int d = top; d >= 0; --d) {
        topn_hist_kernel<<<(unsigned)hb, TN_THREADS, 0, s>>>(spikes, pool, d, top, sc.hist, sc.ctrl); ++L;
        topn_pick_kernel<<<1, 256, 0, s>>>(sc.hist, sc.ctrl, d == 0 ? 1 : 0); ++L;
    }
    const unsigned long long nblocks = (pool + TOPN_BLOCK_ITEMS - 1) / TOPN_BLOCK_ITEMS;
    topn_count_eq_kernel<<<(unsigned)nblocks, TN_THREADS, 0, s>>>(spikes, pool, sc.ctrl, sc.block_counts); ++L;
    topn_scan_kernel<<<1, 1024, 0, s>>>(sc.block_counts, nblocks); ++L;
    topn_gather_kernel<<<(unsigned)nblocks, TN_THREADS, 0, s>>>(spikes, pool, sc.ctrl, sc.block_counts, sc.out_idx,
                                                                 sc.out_spikes); ++L;
    if (n <= 2048) {
        topn_sort_small_kernel<<<1, 1024, 0, s>>>(sc.out_idx, sc.out_spikes, (unsigned)n); ++L;
    } else {
        unsigned long long N = 1;
        while (N < n) N <<= 1;
        if (N > n) { topn_pad_kernel<<<(unsigned)((N - n + 255) / 256), 256, 0, s>>>(sc.out_idx, sc.out_spikes, n, N); ++L; }
        for (unsigned long long k = 2; k <= N; k <<= 1)
            for (unsigned long long j = k >> 1; j > 0; j >>= 1) {
                topn_bitonic_step_kernel<<<(unsigned)((N + 255) / 256), 256, 0, s>>>(sc.out_idx, sc.out_spikes, N, k, j);
                ++L;
            }
    }
    if (launches) *launches += L;
    return cudaGetLastError();
}

}  // namespace nk
