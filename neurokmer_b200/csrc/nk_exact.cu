// nk_exact.cu — the exact side tables of the reference counter, opt-in.
//
// Replaces, when nk_enable_exact_counts(h, 1) was called:
//   counts: DashMap<u64, AtomicU32>             src/spiking_hash.rs:27, filled :110,124,157-165,333,343,442-447
//   get_count                                   src/spiking_hash.rs:675-678
//   kmer_per_neuron: DashMap<usize, u32>        :26, rebuilt :167-172 / :467-473 (= #distinct k-mers per neuron,
//                                               the `uniques` column of top_abundant_neurons :667)
//   process_sequence's variants                 :218-221, 259-263 (# sequences that touched the neuron)
//
// GPU form: the count kernel appends every window's word to a device array (warp-aggregated
// cursor); at the end of the call the array is radix-sorted and run-length encoded into a sorted
// (key, count) table; get_count is a binary search; kmer_per_neuron is a histogram of
// hash(key) % pool over the distinct keys.  The sort and the run-length encode are CUB library
// calls (cub::DeviceRadixSort / DeviceRunLengthEncode / DeviceReduce::ReduceByKey) — this row is
// outside the measured hot path; everything on the hot path is hand-written.
// Memory is O(windows of the call): 16 B per window while sorting.  Counts wrap at 2^32 like the
// reference's AtomicU32::fetch_add.
#include <cub/cub.cuh>

#include "nk_kernels.cuh"

namespace nk {

namespace {

template <bool POW2>
__global__ void uniques_hist_kernel(const unsigned long long* __restrict__ keys, unsigned long long n, FastMod fm,
                                    RotMul rm, unsigned int* uniques) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long w = keys[i];
    const U64 h = siphash13_dev((unsigned)w, (unsigned)(w >> 32), rm);
    atomicAdd(uniques + fastmod_dev<POW2>(h, fm), 1u);
}

// process_sequence: every neuron touched by this sequence gets +1 (local_unique, :204,226,262-264)
template <bool POW2>
__global__ void touched_kernel(const unsigned long long* __restrict__ keys, unsigned long long n, FastMod fm, RotMul rm,
                               unsigned int* flags, unsigned int* uniques, int phase) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long w = keys[i];
    const U64 h = siphash13_dev((unsigned)w, (unsigned)(w >> 32), rm);
    const unsigned idx = fastmod_dev<POW2>(h, fm);
    if (phase == 0) {
        if (atomicExch(flags + idx, 1u) == 0u) atomicAdd(uniques + idx, 1u);
    } else {
        flags[idx] = 0u;
    }
}

__global__ void lookup_kernel(const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ counts,
                              unsigned long long n, unsigned long long key, unsigned long long* out) {
    unsigned long long lo = 0, hi = n;
    while (lo < hi) {
        const unsigned long long mid = lo + (hi - lo) / 2;
        if (keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    const bool found = lo < n && keys[lo] == key;
    out[0] = found ? 1ull : 0ull;
    out[1] = found ? counts[lo] : 0ull;
}

__global__ void gather_uniques_kernel(const unsigned long long* __restrict__ idx, unsigned long long n,
                                      const unsigned int* __restrict__ uniques, unsigned int* out) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = uniques[idx[i]];
}

__global__ void int_to_u32_kernel(const int* __restrict__ in, unsigned int* out, unsigned long long n) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (unsigned int)in[i];
}

#define NKX(expr)                        \
    do {                                 \
        cudaError_t e_ = (expr);         \
        if (e_ != cudaSuccess) return e_; \
    } while (0)

cudaError_t ensure(void** p, unsigned long long* cap, unsigned long long bytes) {
    if (bytes <= *cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    NKX(cudaMalloc(p, bytes));
    *cap = bytes;
    return cudaSuccess;
}

}  // namespace

cudaError_t exact_reserve_words(ExactTable& t, unsigned long long extra, cudaStream_t s) {
    if (!t.cursor) {
        NKX(cudaMalloc(&t.cursor, 4 * sizeof(unsigned long long)));
        NKX(cudaMemsetAsync(t.cursor, 0, 4 * sizeof(unsigned long long), s));
    }
    const unsigned long long need = t.words_bound + extra;
    if (need > t.words_cap) {
        unsigned long long cap = need + need / 2 + 1024;
        unsigned long long* nw = nullptr;
        NKX(cudaMalloc(&nw, cap * sizeof(unsigned long long)));
        if (t.words) {
            NKX(cudaMemcpyAsync(nw, t.words, t.words_bound * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
            NKX(cudaStreamSynchronize(s));
            cudaFree(t.words);
        }
        t.words = nw;
        t.words_cap = cap;
    }
    t.words_bound = need;
    return cudaSuccess;
}

cudaError_t exact_grow_words(ExactTable& t, unsigned long long cap, unsigned long long keep, cudaStream_t s) {
    if (!t.cursor) {
        NKX(cudaMalloc(&t.cursor, 4 * sizeof(unsigned long long)));
        NKX(cudaMemsetAsync(t.cursor, 0, 4 * sizeof(unsigned long long), s));
    }
    if (cap <= t.words_cap) return cudaSuccess;
    unsigned long long* nw = nullptr;
    NKX(cudaMalloc(&nw, cap * sizeof(unsigned long long)));
    if (t.words) {
        if (keep) NKX(cudaMemcpyAsync(nw, t.words, keep * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
        NKX(cudaStreamSynchronize(s));
        cudaFree(t.words);
    }
    t.words = nw;
    t.words_cap = cap;
    return cudaSuccess;
}

namespace {
__global__ void filter_set_kernel(unsigned int* filter, const unsigned long long* __restrict__ idx, unsigned long long n) {
    const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (i < n) atomicOr(filter + (idx[i] >> 5), 1u << (idx[i] & 31u));
}
}  // namespace

cudaError_t launch_filter_set(unsigned int* filter, const unsigned long long* idx, unsigned long long n, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    filter_set_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(filter, idx, n);
    return cudaGetLastError();
}

cudaError_t exact_clear(ExactTable& t, unsigned long long pool, bool tables_too, cudaStream_t s) {
    t.words_bound = 0;
    if (t.cursor) NKX(cudaMemsetAsync(t.cursor, 0, sizeof(unsigned long long), s));
    if (tables_too) {
        t.n_keys = 0;
        t.valid = false;
        if (t.uniques) NKX(cudaMemsetAsync(t.uniques, 0, pool * sizeof(unsigned int), s));
    }
    return cudaSuccess;
}

// words[0..n) -> sorted (keys, counts); merge: add to the existing table instead of replacing it,
// and use process_sequence's "neurons touched by this sequence" rule for the per-neuron column.
cudaError_t exact_finalize(ExactTable& t, const FastMod& fm, unsigned long long pool, unsigned key_bits, bool merge,
                           cudaStream_t s) {
    unsigned long long n = 0;
    if (t.cursor) {
        NKX(cudaMemcpyAsync(&n, t.cursor, sizeof n, cudaMemcpyDeviceToHost, s));
        NKX(cudaStreamSynchronize(s));
    }
    if (n > 0x7FFFFFF0ull) return cudaErrorInvalidValue;  // CUB run-length encode takes int num_items
    if (!t.uniques) {
        NKX(cudaMalloc(&t.uniques, pool * sizeof(unsigned int)));
        NKX(cudaMemsetAsync(t.uniques, 0, pool * sizeof(unsigned int), s));
    }
    if (!merge) {
        t.n_keys = 0;
        NKX(cudaMemsetAsync(t.uniques, 0, pool * sizeof(unsigned int), s));
    }
    const RotMul rm = make_rotmul();
    unsigned long long n_new = 0;
    if (n > 0) {
        // 1. sort the words
        NKX(ensure((void**)&t.alt, &t.alt_cap, n * sizeof(unsigned long long)));
        cub::DoubleBuffer<unsigned long long> db(t.words, t.alt);
        size_t tmp = 0;
        NKX(cub::DeviceRadixSort::SortKeys(nullptr, tmp, db, (long long)n, 0, (int)key_bits, s));
        NKX(ensure(&t.tmp, &t.tmp_cap, tmp));
        NKX(cub::DeviceRadixSort::SortKeys(t.tmp, tmp, db, (long long)n, 0, (int)key_bits, s));
        const unsigned long long* sorted = db.Current();
        // 2. run-length encode into the spare buffers
        NKX(ensure((void**)&t.rk, &t.rk_cap, n * sizeof(unsigned long long)));
        NKX(ensure((void**)&t.rc, &t.rc_cap, n * sizeof(int)));
        size_t tmp2 = 0;
        unsigned long long* d_runs = t.cursor + 1;
        NKX(cub::DeviceRunLengthEncode::Encode(nullptr, tmp2, sorted, t.rk, (int*)t.rc, (int*)d_runs, (int)n, s));
        NKX(ensure(&t.tmp, &t.tmp_cap, tmp2));
        NKX(cudaMemsetAsync(d_runs, 0, sizeof(unsigned long long), s));
        NKX(cub::DeviceRunLengthEncode::Encode(t.tmp, tmp2, sorted, t.rk, (int*)t.rc, (int*)d_runs, (int)n, s));
        unsigned long long runs = 0;
        NKX(cudaMemcpyAsync(&runs, d_runs, sizeof runs, cudaMemcpyDeviceToHost, s));
        NKX(cudaStreamSynchronize(s));
        n_new = runs & 0xFFFFFFFFull;
        const unsigned blocks = (unsigned)((n_new + 255) / 256);
        // 3. per-neuron column
        if (!merge) {
            if (fm.is_pow2) uniques_hist_kernel<true><<<blocks, 256, 0, s>>>(t.rk, n_new, fm, rm, t.uniques);
            else uniques_hist_kernel<false><<<blocks, 256, 0, s>>>(t.rk, n_new, fm, rm, t.uniques);
        } else {
            if (!t.flags) {
                NKX(cudaMalloc(&t.flags, pool * sizeof(unsigned int)));
                NKX(cudaMemsetAsync(t.flags, 0, pool * sizeof(unsigned int), s));
            }
            for (int phase = 0; phase < 2; ++phase) {
                if (fm.is_pow2) touched_kernel<true><<<blocks, 256, 0, s>>>(t.rk, n_new, fm, rm, t.flags, t.uniques, phase);
                else touched_kernel<false><<<blocks, 256, 0, s>>>(t.rk, n_new, fm, rm, t.flags, t.uniques, phase);
            }
        }
        NKX(cudaGetLastError());
    }
    // 4. install / merge the (key, count) table
    if (!merge || t.n_keys == 0) {
        NKX(ensure((void**)&t.keys, &t.keys_cap, (n_new ? n_new : 1) * sizeof(unsigned long long)));
        NKX(ensure((void**)&t.counts, &t.counts_cap, (n_new ? n_new : 1) * sizeof(unsigned int)));
        if (n_new) {
            NKX(cudaMemcpyAsync(t.keys, t.rk, n_new * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
            int_to_u32_kernel<<<(unsigned)((n_new + 255) / 256), 256, 0, s>>>((const int*)t.rc, t.counts, n_new);
        }
        t.n_keys = n_new;
    } else if (n_new > 0) {
        // concatenate old and new pairs, sort by key, reduce equal keys (counts wrap like AtomicU32)
        const unsigned long long m = t.n_keys + n_new;
        unsigned long long *ck = nullptr, *ck2 = nullptr;
        unsigned int *cc = nullptr, *cc2 = nullptr;
        NKX(cudaMalloc(&ck, m * 8)); NKX(cudaMalloc(&ck2, m * 8)); NKX(cudaMalloc(&cc, m * 4)); NKX(cudaMalloc(&cc2, m * 4));
        NKX(cudaMemcpyAsync(ck, t.keys, t.n_keys * 8, cudaMemcpyDeviceToDevice, s));
        NKX(cudaMemcpyAsync(ck + t.n_keys, t.rk, n_new * 8, cudaMemcpyDeviceToDevice, s));
        NKX(cudaMemcpyAsync(cc, t.counts, t.n_keys * 4, cudaMemcpyDeviceToDevice, s));
        int_to_u32_kernel<<<(unsigned)((n_new + 255) / 256), 256, 0, s>>>((const int*)t.rc, cc + t.n_keys, n_new);
        size_t tmp3 = 0;
        NKX(cub::DeviceRadixSort::SortPairs(nullptr, tmp3, ck, ck2, cc, cc2, (long long)m, 0, (int)key_bits, s));
        NKX(ensure(&t.tmp, &t.tmp_cap, tmp3));
        NKX(cub::DeviceRadixSort::SortPairs(t.tmp, tmp3, ck, ck2, cc, cc2, (long long)m, 0, (int)key_bits, s));
        unsigned long long* d_runs = t.cursor + 1;
        size_t tmp4 = 0;
        NKX(cub::DeviceReduce::ReduceByKey(nullptr, tmp4, ck2, ck, cc2, cc, (int*)d_runs, cub::Sum(), (int)m, s));
        NKX(ensure(&t.tmp, &t.tmp_cap, tmp4));
        NKX(cudaMemsetAsync(d_runs, 0, sizeof(unsigned long long), s));
        NKX(cub::DeviceReduce::ReduceByKey(t.tmp, tmp4, ck2, ck, cc2, cc, (int*)d_runs, cub::Sum(), (int)m, s));
        unsigned long long runs = 0;
        NKX(cudaMemcpyAsync(&runs, d_runs, sizeof runs, cudaMemcpyDeviceToHost, s));
        NKX(cudaStreamSynchronize(s));
        runs &= 0xFFFFFFFFull;
        NKX(ensure((void**)&t.keys, &t.keys_cap, runs * 8));
        NKX(ensure((void**)&t.counts, &t.counts_cap, runs * 4));
        NKX(cudaMemcpyAsync(t.keys, ck, runs * 8, cudaMemcpyDeviceToDevice, s));
        NKX(cudaMemcpyAsync(t.counts, cc, runs * 4, cudaMemcpyDeviceToDevice, s));
        NKX(cudaStreamSynchronize(s));
        cudaFree(ck); cudaFree(ck2); cudaFree(cc); cudaFree(cc2);
        t.n_keys = runs;
    }
    t.valid = true;
    // the words of this call are consumed
    t.words_bound = 0;
    if (t.cursor) NKX(cudaMemsetAsync(t.cursor, 0, sizeof(unsigned long long), s));
    return cudaGetLastError();
}

cudaError_t exact_lookup(const ExactTable& t, unsigned long long key, unsigned long long* d_out2, cudaStream_t s) {
    lookup_kernel<<<1, 1, 0, s>>>(t.keys, t.counts, t.n_keys, key, d_out2);
    return cudaGetLastError();
}

cudaError_t exact_gather_uniques(const ExactTable& t, const unsigned long long* idx, unsigned long long n,
                                 unsigned int* out, cudaStream_t s) {
    if (n == 0) return cudaSuccess;
    gather_uniques_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(idx, n, t.uniques, out);
    return cudaGetLastError();
}

void exact_free(ExactTable& t) {
    cudaFree(t.words); cudaFree(t.alt); cudaFree(t.cursor); cudaFree(t.tmp); cudaFree(t.rk); cudaFree(t.rc);
    cudaFree(t.keys); cudaFree(t.counts); cudaFree(t.uniques); cudaFree(t.flags);
    t = ExactTable{};
}

}  // namespace nk
