// nk_post.cu — ONE cooperative launch for everything after the count kernel (sm_100a):
//   fold (u32 acc -> u64 currents)            src/spiking_hash.rs:174-176, :463-465
//   LIF via the per-count table               src/spiking_hash.rs:187-200 / :544-659 (see nk_lif.cu)
//   EnergyTracker total                       src/models.rs:159-172
//   top-N of the spike counts                 src/spiking_hash.rs:661-673 (see nk_topn.cu)
// The separate kernels (fold, LIF apply, 8 top-N launches, three small D2H copies) cost ~0.2 ms
// of launch latency and pool re-reads per job against 0.9 ms of counting; here the phases are
// separated by grid-wide barriers of a persistent grid (cooperative groups) and the result
// (scalars + sorted top-N rows) lands in one contiguous pack for a single D2H copy.
// Used when the LIF table path applies and N <= 2048; otherwise the separate kernels run.
#include <cooperative_groups.h>

#include "nk_kernels.cuh"

namespace cg = cooperative_groups;

namespace nk {

namespace {

struct PeerMail {
    unsigned char* mail[DIST_MAX_WORLD];
};

constexpr int PT = 256;                               // threads per block
// Work partition (ncu, profiles/r01_post_kernel_ncu.md): with 4096-neuron segments dealt round-robin,
// 2 M neurons are 489 segments for 444 resident blocks — 45 blocks do two rounds while 399 wait at the
// grid barrier, i.e. ~45 % of the kernel was barrier wait.  The elementwise phases (1, 2) therefore run
// flat grid-stride loops (every thread gets the same number of neurons +-1), and the ordered phases
// (3, 4) use 1024-neuron segments in CONTIGUOUS per-block ranges (balanced to one small segment, and
// the running tie prefix is carried instead of being re-summed per segment).
constexpr int SEG = POST_SEG_ITEMS;                   // neurons per segment of the ordered phases (1024)
constexpr int ITEMS = SEG / PT;                       // 4

__device__ __forceinline__ bool before(unsigned long long sa, unsigned long long ia, unsigned long long sb,
                                       unsigned long long ib) {
    return sa > sb || (sa == sb && ia < ib);
}

__device__ __forceinline__ unsigned block_sum(unsigned v, unsigned* s_warp) {
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    unsigned t = 0;
#pragma unroll
    for (int w = 0; w < PT / 32; ++w) t += s_warp[w];
    __syncthreads();
    return t;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// one thread: wait until flags[r] >= epoch for all r < world; false on timeout
__device__ bool wait_flags(const unsigned long long* flags, int world, unsigned long long epoch, unsigned long long timeout_ns) {
    const unsigned long long t0 = global_ns();
    for (int r = 0; r < world; ++r) {
        while (ld_acquire_sys(flags + r) < epoch) {
            if (global_ns() - t0 > timeout_ns) return false;
            __nanosleep(100);
        }
    }
    return true;
}

__global__ void dist_signal_kernel(PeerMail pm, int world, int rank, int which, unsigned long long epoch) {
    const int r = threadIdx.x;
    if (r >= world) return;
    __threadfence_system();  // this rank's counts (earlier kernels of the stream) before the flag
    st_release_sys(reinterpret_cast<unsigned long long*>(pm.mail[r]) + which * DIST_MAX_WORLD + rank, epoch);
}

// block-wide: s_pick <- the bin (walking DOWN from the largest total) that holds the `rank`-th largest value
// of the POST_EXACT_BINS-bin histogram `hist` (shared memory copy); returns through refs the value, how many
// totals are strictly greater, and how many of the ties are still needed
__device__ __forceinline__ void select_from_exact_hist(const unsigned* hist, unsigned long long n, unsigned long long* s_above,
                                                       unsigned* s_part, unsigned* s_pick, unsigned long long& T,
                                                       unsigned long long& gt, unsigned long long& need) {
    const unsigned tid = threadIdx.x;
    constexpr unsigned PER = POST_EXACT_BINS / PT;  // thread t owns the bins [8t, 8t+8)
    unsigned part = 0;
    for (unsigned b = 0; b < PER; ++b) part += hist[tid * PER + b];
    // s_above[t] = number of totals in the bins of the threads ABOVE t (suffix sum): warp shuffles, then 8 warp totals
    const unsigned lane = tid & 31u, warp = tid >> 5;
    unsigned incl = part;  // inclusive suffix sum inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_down_sync(0xFFFFFFFFu, incl, o);
        if (lane + o < 32u) incl += v;
    }
    if (lane == 0) s_part[warp] = incl;  // the warp's total
    __syncthreads();
    unsigned long long above = incl - part;
    for (unsigned w = warp + 1; w < PT / 32; ++w) above += s_part[w];
    s_above[tid] = above;
    __syncthreads();
    if (s_above[tid] < n && n <= s_above[tid] + part) *s_pick = tid;
    __syncthreads();
    const unsigned c = *s_pick;
    unsigned long long run = s_above[c];
    unsigned long long prefix = 0;
    for (int b = (int)PER - 1; b >= 0; --b) {
        const unsigned hb = hist[c * PER + b];
        if (run < n && n <= run + hb) { prefix = c * PER + b; break; }
        run += hb;
    }
    T = prefix;
    gt = run;
    need = n - run;
    __syncthreads();
}

// Phase structure (single-histogram case: every spike total < POST_EXACT_BINS, which the reference's parameters
// guarantee — a call adds at most 334):
//   1  every block owns a CONTIGUOUS neuron range: fold + LIF table look-up + write-back, and a shared-memory
//      histogram of the new totals that is (a) added to the global histogram and (b) kept;
//      -- the ONE grid barrier --
//   2  every block selects the n-th largest total T from the global histogram, takes its OWN number of ties from
//      the histogram it kept, publishes it, and sums the published counts of the blocks before it (they are all
//      resident: a short spin, no barrier);
//   3  only blocks that hold a total > T or one of the `need` lowest-index ties re-read their range and emit rows;
//   4  the last block to finish (atomic ticket) sorts the n rows, packs the result, delivers it (multi-GPU), and
//      zeroes the scratch so that the next launch needs no memset.
// Round 1 ran four grid barriers and two more passes over the pool here (profiles/r01_post_kernel_ncu.md: 8.1 barrier
// stall cycles per issued instruction).  Totals beyond the single histogram take the radix passes below (rare).
__global__ void __launch_bounds__(PT, 3) post_kernel(const PostParams q) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned int s_hist[POST_EXACT_BINS];  // 256 radix bins, or one bin per spike total (single pass)
    __shared__ unsigned long long s_above[256];
    __shared__ unsigned int s_warp[PT / 32];
    __shared__ unsigned long long s_red[PT / 32];
    __shared__ unsigned int s_pick;
    __shared__ unsigned int s_part[PT];
    __shared__ unsigned long long s_idx[2048], s_spk[2048];
    __shared__ int s_last;
    const unsigned tid = threadIdx.x;
    const LifParams& p = q.lif;
    const unsigned long long nseg = (p.pool + SEG - 1) / SEG;
    const int top = q.passes - 1;

    // ---- peer-signalled mode: every rank must have finished counting before its counts are read ----
    __shared__ int s_timeout;
    __shared__ unsigned s_spill_mask;
    unsigned long long t_start = 0, t_ready = 0, t_phase1 = 0;
    if (blockIdx.x == 0 && tid == 0) t_start = global_ns();
    if (q.wait_flags) {
        if (tid == 0) s_timeout = wait_flags(q.wait_flags, q.npeers, q.epoch, q.timeout_ns) ? 0 : 1;
        __syncthreads();
        // a timed-out block keeps going (leaving would deadlock grid.sync); the pack carries the error
        if (tid == 0 && s_timeout) atomicExch(&q.ctrl[1], 1ull);
    }
    // which ranks spilled part of this job's counts into their u64 spill array (u32 overflow guard)
    if (q.npeers) {
        if (tid == 0) s_spill_mask = 0u;
        __syncthreads();
        if (tid < (unsigned)q.npeers && q.peer_mail[tid] != nullptr &&
            ld_acquire_sys(reinterpret_cast<const unsigned long long*>(q.peer_mail[tid]) + 2 * DIST_MAX_WORLD) != 0ull)
            atomicOr(&s_spill_mask, 1u << tid);
        __syncthreads();
    }
    const unsigned spill_mask = q.npeers ? s_spill_mask : 0u;
    if (blockIdx.x == 0 && tid == 0) t_ready = global_ns();

    // ---- phase 1: fold + LIF table apply + histogram of the new spike totals --------------------------------
    const bool single = q.single_pass != 0;
    for (unsigned b = tid; b < POST_EXACT_BINS; b += PT) s_hist[b] = 0;
    __syncthreads();
    unsigned long long fired_sum = 0;
    // loads are issued in batches of B independent items (memory-level parallelism: a persistent grid
    // has only ~150 k threads for millions of neurons, so every thread must keep several loads in flight)
    // (6, not 8: 2 M neurons over 444 blocks are 17.6 per thread = 3 nearly full rounds of 6 instead of 2.2 of 8)
    constexpr int B = 6;
    unsigned int* __restrict__ acc = p.acc;
    unsigned long long* __restrict__ currents = p.currents;
    unsigned long long* __restrict__ spikes = p.spikes;
    float* __restrict__ vv = p.v;
    unsigned int* __restrict__ rr = p.r;
    const unsigned int* __restrict__ tsp = q.table.spikes;
    const float* __restrict__ tv = q.table.v;
    const unsigned int* __restrict__ tr = q.table.r;
    const unsigned long long gthreads = (unsigned long long)gridDim.x * PT;
    // this block's contiguous neuron range (a multiple of ITEMS neurons, so a thread's ITEMS consecutive neurons
    // of phase 3 never straddle two blocks)
    unsigned long long per_block = (p.pool + gridDim.x - 1) / gridDim.x;
    per_block = (per_block + ITEMS - 1) / ITEMS * ITEMS;
    const unsigned long long my_lo = (unsigned long long)blockIdx.x * per_block < p.pool ? (unsigned long long)blockIdx.x * per_block : p.pool;
    const unsigned long long my_hi = my_lo + per_block < p.pool ? my_lo + per_block : p.pool;
    {
        for (unsigned long long i0 = my_lo + tid; i0 < my_hi; i0 += (unsigned long long)PT * B) {
            unsigned long long count[B], total[B];
            bool ok[B];
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const unsigned long long i = i0 + (unsigned long long)u * PT;
                ok[u] = i < my_hi;
                count[u] = (ok[u] && p.fold_mode != 2) ? currents[i] : 0ull;
                if (q.npeers) {
                    // reduce-scatter by peer loads: this neuron's count on every rank (NVLink P2P)
                    if (ok[u]) {
                        unsigned long long sum = 0;
                        for (int r = 0; r < q.npeers; ++r) sum += __ldcg(q.peer_acc[r] + q.slice_lo + i);
                        if (spill_mask)  // rare: a rank counted more than 2^32-1 window starts in this job
                            for (int r = 0; r < q.npeers; ++r)
                                if ((spill_mask >> r) & 1u) sum += __ldcg(q.peer_spill[r] + q.slice_lo + i);
                        count[u] += sum;
                    }
                } else if (ok[u] && p.fold_mode) {
                    count[u] += acc[i];
                }
                total[u] = (ok[u] && !p.zero_state) ? spikes[i] : 0ull;
            }
            unsigned c[B], f[B], tr_[B];
            float tv_[B];
#pragma unroll
            for (int u = 0; u < B; ++u) {
                c[u] = (unsigned)(count[u] < q.table_n - 1 ? count[u] : q.table_n - 1);
                f[u] = tsp[c[u]];
                tv_[u] = tv[c[u]];
                tr_[u] = tr[c[u]];
            }
#pragma unroll
            for (int u = 0; u < B; ++u) {
                const unsigned long long i = i0 + (unsigned long long)u * PT;
                if (!ok[u]) continue;
                if (p.fold_mode) {
                    if (!q.npeers) acc[i] = 0u;  // sharded: peers may still be reading; the host clears it later
                    currents[i] = count[u];
                }
                if (!(p.skip_zero && count[u] == 0)) {
                    vv[i] = tv_[u];
                    rr[i] = tr_[u];
                    fired_sum += f[u];
                    total[u] += f[u];
                    spikes[i] = total[u];
                } else if (p.zero_state) {
                    vv[i] = 0.0f;
                    rr[i] = 0u;
                    spikes[i] = 0ull;
                }
                atomicAdd(&s_hist[single ? (unsigned)(total[u] < POST_EXACT_BINS - 1 ? total[u] : POST_EXACT_BINS - 1)
                                         : (unsigned)(total[u] >> (8 * top)) & 255u], 1u);
            }
        }
    }
    // spikes fired by this call (EnergyTracker)
    for (int o = 16; o > 0; o >>= 1) fired_sum += __shfl_down_sync(0xFFFFFFFFu, fired_sum, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = fired_sum;
    __syncthreads();
    if (tid == 0) {
        unsigned long long f = 0;
        for (int w = 0; w < PT / 32; ++w) f += s_red[w];
        if (f) atomicAdd(p.total_new, f);
    }
    if (single) {
        for (unsigned b = tid; b < POST_EXACT_BINS; b += PT)
            if (s_hist[b]) atomicAdd(&q.hist[b], s_hist[b]);
    } else if (s_hist[tid]) {
        atomicAdd(&q.hist[top * 256 + tid], s_hist[tid]);
    }
    grid.sync();
    if (blockIdx.x == 0 && tid == 0) {
        t_phase1 = global_ns();
        q.ctrl[4] = t_start; q.ctrl[5] = t_ready; q.ctrl[6] = t_phase1;  // for whichever block packs the result
    }

    if (single) {
        // ---- phase 2: select T from the global histogram; ties of the blocks before this one -----------------
        unsigned* s_ghist = reinterpret_cast<unsigned*>(s_idx);  // s_idx is not needed before the final sort
        for (unsigned b = tid; b < POST_EXACT_BINS; b += PT) s_ghist[b] = __ldcg(&q.hist[b]);
        __syncthreads();
        unsigned long long T, gt, need;
        select_from_exact_hist(s_ghist, q.n, s_above, s_part, &s_pick, T, gt, need);
        if (q.trace && blockIdx.x == 0 && tid == 0) q.ctrl[3] = global_ns();
        const unsigned my_ties = s_hist[T];                       // this block's own range (kept from phase 1)
        unsigned my_gt = 0;
        for (unsigned b = (unsigned)T + 1 + tid; b < POST_EXACT_BINS; b += PT) my_gt += s_hist[b];
        my_gt = block_sum(my_gt, s_warp);
        if (tid == 0) {
            // published as count + 1 (0 = not yet); release so that a reader that sees it sees a finished value
            unsigned* slot = q.block_ties + blockIdx.x;
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(slot), "r"(my_ties + 1u) : "memory");
        }
        unsigned long long ties_before = 0;
        if (my_ties != 0) {  // (a block without ties has no use for the prefix)
            for (unsigned bb = tid; bb < blockIdx.x; bb += PT) {
                unsigned v;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(q.block_ties + bb) : "memory");
                } while (v == 0u);
                ties_before += v - 1u;
            }
            for (int o = 16; o > 0; o >>= 1) ties_before += __shfl_down_sync(0xFFFFFFFFu, ties_before, o);
            if ((tid & 31) == 0) s_red[tid >> 5] = ties_before;
            __syncthreads();
            ties_before = 0;
            for (int w = 0; w < PT / 32; ++w) ties_before += s_red[w];
            __syncthreads();
        }
        if (q.trace && blockIdx.x == 0 && tid == 0) q.ctrl[7] = global_ns();
        // ---- phase 3: ordered gather, only where this block contributes rows ---------------------------------
        if (my_gt != 0 || (my_ties != 0 && ties_before < need)) {
            unsigned long long seg_prefix = ties_before;
            for (unsigned long long base0 = my_lo; base0 < my_hi; base0 += SEG) {
                // thread t owns ITEMS consecutive neurons (index order)
                const unsigned long long base = base0 + (unsigned long long)tid * ITEMS;
                unsigned long long v[ITEMS];
                unsigned eq = 0;
#pragma unroll
                for (int it = 0; it < ITEMS; ++it) {
                    const unsigned long long i = base + it;
                    v[it] = i < my_hi ? spikes[i] : 0ull;
                    if (i < my_hi && v[it] == T) ++eq;
                    if (i < my_hi && v[it] > T) {
                        const unsigned long long slot = atomicAdd(&q.ctrl[0], 1ull);
                        q.out_idx[slot] = i;
                        q.out_spikes[slot] = v[it];
                    }
                }
                // block-uniform: only while ties are still wanted, and only in segments that hold one
                if (seg_prefix >= need) {
                    if (my_gt == 0) break;  // nothing above T in this range and all wanted ties are placed
                } else if (__syncthreads_or(eq != 0)) {
                    s_part[tid] = eq;
                    __syncthreads();
                    for (int o = 1; o < PT; o <<= 1) {
                        const unsigned a = tid >= (unsigned)o ? s_part[tid - o] : 0u;
                        __syncthreads();
                        s_part[tid] += a;
                        __syncthreads();
                    }
                    const unsigned seg_ties = s_part[PT - 1];
                    unsigned long long r = seg_prefix + (s_part[tid] - eq);
                    if (eq != 0 && r < need) {
#pragma unroll
                        for (int it = 0; it < ITEMS; ++it) {
                            const unsigned long long i = base + it;
                            if (i < my_hi && v[it] == T) {
                                if (r < need) {
                                    q.out_idx[gt + r] = i;
                                    q.out_spikes[gt + r] = T;
                                }
                                ++r;
                            }
                        }
                    }
                    __syncthreads();
                    seg_prefix += seg_ties;
                }
            }
        }
        // ---- the last block to get here does the rest --------------------------------------------------------
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(&q.ctrl[2], 1ull) == (unsigned long long)gridDim.x - 1ull ? 1 : 0;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        if (q.trace && tid == 0) q.pack[10] = global_ns();
    } else {
    // ---- radix passes (totals beyond the single histogram): MSB-first select, every block redundantly ----
    unsigned long long prefix = 0, rank = q.n, gt = 0;
    for (int d = top; d >= 0; --d) {
        if (d != top) {
            s_hist[tid] = 0;
            __syncthreads();
            const int hs = 8 * (d + 1);
            for (unsigned long long i0 = (unsigned long long)blockIdx.x * PT + tid; i0 < p.pool; i0 += gthreads * B) {
                unsigned long long v[B];
#pragma unroll
                for (int it = 0; it < B; ++it) {
                    const unsigned long long i = i0 + (unsigned long long)it * gthreads;
                    v[it] = i < p.pool ? spikes[i] : ~0ull;
                }
#pragma unroll
                for (int it = 0; it < B; ++it) {
                    const unsigned long long i = i0 + (unsigned long long)it * gthreads;
                    if (i < p.pool && (v[it] >> hs) == prefix) atomicAdd(&s_hist[(v[it] >> (8 * d)) & 255u], 1u);
                }
            }
            __syncthreads();
            if (s_hist[tid]) atomicAdd(&q.hist[d * 256 + tid], s_hist[tid]);
            grid.sync();
        }
        const unsigned hb = __ldcg(&q.hist[d * 256 + tid]);
        s_hist[tid] = hb;
        __syncthreads();
        if (tid == 0) {
            unsigned long long run = 0;
            for (int x = 255; x >= 0; --x) { s_above[x] = run; run += s_hist[x]; }
        }
        __syncthreads();
        if (s_above[tid] < rank && rank <= s_above[tid] + hb) s_pick = tid;
        __syncthreads();
        const unsigned b = s_pick;
        prefix = (prefix << 8) | b;
        gt += s_above[b];
        rank -= s_above[b];
        __syncthreads();
    }
    const unsigned long long T = prefix, need = rank;

    // ties per segment (each block owns a contiguous range of segments)
    const unsigned long long seg_per = (nseg + gridDim.x - 1) / gridDim.x;
    const unsigned long long seg_lo = (unsigned long long)blockIdx.x * seg_per;
    const unsigned long long seg_hi = seg_lo + seg_per < nseg ? seg_lo + seg_per : nseg;
    for (unsigned long long seg = seg_lo; seg < seg_hi; ++seg) {
        unsigned c = 0;
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const unsigned long long i = seg * SEG + (unsigned long long)it * PT + tid;
            if (i < p.pool && spikes[i] == T) ++c;
        }
        c = block_sum(c, s_warp);
        if (tid == 0) q.seg_counts[seg] = c;
    }
    grid.sync();

    // ordered gather: every total > T, and the `need` lowest-index totals == T
    if (seg_lo < seg_hi) {
        unsigned long long before_me = 0;  // ties in the segments before this block's range
        for (unsigned long long s = tid; s < seg_lo; s += PT) before_me += __ldcg(&q.seg_counts[s]);
        for (int o = 16; o > 0; o >>= 1) before_me += __shfl_down_sync(0xFFFFFFFFu, before_me, o);
        if ((tid & 31) == 0) s_red[tid >> 5] = before_me;
        __syncthreads();
        unsigned long long seg_prefix = 0;
        for (int w = 0; w < PT / 32; ++w) seg_prefix += s_red[w];
        __syncthreads();
        for (unsigned long long seg = seg_lo; seg < seg_hi; ++seg) {
            // thread t owns ITEMS consecutive neurons (index order)
            const unsigned long long base = seg * SEG + (unsigned long long)tid * ITEMS;
            unsigned long long v[ITEMS];
            unsigned eq = 0;
#pragma unroll
            for (int it = 0; it < ITEMS; ++it) {
                const unsigned long long i = base + it;
                v[it] = i < p.pool ? spikes[i] : 0ull;
                if (i < p.pool && v[it] == T) ++eq;
                if (i < p.pool && v[it] > T) {
                    const unsigned long long slot = atomicAdd(&q.ctrl[0], 1ull);
                    q.out_idx[slot] = i;
                    q.out_spikes[slot] = v[it];
                }
            }
            const unsigned seg_ties = __ldcg(&q.seg_counts[seg]);
            if (seg_prefix < need && seg_ties != 0) {  // block-uniform: only segments that still contribute ties scan
                s_hist[tid] = eq;
                __syncthreads();
                for (int o = 1; o < PT; o <<= 1) {
                    const unsigned a = tid >= (unsigned)o ? s_hist[tid - o] : 0u;
                    __syncthreads();
                    s_hist[tid] += a;
                    __syncthreads();
                }
                unsigned long long r = seg_prefix + (s_hist[tid] - eq);
                if (eq != 0 && r < need) {
#pragma unroll
                    for (int it = 0; it < ITEMS; ++it) {
                        const unsigned long long i = base + it;
                        if (i < p.pool && v[it] == T) {
                            if (r < need) {
                                q.out_idx[gt + r] = i;
                                q.out_spikes[gt + r] = T;
                            }
                            ++r;
                        }
                    }
                }
                __syncthreads();
            }
            seg_prefix += seg_ties;
        }
    }
    grid.sync();
    if (blockIdx.x != 0) return;
    }  // radix passes

    // ---- final phase (one block): sort the n candidates (spikes desc, idx asc) and pack the result ----
    const unsigned n = (unsigned)q.n;
    unsigned N = 1;
    while (N < n) N <<= 1;
    __syncthreads();
    // the scalars of the pack: loaded now, so that their L2 round trips overlap the sort
    unsigned long long pk_fired = 0, pk_err = 0, pk_kmers = 0, pk_t0 = 0, pk_t1 = 0, pk_t2 = 0;
    if (tid == 0) {
        pk_fired = __ldcg(p.total_new);
        pk_err = __ldcg(&q.ctrl[1]);
        pk_kmers = __ldcg(q.kmers);
        pk_t0 = __ldcg(&q.ctrl[4]); pk_t1 = __ldcg(&q.ctrl[5]); pk_t2 = __ldcg(&q.ctrl[6]);
    }
    for (unsigned i = tid; i < N; i += PT) {
        s_idx[i] = i < n ? __ldcg(&q.out_idx[i]) : ~0ull;
        s_spk[i] = i < n ? __ldcg(&q.out_spikes[i]) : 0ull;
    }
    __syncthreads();
    for (unsigned k = 2; k <= N; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = tid; i < N; i += PT) {
                const unsigned l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const bool swap = up ? before(s_spk[l], s_idx[l], s_spk[i], s_idx[i])
                                         : before(s_spk[i], s_idx[i], s_spk[l], s_idx[l]);
                    if (swap) {
                        const unsigned long long a = s_idx[i], b = s_spk[i];
                        s_idx[i] = s_idx[l]; s_spk[i] = s_spk[l];
                        s_idx[l] = a; s_spk[l] = b;
                    }
                }
            }
            __syncthreads();
        }
    }
    // pack: [0] spikes fired by this call, [1] error, [2] k-mers of the call, [3] n, [4..11] time stamps
    // (%globaltimer ns of THIS GPU: kernel start, peers' counts readable, end of the fused reduce + LIF
    // phase, pack written; [8..10] belong to the merge kernel), then idx[n], spikes[n]
    for (unsigned i = tid; i < n; i += PT) {
        q.out_idx[i] = s_idx[i];
        q.out_spikes[i] = s_spk[i];
        q.pack[PACK_HDR + i] = s_idx[i] + q.slice_lo;
        q.pack[PACK_HDR + n + i] = s_spk[i];
    }
    if (tid == 0) {
        q.pack[0] = pk_fired;
        q.pack[1] = pk_err;  // 1: a peer's "counting finished" signal timed out
        q.pack[2] = pk_kmers;
        q.pack[3] = n;
        q.pack[4] = pk_t0; q.pack[5] = pk_t1; q.pack[6] = pk_t2;  // block 0's stamps travel through the scratch
        q.pack[7] = global_ns();
        if (q.trace) { q.pack[8] = __ldcg(&q.ctrl[3]); q.pack[9] = __ldcg(&q.ctrl[7]); q.pack[11] = blockIdx.x; }
    }
    if (q.wait_flags) {
        // deliver the pack into every rank's mailbox (slot = this rank), then raise "pack delivered"
        __syncthreads();
        const unsigned words = (unsigned)PACK_HDR + 2u * n;
        for (int r = 0; r < q.npeers; ++r) {
            unsigned long long* slot = reinterpret_cast<unsigned long long*>(q.peer_mail[r]) + DIST_FLAG_ROWS * DIST_MAX_WORLD +
                                       (unsigned long long)q.rank * DIST_PACK_SLOT_U64;
            for (unsigned i = tid; i < words; i += PT) slot[i] = q.pack[i];
        }
        __threadfence_system();
        __syncthreads();
        if (tid < (unsigned)q.npeers)
            st_release_sys(reinterpret_cast<unsigned long long*>(q.peer_mail[tid]) + DIST_MAX_WORLD + q.rank, q.epoch);
    }
    // leave the scratch as the next launch wants it: no memsets between jobs
    __syncthreads();
    for (unsigned b = tid; b < POST_EXACT_BINS; b += PT) q.hist[b] = 0u;
    for (unsigned b = tid; b < POST_MAX_GRID; b += PT) q.block_ties[b] = 0u;
    if (tid < 8) q.ctrl[tid] = 0ull;
    if (tid == 0) { *p.total_new = 0ull; *const_cast<unsigned long long*>(q.kmers) = 0ull; }
}

// one block: sum the scalars, sort the union of the per-rank rows, keep the best n_out
// `flags` != null (peer-signalled mode): first wait until every rank's pack has been delivered.
// No __restrict__ / non-coherent loads on g: in that mode it is written by the peers while this kernel waits.
__global__ void __launch_bounds__(PT) merge_packs_kernel(const unsigned long long* g, int world,
                                                         unsigned long long n_each, unsigned long long stride,
                                                         unsigned long long n_out, unsigned long long* out,
                                                         const unsigned long long* flags, unsigned long long epoch,
                                                         unsigned long long timeout_ns, int rank) {
    __shared__ unsigned long long s_idx[2048], s_spk[2048];
    __shared__ int s_err;
    const unsigned tid = threadIdx.x;
    unsigned long long t_start = 0, t_ready = 0;
    if (tid == 0) { s_err = 0; t_start = global_ns(); }
    if (flags) {
        if (tid == 0) s_err = wait_flags(flags, world, epoch, timeout_ns) ? 0 : 2;
        __syncthreads();
        __threadfence_system();
        if (s_err) {  // a rank never delivered its pack: report, do not touch the slots
            if (tid == 0) { out[0] = 0; out[1] = (unsigned long long)s_err; out[2] = 0; out[3] = 0; }
            return;
        }
    }
    if (tid == 0) t_ready = global_ns();
    unsigned total = 0;
    for (int r = 0; r < world; ++r) total += (unsigned)__ldcg(&g[r * stride + 3]);
    unsigned N = 1;
    while (N < total) N <<= 1;
    for (unsigned i = tid; i < N; i += PT) { s_idx[i] = ~0ull; s_spk[i] = 0ull; }
    __syncthreads();
    unsigned off = 0;
    for (int r = 0; r < world; ++r) {
        const unsigned nr = (unsigned)__ldcg(&g[r * stride + 3]);
        for (unsigned i = tid; i < nr; i += PT) {
            s_idx[off + i] = __ldcg(&g[r * stride + PACK_HDR + i]);
            s_spk[off + i] = __ldcg(&g[r * stride + PACK_HDR + nr + i]);
        }
        off += nr;
    }
    __syncthreads();
    for (unsigned k = 2; k <= N; k <<= 1) {
        for (unsigned j = k >> 1; j > 0; j >>= 1) {
            for (unsigned i = tid; i < N; i += PT) {
                const unsigned l = i ^ j;
                if (l > i) {
                    const bool up = (i & k) == 0;
                    const bool swap = up ? before(s_spk[l], s_idx[l], s_spk[i], s_idx[i])
                                         : before(s_spk[i], s_idx[i], s_spk[l], s_idx[l]);
                    if (swap) {
                        const unsigned long long a = s_idx[i], b = s_spk[i];
                        s_idx[i] = s_idx[l]; s_spk[i] = s_spk[l];
                        s_idx[l] = a; s_spk[l] = b;
                    }
                }
            }
            __syncthreads();
        }
    }
    const unsigned n = (unsigned)(n_out < total ? n_out : total);
    for (unsigned i = tid; i < n; i += PT) {
        out[PACK_HDR + i] = s_idx[i];
        out[PACK_HDR + n + i] = s_spk[i];
    }
    if (tid == 0) {
        unsigned long long fired = 0, kmers = 0, err = (unsigned long long)s_err;
        for (int r = 0; r < world; ++r) { fired += __ldcg(&g[r * stride + 0]); err |= __ldcg(&g[r * stride + 1]); kmers += __ldcg(&g[r * stride + 2]); }
        out[0] = fired; out[1] = err; out[2] = kmers; out[3] = n;
        // this rank's own slice-kernel stamps travel with the merged pack; then the merge kernel's
        for (int x = 4; x < 8; ++x) out[x] = __ldcg(&g[rank * stride + x]);
        out[8] = t_start; out[9] = t_ready; out[10] = global_ns(); out[11] = 0;
    }
}

}  // namespace

cudaError_t launch_merge_packs(const unsigned long long* gathered, int world, int rank, unsigned long long n_each,
                               unsigned long long n_out, unsigned long long* pack_out, cudaStream_t s) {
    merge_packs_kernel<<<1, PT, 0, s>>>(gathered, world, n_each, PACK_HDR + 2 * n_each, n_out, pack_out, nullptr, 0, 0, rank);
    return cudaGetLastError();
}

cudaError_t launch_merge_mailbox(unsigned char* mail, int world, int rank, unsigned long long n_each, unsigned long long n_out,
                                 unsigned long long epoch, unsigned long long timeout_ns, unsigned long long* pack_out,
                                 cudaStream_t s) {
    merge_packs_kernel<<<1, PT, 0, s>>>(dist_mail_slot(mail, 0), world, n_each, DIST_PACK_SLOT_U64, n_out, pack_out,
                                        dist_mail_flags(mail, 1), epoch, timeout_ns, rank);
    return cudaGetLastError();
}

cudaError_t launch_dist_signal(unsigned char* const* peer_mail, int world, int rank, int which, unsigned long long epoch,
                               cudaStream_t s) {
    PeerMail pm{};
    for (int r = 0; r < world; ++r) pm.mail[r] = peer_mail[r];
    dist_signal_kernel<<<1, 32, 0, s>>>(pm, world, rank, which, epoch);
    return cudaGetLastError();
}

namespace {
}  // namespace

cudaError_t post_max_grid(int device, int* grid) {
    int sms = 0, per_sm = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, post_kernel, PT, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 6) per_sm = 6;
    *grid = sms * per_sm;
    return cudaSuccess;
}

cudaError_t launch_post(const PostParams& q, int max_grid, cudaStream_t s) {
    const unsigned long long nseg = (q.lif.pool + SEG - 1) / SEG;
    int grid = (int)(nseg < (unsigned long long)max_grid ? nseg : (unsigned long long)max_grid);
    if (grid < 1) grid = 1;
    if (grid > POST_MAX_GRID) grid = POST_MAX_GRID;
    void* args[] = {(void*)&q};
    return cudaLaunchCooperativeKernel((const void*)post_kernel, dim3(grid), dim3(PT), args, 0, s);
}

}  // namespace nk
