// nk_lif.cu — pool fold + Leaky Integrate-and-Fire simulation (sm_100a).
//
// Replaces:
//   currents store / overwrite                 src/spiking_hash.rs:174-176, :463-465
//   LifNeuron::update                          src/models.rs:34-51
//   in-memory LIF driver (skips zero current)  src/spiking_hash.rs:187-200
//   simulate_spikes_simd (steps every neuron)  src/spiking_hash.rs:544-659
//   process_sequence's single tick             src/spiking_hash.rs:266-272
//   EnergyTracker totals                       src/models.rs:159-172
//
// Arithmetic is IEEE f32 with a SEPARATE multiply and add (__fmul_rn/__fadd_rn: rustc
// and _mm256_mul_ps/_mm256_add_ps never fuse), input current = f32(f64(count)/f64(steps)).
//
// State invariant used below: refractory_ticks > 0 implies voltage == 0 (refractory is
// only ever set together with voltage = 0 and voltage is frozen while it counts down),
// so "inactive lane keeps its voltage" == "inactive lane has voltage 0", and the SIMD
// driver's inactive-lane compare against f32::MAX (:611-613) can never fire.
#include "nk_kernels.cuh"

namespace nk {

namespace {

constexpr int LIF_THREADS = 256;

__global__ void fold_kernel(unsigned int* __restrict__ acc, unsigned long long* __restrict__ currents,
                            unsigned long long pool, int overwrite) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < pool; i += stride) {
        const unsigned a = acc[i];
        currents[i] = (overwrite ? 0ull : currents[i]) + a;
        acc[i] = 0u;
    }
}

// u64 spill array (sharded-pool handles) -> currents; the spill array is left as it is (overwritten by its next use)
__global__ void fold64_kernel(const unsigned long long* __restrict__ spill, unsigned long long* __restrict__ currents,
                              unsigned long long pool, int overwrite) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < pool; i += stride)
        currents[i] = (overwrite ? 0ull : currents[i]) + spill[i];
}

// Single-process multi-GPU groups, carried-state path: the leader GPU sums every member's u32 counts (and the
// u64 spill arrays of the members that spilled) straight out of peer memory into its u64 currents
// (= the reference's reduce over per-thread vectors, src/spiking_hash.rs:145-154, then :174-176).
__global__ void peer_sum_kernel(const PeerSumParams ps, unsigned long long* __restrict__ currents,
                                unsigned long long pool, int overwrite, unsigned long long* kmers_out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < pool; i += stride) {
        unsigned long long sum = overwrite ? 0ull : currents[i];
        for (int r = 0; r < ps.n; ++r) sum += __ldcg(ps.acc[r] + i);
        if (ps.spill_mask)
            for (int r = 0; r < ps.n; ++r)
                if ((ps.spill_mask >> r) & 1u) sum += __ldcg(ps.spill[r] + i);
        currents[i] = sum;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long k = 0;
        for (int r = 0; r < ps.n; ++r) k += __ldcg(ps.kmers[r]);
        *kmers_out = k;
    }
}

struct NeuronResult {
    float v;
    unsigned r;
    unsigned fired;
};

// `steps` ticks of one neuron.  Refractory state is carried as the index of the next
// active tick (t_next): active at tick t iff t >= t_next; a spike at tick t sets
// t_next = t + period + 1.  Equivalent to the reference's countdown:
//   r > 0  -> r -= 1, no integration              (models.rs:35-38)
//   else   -> v = v*leak + I; v >= thr -> spike   (models.rs:41-47)
__device__ __forceinline__ NeuronResult lif_run(float v, unsigned r, float I, unsigned long long steps,
                                                float thr, float leak, unsigned period) {
    unsigned long long t_next = r;  // first active tick
    unsigned fired = 0;
    const unsigned long long p1 = (unsigned long long)period + 1ull;
    if (steps < 0x7FFFFFFFull && p1 < 0x7FFFFFFFull && r < 0x7FFFFFFFu) {
        // 32-bit tick arithmetic (every realistic configuration)
        const unsigned n = (unsigned)steps, q1 = (unsigned)p1;
        unsigned tn = r;
#pragma unroll 8
        for (unsigned t = 0; t < n; ++t) {
            const bool active = t >= tn;
            const float vn = __fadd_rn(__fmul_rn(v, leak), I);
            const bool spike = active && (vn >= thr);
            const bool keep = active && !(vn >= thr);
            v = keep ? vn : (active ? 0.0f : v);
            if (spike) {
                ++fired;
                tn = t + q1;
            }
        }
        NeuronResult o;
        o.v = v;
        o.r = tn > n ? tn - n : 0u;
        o.fired = fired;
        return o;
    }
    for (unsigned long long t = 0; t < steps; ++t) {
        if (t < t_next) continue;
        v = __fadd_rn(__fmul_rn(v, leak), I);
        if (v >= thr) {
            v = 0.0f;
            ++fired;
            t_next = t + p1;
        }
    }
    NeuronResult o;
    o.v = v;
    o.r = t_next > steps ? (unsigned)(t_next - steps) : 0u;
    o.fired = fired;
    return o;
}

__device__ __forceinline__ float input_current(unsigned long long count, unsigned long long steps) {
    // (total_current / steps as f64) as f32      src/spiking_hash.rs:193-196, :581-582
    return __double2float_rn(__ull2double_rn(count) / __ull2double_rn(steps));
}

// this call's total for neuron i: optionally folds the u32 batch accumulator in (the fold
// kernel fused into the LIF pass: saves a launch and one 24 MB round trip over the pool)
__device__ __forceinline__ unsigned long long load_count(const LifParams& p, unsigned long long i) {
    unsigned long long count = p.fold_mode == 2 ? 0ull : p.currents[i];
    if (p.fold_mode) {
        count += p.acc[i];
        p.acc[i] = 0u;
        p.currents[i] = count;
    }
    return count;
}

__device__ __forceinline__ void block_totals(unsigned long long fired, unsigned long long maxs,
                                             unsigned long long* total_new, unsigned long long* max_spikes) {
    __shared__ unsigned long long s_f[LIF_THREADS / 32], s_m[LIF_THREADS / 32];
    for (int o = 16; o > 0; o >>= 1) {
        fired += __shfl_down_sync(0xFFFFFFFFu, fired, o);
        const unsigned long long m = __shfl_down_sync(0xFFFFFFFFu, maxs, o);
        maxs = m > maxs ? m : maxs;
    }
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (lane == 0) { s_f[warp] = fired; s_m[warp] = maxs; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long f = 0, m = 0;
        for (int w = 0; w < LIF_THREADS / 32; ++w) { f += s_f[w]; m = s_m[w] > m ? s_m[w] : m; }
        if (f) atomicAdd(total_new, f);
        if (m) atomicMax(max_spikes, m);
    }
}

__global__ void __launch_bounds__(LIF_THREADS) lif_kernel(const LifParams p, const unsigned long long* only_if) {
    if (only_if && *only_if == 0ull) return;  // device-side fallback of the memo path: nothing to do
    const unsigned long long i = blockIdx.x * (unsigned long long)LIF_THREADS + threadIdx.x;
    unsigned long long fired = 0, total = 0;
    if (i < p.pool) {
        const unsigned long long count = load_count(p, i);
        total = p.spikes[i];
        if (!(p.skip_zero && count == 0)) {
            const NeuronResult o =
                lif_run(p.v[i], p.r[i], input_current(count, p.steps), p.steps, p.thr, p.leak, p.period);
            p.v[i] = o.v;
            p.r[i] = o.r;
            fired = o.fired;
            total += fired;
            p.spikes[i] = total;
        }
    }
    block_totals(fired, total, p.total_new, p.max_spikes);
}

// ---- uniform-fresh-state fast path ------------------------------------------------
// While every neuron is still in its initial state (v = 0, r = 0) the result of the
// simulation is a pure function of the neuron's count, and all counts >= c_sat (the
// first count whose input current reaches the threshold) follow the same trajectory
// (tick 0: v = 0*leak + I >= thr -> spike -> v = 0).  So: simulate once per distinct
// count value c in [0, table_n) and look the result up per neuron.  Bit-identical to
// lif_kernel (tests/test_parity_gpu.py::test_lif_table_equals_direct).
__global__ void __launch_bounds__(LIF_THREADS) lif_table_build_kernel(const LifParams p, const LifTable t,
                                                                     unsigned long long table_n) {
    const unsigned long long c = blockIdx.x * (unsigned long long)LIF_THREADS + threadIdx.x;
    if (c >= table_n) return;
    const NeuronResult o = lif_run(0.0f, 0u, input_current(c, p.steps), p.steps, p.thr, p.leak, p.period);
    t.spikes[c] = o.fired;
    t.v[c] = o.v;
    t.r[c] = o.r;
}

__global__ void __launch_bounds__(LIF_THREADS) lif_table_apply_kernel(const LifParams p, const LifTable t,
                                                                     unsigned long long table_n) {
    const unsigned long long i = blockIdx.x * (unsigned long long)LIF_THREADS + threadIdx.x;
    unsigned long long fired = 0, total = 0;
    if (i < p.pool) {
        const unsigned long long count = load_count(p, i);
        total = p.zero_state ? 0ull : p.spikes[i];
        if (!(p.skip_zero && count == 0)) {
            const unsigned long long c = count < table_n - 1 ? count : table_n - 1;
            p.v[i] = t.v[c];
            p.r[i] = t.r[c];
            fired = t.spikes[c];
            total += fired;
            p.spikes[i] = total;
        } else if (p.zero_state) {  // lazily-zero pool: the untouched neuron's state is written here
            p.v[i] = 0.0f;
            p.r[i] = 0u;
            p.spikes[i] = 0ull;
        }
    }
    block_totals(fired, total, p.total_new, p.max_spikes);
}

// ---- carried-state fast path: memoised simulation --------------------------------------------------
// Key = (state, clamped count) in 54 bits.  State: the invariant above (r > 0 implies v == 0) makes it either a
// voltage (r == 0) or a refractory count (v == 0): 33 bits.  Count: clamped to `sat`, the first count whose input
// current reaches the threshold — with leak >= 0 the voltage is never negative, so for such a count
// fl(fl(v*leak) + I) >= I >= thr on EVERY active tick whatever v is: all saturating counts follow one
// trajectory from any reachable state (the argument of the fresh-state table, extended to carried state).
constexpr unsigned long long MEMO_EMPTY = ~0ull;
constexpr unsigned int MEMO_NONE = 0xFFFFFFFFu;

__device__ __forceinline__ unsigned long long memo_key(float v, unsigned r, unsigned long long c) {
    const unsigned long long state = r ? ((1ull << 32) | r) : (unsigned long long)__float_as_uint(v);
    return (state << 21) | c;
}

__global__ void __launch_bounds__(LIF_THREADS) lif_memo_insert_kernel(const LifParams p, const LifMemo m, unsigned long long sat) {
    const unsigned long long i = blockIdx.x * (unsigned long long)LIF_THREADS + threadIdx.x;
    if (i >= p.pool) return;
    const unsigned long long count = load_count(p, i);  // folds the batch accumulators in (once: the fallback must not)
    if (p.skip_zero && count == 0) {
        m.slot_of[i] = MEMO_NONE;
        return;
    }
    const unsigned long long key = memo_key(p.v[i], p.r[i], count < sat ? count : sat);
    unsigned long long slot = splitmix64(key) & (LIF_MEMO_SLOTS - 1);
    for (unsigned long long probe = 0; probe < LIF_MEMO_SLOTS; ++probe) {
        if (*reinterpret_cast<volatile unsigned long long*>(m.ctrl + 1)) break;  // already given up: do not fill the table
        // most neurons find their key already present: a plain load first keeps them off the atomic unit
        unsigned long long seen = *reinterpret_cast<volatile unsigned long long*>(m.keys + slot);
        if (seen == key) break;
        if (seen == MEMO_EMPTY) seen = atomicCAS(m.keys + slot, MEMO_EMPTY, key);
        if (seen == MEMO_EMPTY) {  // this thread inserted the key
            if (atomicAdd(m.ctrl + 0, 1ull) >= LIF_MEMO_SLOTS / 2) m.ctrl[1] = 1ull;  // too diverse: direct kernel
            break;
        }
        if (seen == key) break;
        slot = (slot + 1) & (LIF_MEMO_SLOTS - 1);
        if (probe + 1 == LIF_MEMO_SLOTS) m.ctrl[1] = 1ull;
    }
    m.slot_of[i] = (unsigned)slot;
}

__global__ void __launch_bounds__(LIF_THREADS) lif_memo_compact_kernel(const LifMemo m) {
    if (m.ctrl[1]) return;
    const unsigned long long slot = blockIdx.x * (unsigned long long)LIF_THREADS + threadIdx.x;
    if (slot < LIF_MEMO_SLOTS && m.keys[slot] != MEMO_EMPTY) m.dense[atomicAdd(m.ctrl + 2, 1ull)] = (unsigned)slot;
}

__global__ void __launch_bounds__(LIF_THREADS) lif_memo_run_kernel(const LifParams p, const LifMemo m) {
    if (m.ctrl[1]) return;
    const unsigned long long id = blockIdx.x * (unsigned long long)LIF_THREADS + threadIdx.x;
    if (id >= m.ctrl[0]) return;
    const unsigned slot = m.dense[id];
    const unsigned long long key = m.keys[slot];
    const unsigned long long c = key & ((1ull << 21) - 1), state = key >> 21;
    const bool refr = (state >> 32) != 0;
    const NeuronResult o = lif_run(refr ? 0.0f : __uint_as_float((unsigned)state), refr ? (unsigned)state : 0u,
                                   input_current(c, p.steps), p.steps, p.thr, p.leak, p.period);
    m.res_v[slot] = o.v;
    m.res_r[slot] = o.r;
    m.res_f[slot] = o.fired;
}

__global__ void __launch_bounds__(LIF_THREADS) lif_memo_apply_kernel(const LifParams p, const LifMemo m) {
    if (m.ctrl[1]) return;  // block-uniform: the direct kernel behind this one does the work
    const unsigned long long i = blockIdx.x * (unsigned long long)LIF_THREADS + threadIdx.x;
    unsigned long long fired = 0, total = 0;
    if (i < p.pool) {
        total = p.spikes[i];
        const unsigned slot = m.slot_of[i];
        if (slot != MEMO_NONE) {
            p.v[i] = m.res_v[slot];
            p.r[i] = m.res_r[slot];
            fired = m.res_f[slot];
            total += fired;
            p.spikes[i] = total;
        }
    }
    block_totals(fired, total, p.total_new, p.max_spikes);
}

// process_sequence tail (src/spiking_hash.rs:266-272): one update(count as f32) per
// neuron with count > 0, then currents[i] = 0.
__global__ void __launch_bounds__(LIF_THREADS) lif_single_tick_kernel(const LifParams p,
                                                                     unsigned long long* currents_rw) {
    const unsigned long long i = blockIdx.x * (unsigned long long)LIF_THREADS + threadIdx.x;
    unsigned long long fired = 0, total = 0;
    if (i < p.pool) {
        const unsigned long long count = currents_rw[i];
        total = p.spikes[i];
        if (count > 0) {
            float v = p.v[i];
            unsigned r = p.r[i];
            if (r > 0) {
                r -= 1;
            } else {
                v = __fadd_rn(__fmul_rn(v, p.leak), __double2float_rn(__ull2double_rn(count)));
                if (v >= p.thr) { v = 0.0f; r = p.period; fired = 1; }
            }
            p.v[i] = v;
            p.r[i] = r;
            total += fired;
            p.spikes[i] = total;
            currents_rw[i] = 0;
        }
    }
    block_totals(fired, total, p.total_new, p.max_spikes);
}

}  // namespace

cudaError_t launch_fold(unsigned int* acc, unsigned long long* currents, unsigned long long pool,
                        bool overwrite, cudaStream_t s) {
    unsigned long long blocks = (pool + 255) / 256;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    if (blocks == 0) blocks = 1;
    fold_kernel<<<(unsigned)blocks, 256, 0, s>>>(acc, currents, pool, overwrite ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_fold64(const unsigned long long* spill, unsigned long long* currents, unsigned long long pool,
                          bool overwrite, cudaStream_t s) {
    unsigned long long blocks = (pool + 255) / 256;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    if (blocks == 0) blocks = 1;
    fold64_kernel<<<(unsigned)blocks, 256, 0, s>>>(spill, currents, pool, overwrite ? 1 : 0);
    return cudaGetLastError();
}

cudaError_t launch_peer_sum(const PeerSumParams& ps, unsigned long long* currents, unsigned long long pool, bool overwrite,
                            unsigned long long* kmers_out, cudaStream_t s) {
    unsigned long long blocks = (pool + 255) / 256;
    if (blocks > 148ull * 16) blocks = 148ull * 16;
    if (blocks == 0) blocks = 1;
    peer_sum_kernel<<<(unsigned)blocks, 256, 0, s>>>(ps, currents, pool, overwrite ? 1 : 0, kmers_out);
    return cudaGetLastError();
}

static unsigned lif_blocks(unsigned long long n) { return (unsigned)((n + LIF_THREADS - 1) / LIF_THREADS); }

cudaError_t launch_lif(const LifParams& p, cudaStream_t s) {
    if (p.pool == 0 || p.steps == 0) return cudaSuccess;  // steps == 0: :548-551 / empty loop :195
    lif_kernel<<<lif_blocks(p.pool), LIF_THREADS, 0, s>>>(p, nullptr);
    return cudaGetLastError();
}

cudaError_t launch_lif_if(const LifParams& p, const unsigned long long* only_if, cudaStream_t s) {
    if (p.pool == 0 || p.steps == 0) return cudaSuccess;
    lif_kernel<<<lif_blocks(p.pool), LIF_THREADS, 0, s>>>(p, only_if);
    return cudaGetLastError();
}

cudaError_t launch_lif_memo(const LifParams& p, const LifMemo& m, unsigned long long sat, cudaStream_t s, uint64_t* launches) {
    if (p.pool == 0 || p.steps == 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(m.keys, 0xFF, LIF_MEMO_SLOTS * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(m.ctrl, 0, 4 * sizeof(unsigned long long), s);
    if (e != cudaSuccess) return e;
    lif_memo_insert_kernel<<<lif_blocks(p.pool), LIF_THREADS, 0, s>>>(p, m, sat);
    lif_memo_compact_kernel<<<lif_blocks(LIF_MEMO_SLOTS), LIF_THREADS, 0, s>>>(m);
    lif_memo_run_kernel<<<lif_blocks(LIF_MEMO_SLOTS / 2), LIF_THREADS, 0, s>>>(p, m);
    lif_memo_apply_kernel<<<lif_blocks(p.pool), LIF_THREADS, 0, s>>>(p, m);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // too diverse a pool: the direct kernel runs instead (the counts are already folded into `currents`)
    LifParams d = p;
    d.fold_mode = 0;
    e = launch_lif_if(d, m.ctrl + 1, s);
    if (launches) *launches += 5;
    return e;
}

cudaError_t launch_lif_table_build(const LifParams& p, const LifTable& t, unsigned long long table_n, cudaStream_t s) {
    lif_table_build_kernel<<<lif_blocks(table_n), LIF_THREADS, 0, s>>>(p, t, table_n);
    return cudaGetLastError();
}

cudaError_t launch_lif_table_apply(const LifParams& p, const LifTable& t, unsigned long long table_n, cudaStream_t s) {
    if (p.pool == 0 || p.steps == 0) return cudaSuccess;
    lif_table_apply_kernel<<<lif_blocks(p.pool), LIF_THREADS, 0, s>>>(p, t, table_n);
    return cudaGetLastError();
}

cudaError_t launch_lif_single_tick(const LifParams& p, unsigned long long* currents_rw, cudaStream_t s) {
    if (p.pool == 0) return cudaSuccess;
    lif_single_tick_kernel<<<lif_blocks(p.pool), LIF_THREADS, 0, s>>>(p, currents_rw);
    return cudaGetLastError();
}

}  // namespace nk
