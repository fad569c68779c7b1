// nk_multi.cu — ONE input sharded over several GPUs inside ONE process (nk_create_multi).
//
// The reference parallelises over whole sequences inside one process (rayon fold/reduce over per-thread
// current vectors, src/spiking_hash.rs:94-154; worker threads + channels, :292-403).  Here the unit is the
// GPU: a batch is cut by WINDOW START into one contiguous range per device — sequences are cut wherever the
// range ends, the piece that owns starts [a, b) reads bytes [a, min(b + k-1, end of sequence)) — every
// device counts its pieces into its own full accumulator array, and the exchange is the sharded pool of
// nk_post.cu: device r owns neurons [r·P/N, (r+1)·P/N), its fused fold + LIF + top-N kernel sums that slice
// of EVERY device's counts through NVLink peer memory (cudaDeviceEnablePeerAccess; no IPC, no NCCL), writes
// its result pack straight into the leader device's memory, and a one-block kernel there merges the packs.
// Inside one process the ordering between devices is plain CUDA events (no spinning kernels, no timeouts).
//
// The group handle ("leader") owns no device state: group[r] are ordinary single-GPU handles.  Neuron state
// lives either in slices on the members (after a fresh-state job: the fast path above) or entirely on
// group[0] (carried state: every later job on the same counter, parameters the per-count LIF table cannot
// serve, more result rows than the slices computed).  In the second mode group[0] sums all members' counts
// out of peer memory (peer_sum_kernel) and runs the single-GPU LIF / top-N machinery on the whole pool.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>

#include "nk_internal.h"

using namespace nkd;

namespace nkd {

int exact_finalize_group(nk_counter* g);

namespace {

enum { kFresh = 0, kSliced = 1, kLeader = 2 };

struct DeviceGuard {  // entry points leave the caller's current device as they found it
    int prev = -1;
    DeviceGuard() { cudaGetDevice(&prev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// member r owns window starts [cut[r], cut[r+1]) of the concatenated batch.  Cuts are multiples of the count
// kernel's tile (4096): 16-byte aligned pointers for the in-place (zero-copy) reads of pinned batches, whole
// words of the pre-packed arrays.
void plan_cuts(uint64_t nbytes, int n, std::vector<uint64_t>& cut) {
    uint64_t per = (nbytes + (uint64_t)n - 1) / (uint64_t)n;
    per = (per + nk::COUNT_TILE - 1) / nk::COUNT_TILE * nk::COUNT_TILE;
    cut.resize((size_t)n + 1);
    for (int r = 0; r <= n; ++r) cut[(size_t)r] = std::min<uint64_t>(nbytes, per * (uint64_t)r);
}

// offsets (relative to a) of the pieces the owner of starts [a, b) counts: every sequence clipped to
// [a, b + k-1).  A piece that reaches b + k-1 has its last window start at b-1; a sequence that begins in
// [b, b + k-1) owns no start here.  Consecutive pieces are contiguous in memory (the batch is concatenated).
void shard_pieces(const uint64_t* offsets, uint64_t nseq, uint64_t a, uint64_t b, unsigned k, std::vector<uint64_t>& out) {
    out.clear();
    out.push_back(0);
    if (b <= a) return;
    const uint64_t lim = b + (uint64_t)(k - 1);
    const uint64_t* first = std::upper_bound(offsets + 1, offsets + nseq + 1, a);  // first sequence that ends after a
    for (uint64_t s = (uint64_t)(first - (offsets + 1)); s < nseq && offsets[s] < b; ++s) {
        const uint64_t p0 = std::max(offsets[s], a), p1 = std::min(offsets[s + 1], lim);
        if (p1 > p0) out.push_back(p1 - a);
    }
}

int member_after_job(nk_counter* c, bool holds_slice) {
    // accumulators are all-zero between jobs (every reader is done: ordered by the caller's events)
    NK_CUDA(cudaMemsetAsync(c->acc, 0, c->cfg.pool_size * sizeof(unsigned int), c->stream));
    if (c->spill_dirty) {
        NK_CUDA(cudaMemsetAsync(nk::dist_mail_flags(own_mail(c), 2), 0, 8, c->stream));
        c->spill_dirty = false;
    }
    c->acc_dirty = false;
    c->acc_kmers = 0;
    c->currents_valid_overwrite = false;
    c->streaming = false;
    c->top_cache_valid = false;
    if (holds_slice) {
        c->lazy_zero = false;
        c->fresh = false;
        c->slice_only = true;
        c->spike_bound += (c->cfg.steps + c->cfg.refractory) / ((unsigned long long)c->cfg.refractory + 1ull);
    }
    return NK_OK;
}

bool sliced_path_ok(const nk_counter* g) {
    const nk_counter* c0 = g->group[0];
    const int n = (int)g->group.size();
    if (g->group_state != kFresh || !g->group_can_peer || g->exact) return false;   // (exact tables: LIF on group[0])
    if (c0->force_direct || c0->cfg.steps == 0 || !std::isfinite(c0->cfg.threshold) || !std::isfinite(c0->cfg.leak)) return false;
    if (saturation_count(c0->cfg) >= (1ull << 20)) return false;
    const unsigned long long per = (c0->cfg.pool_size + n - 1) / n;
    const unsigned long long n_top = std::min<unsigned long long>(c0->topn_hint, per);
    if (n_top < 1 || n_top * n > 2048) return false;
    for (const nk_counter* c : g->group)
        if (c->dist_len == 0) return false;
    return true;
}

// every member's slice of the neuron state -> group[0] (peer copies on group[0]'s stream)
int gather_to_leader(nk_counter* g) {
    if (g->group_state != kSliced) return NK_OK;
    nk_counter* c0 = g->group[0];
    for (size_t r = 1; r < g->group.size(); ++r) {
        nk_counter* c = g->group[r];
        NK_CUDA(cudaSetDevice(c->cfg.device));
        NK_CUDA(cudaEventRecord(g->ev_posted[r], c->stream));
        NK_CUDA(cudaSetDevice(c0->cfg.device));
        NK_CUDA(cudaStreamWaitEvent(c0->stream, g->ev_posted[r], 0));
        const unsigned long long lo = c->dist_lo, len = c->dist_len;
        NK_CUDA(cudaMemcpyPeerAsync(c0->currents + lo, c0->cfg.device, c->currents + lo, c->cfg.device, len * 8, c0->stream));
        NK_CUDA(cudaMemcpyPeerAsync(c0->spikes + lo, c0->cfg.device, c->spikes + lo, c->cfg.device, len * 8, c0->stream));
        NK_CUDA(cudaMemcpyPeerAsync(c0->v + lo, c0->cfg.device, c->v + lo, c->cfg.device, len * 4, c0->stream));
        NK_CUDA(cudaMemcpyPeerAsync(c0->r + lo, c0->cfg.device, c->r + lo, c->cfg.device, len * 4, c0->stream));
        c->slice_only = false;
    }
    NK_CUDA(cudaSetDevice(c0->cfg.device));
    NK_CUDA(cudaStreamSynchronize(c0->stream));
    c0->slice_only = false;
    g->group_state = kLeader;
    return NK_OK;
}

int end_sliced(nk_counter* g, bool skip_zero) {
    NvtxRange nvtx("nk:exchange (multi-GPU group: slice reduce/LIF/top-N over peer memory + pack merge)");
    const int n = (int)g->group.size();
    nk_counter* c0 = g->group[0];
    for (int r = 0; r < n; ++r) {
        nk_counter* c = g->group[(size_t)r];
        NK_CUDA(cudaSetDevice(c->cfg.device));
        NK_CUDA(cudaEventRecord(g->ev_counted[(size_t)r], c->stream));
    }
    unsigned long long n_top = 0;
    for (int r = 0; r < n; ++r) {
        nk_counter* c = g->group[(size_t)r];
        NK_CUDA(cudaSetDevice(c->cfg.device));
        for (int q = 0; q < n; ++q)
            if (q != r) NK_CUDA(cudaStreamWaitEvent(c->stream, g->ev_counted[(size_t)q], 0));
        nk::PostParams q{};
        NK_TRY(dist_build_post(c, q, &n_top));
        q.lif.skip_zero = skip_zero ? 1 : 0;
        q.pack = g->m_gathered + (unsigned long long)r * (nk::PACK_HDR + 2 * n_top);  // group[0]'s memory
        NK_CUDA(nk::launch_post(q, c->post_grid, c->stream));
        ++c->last.launches;
        c->last.lif_path = 4;
        NK_CUDA(cudaEventRecord(g->ev_posted[(size_t)r], c->stream));
    }
    for (int r = 0; r < n; ++r) {
        nk_counter* c = g->group[(size_t)r];
        NK_CUDA(cudaSetDevice(c->cfg.device));
        for (int q = 0; q < n; ++q)
            if (q != r) NK_CUDA(cudaStreamWaitEvent(c->stream, g->ev_posted[(size_t)q], 0));
        if (r == 0) {
            const unsigned long long n_out = std::min<unsigned long long>(c0->topn_hint, c0->cfg.pool_size);
            NK_CUDA(nk::launch_merge_packs(g->m_gathered, n, 0, n_top, n_out, merged_pack_out(c0), c0->stream));
            ++c0->last.launches;
            NK_TRY(dist_finish(c0, n_out));  // queues the read-back of the merged pack
        } else {
            NK_TRY(member_after_job(c, true));
        }
    }
    g->group_state = kSliced;
    return NK_OK;
}

int end_on_leader(nk_counter* g, bool skip_zero) {
    const int n = (int)g->group.size();
    nk_counter* c0 = g->group[0];
    NK_TRY(gather_to_leader(g));
    nk::PeerSumParams ps{};
    ps.n = n;
    for (int r = 0; r < n; ++r) {
        nk_counter* c = g->group[(size_t)r];
        NK_CUDA(cudaSetDevice(c->cfg.device));
        NK_CUDA(cudaEventRecord(g->ev_counted[(size_t)r], c->stream));
        ps.acc[r] = c->acc;
        ps.spill[r] = c->spill;
        ps.kmers[r] = c->scalars + 2;
        if (c->spill_dirty) ps.spill_mask |= 1u << r;
    }
    NK_CUDA(cudaSetDevice(c0->cfg.device));
    for (int r = 1; r < n; ++r) NK_CUDA(cudaStreamWaitEvent(c0->stream, g->ev_counted[(size_t)r], 0));
    PhaseEvents pe = c0->stream_pe;
    NK_TRY(get_event(c0, &pe.fold0));
    NK_CUDA(cudaEventRecord(pe.fold0, c0->stream));
    NK_TRY(materialize_zero(c0));
    // totals of THIS call overwrite the stored currents (src/spiking_hash.rs:174-176, :463-465)
    NK_CUDA(nk::launch_peer_sum(ps, c0->currents, c0->cfg.pool_size, /*overwrite=*/true, c0->scalars + 2, c0->stream));
    ++c0->last.launches;
    NK_CUDA(cudaEventRecord(g->ev_leader, c0->stream));
    NK_CUDA(cudaMemsetAsync(c0->acc, 0, c0->cfg.pool_size * sizeof(unsigned int), c0->stream));
    if (c0->spill_dirty) {
        NK_CUDA(cudaMemsetAsync(nk::dist_mail_flags(own_mail(c0), 2), 0, 8, c0->stream));
        c0->spill_dirty = false;
    }
    c0->acc_dirty = false;
    c0->acc_kmers = 0;
    c0->currents_valid_overwrite = false;
    NK_TRY(get_event(c0, &pe.fold1));
    NK_CUDA(cudaEventRecord(pe.fold1, c0->stream));
    NK_TRY(simulate(c0, skip_zero, /*with_topn=*/true));
    NK_TRY(get_event(c0, &pe.lif1));
    NK_CUDA(cudaEventRecord(pe.lif1, c0->stream));
    NK_TRY(finish_call(c0, true, &pe));
    c0->streaming = false;
    for (int r = 1; r < n; ++r) {
        nk_counter* c = g->group[(size_t)r];
        NK_CUDA(cudaSetDevice(c->cfg.device));
        NK_CUDA(cudaStreamWaitEvent(c->stream, g->ev_leader, 0));
        NK_TRY(member_after_job(c, false));
    }
    g->group_state = kLeader;
    return NK_OK;
}

}  // namespace

int group_unsupported(const char* what) {
    return fail(NK_ERR_UNSUPPORTED, "%s is not available on a multi-GPU group handle (nk_create_multi); use a single-GPU "
                "counter, or one handle per GPU with the nk_dist_* calls", what);
}

int group_destroy(nk_counter* g) {
    for (nk_counter* c : g->group) nk_destroy(c);
    if (!g->group.empty()) cudaSetDevice(g->cfg.device);
    cudaFree(g->m_gathered);
    for (cudaEvent_t e : g->ev_counted) cudaEventDestroy(e);
    for (cudaEvent_t e : g->ev_posted) cudaEventDestroy(e);
    if (g->ev_leader) cudaEventDestroy(g->ev_leader);
    for (int i = 0; i < 2; ++i)
        if (g->file_batch[i]) cudaFreeHost(g->file_batch[i]);
    delete g;
    return NK_OK;
}

int group_reset(nk_counter* g) {
    DeviceGuard dg;
    for (nk_counter* c : g->group) {
        NK_TRY(nk_reset(c));
        c->xs.n_keys = 0;
        c->xs.valid = false;
    }
    g->group_uniques_valid = false;
    g->group_state = kFresh;
    g->group_streaming = false;
    g->group_counted = false;
    return NK_OK;
}

int group_set_steps(nk_counter* g, uint64_t steps) {
    g->cfg.steps = steps;
    for (nk_counter* c : g->group) c->cfg.steps = steps;
    return NK_OK;
}

int group_begin(nk_counter* g) {
    if (g->group_streaming) return fail(NK_ERR_STATE, "nk_stream_begin called twice");
    DeviceGuard dg;
    for (nk_counter* c : g->group) NK_TRY(nk_stream_begin(c));
    g->group_streaming = true;
    g->group_counted = false;
    return NK_OK;
}

int group_push(nk_counter* g, const uint8_t* bases, const uint32_t* codes, const uint32_t* other, const uint64_t* offsets,
               uint64_t nseq) {
    if (!g->group_streaming) return fail(NK_ERR_STATE, "nk_stream_push without nk_stream_begin");
    if (nseq == 0) return NK_OK;
    for (uint64_t s = 0; s < nseq; ++s)
        if (offsets[s + 1] < offsets[s]) return fail(NK_ERR_BAD_ARG, "offsets must be non-decreasing (at %llu)", (unsigned long long)s);
    const uint64_t nbytes = offsets[nseq];
    if (nbytes == 0) return NK_OK;
    DeviceGuard dg;
    const int n = (int)g->group.size();
    std::vector<uint64_t> cut;
    plan_cuts(nbytes, n, cut);
    g->shard_offsets.resize((size_t)n);
    int rc = NK_OK;
    std::vector<char> pushed((size_t)n, 0);
    for (int r = 0; r < n && rc == NK_OK; ++r) {
        const uint64_t a = cut[(size_t)r], b = cut[(size_t)r + 1];
        if (b <= a) continue;
        std::vector<uint64_t>& po = g->shard_offsets[(size_t)r];
        shard_pieces(offsets, nseq, a, b, g->cfg.k, po);
        if (po.size() < 2) continue;
        nk_counter* c = g->group[(size_t)r];
        if (cudaSetDevice(c->cfg.device) != cudaSuccess) { rc = fail(NK_ERR_CUDA, "cudaSetDevice(%d)", c->cfg.device); break; }
        // deferred: every device's copies and kernels are enqueued before this thread waits for any of them
        if (bases) rc = count_host_batch(c, bases + a, po.data(), po.size() - 1, &c->stream_pe, kPushDeferred);
        else rc = count_host_batch_packed(c, codes + a / 16, other ? other + a / 32 : nullptr, po.data(), po.size() - 1,
                                          &c->stream_pe, kPushDeferred);
        pushed[(size_t)r] = 1;
    }
    // the caller's buffers are reusable on return: wait for the copies, and for the kernels of the members
    // that read their range in place (pinned input)
    for (int r = 0; r < n; ++r) {
        if (!pushed[(size_t)r]) continue;
        nk_counter* c = g->group[(size_t)r];
        cudaSetDevice(c->cfg.device);
        cudaError_t e = cudaStreamSynchronize(c->copy_stream);
        if (e == cudaSuccess && c->last_push_zc) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess && rc == NK_OK) rc = fail(NK_ERR_CUDA, "multi-GPU push: %s", cudaGetErrorString(e));
    }
    g->group_counted = true;
    return rc;
}

int group_end(nk_counter* g, bool skip_zero) {
    if (!g->group_streaming) return fail(NK_ERR_STATE, "nk_stream_end without nk_stream_begin");
    DeviceGuard dg;
    g->group_streaming = false;
    NK_TRY(sliced_path_ok(g) ? end_sliced(g, skip_zero) : end_on_leader(g, skip_zero));
    if (g->exact) NK_TRY(exact_finalize_group(g));
    return NK_OK;
}

int group_process_batch(nk_counter* g, const uint8_t* bases, const uint32_t* codes, const uint32_t* other,
                        const uint64_t* offsets, uint64_t nseq) {
    if (g->group_streaming) return fail(NK_ERR_STATE, "nk_process_batch inside nk_stream_begin/end");
    NK_TRY(group_begin(g));
    int rc = group_push(g, bases, codes, other, offsets, nseq);
    if (rc != NK_OK) {
        group_reset(g);
        return rc;
    }
    return group_end(g, /*skip_zero=*/true);  // in-memory LIF driver (src/spiking_hash.rs:187-200)
}

int group_simulate(nk_counter* g) {
    if (g->group_streaming) return fail(NK_ERR_STATE, "nk_simulate inside nk_stream_begin/end");
    DeviceGuard dg;
    NK_TRY(gather_to_leader(g));
    NK_TRY(nk_simulate(g->group[0]));
    g->group_state = kLeader;
    return NK_OK;
}

int group_top_n(nk_counter* g, uint64_t top_n, nk_top_entry* out, uint64_t* n_out) {
    DeviceGuard dg;
    nk_counter* c0 = g->group[0];
    const uint64_t n = std::min<uint64_t>(top_n, g->cfg.pool_size);
    if (n <= 2048)
        for (nk_counter* c : g->group) c->topn_hint = std::max<unsigned long long>(c->topn_hint, n);
    // more rows than the slices computed: bring the state to group[0] and select there
    if (g->group_state == kSliced && !(c0->top_cache_valid && n <= c0->top_cached_n)) NK_TRY(gather_to_leader(g));
    NK_TRY(nk_top_n(c0, top_n, out, n_out));
    if (g->exact && out && n_out) {  // `uniques` = kmer_per_neuron of the MERGED table (group[0] only knows its own windows)
        if (!g->group_uniques_valid) {
            g->group_uniques.assign(g->cfg.pool_size, 0u);
            NK_TRY(group_copy_uniques(g, g->group_uniques.data()));
            g->group_uniques_valid = true;
        }
        for (uint64_t i = 0; i < *n_out; ++i) out[i].uniques = g->group_uniques[out[i].idx];
    }
    return NK_OK;
}

int group_copy(nk_counter* g, int which, void* out) {
    if (!out) return fail(NK_ERR_BAD_ARG, "null argument");
    DeviceGuard dg;
    nk_counter* c0 = g->group[0];
    auto src_of = [&](nk_counter* c) -> const unsigned char* {
        switch (which) {
            case 0: return reinterpret_cast<const unsigned char*>(c->currents);
            case 1: return reinterpret_cast<const unsigned char*>(c->spikes);
            case 2: return reinterpret_cast<const unsigned char*>(c->v);
            default: return reinterpret_cast<const unsigned char*>(c->r);
        }
    };
    const size_t esz = which <= 1 ? 8 : 4;
    if (g->group_state != kSliced) {
        NK_CUDA(cudaSetDevice(c0->cfg.device));
        return copy_out(c0, out, src_of(c0), c0->cfg.pool_size * esz);
    }
    NK_CUDA(cudaSetDevice(c0->cfg.device));
    NK_TRY(resolve(c0));
    for (nk_counter* c : g->group) {
        NK_CUDA(cudaSetDevice(c->cfg.device));
        NK_CUDA(cudaMemcpyAsync(static_cast<unsigned char*>(out) + c->dist_lo * esz, src_of(c) + c->dist_lo * esz,
                                c->dist_len * esz, cudaMemcpyDeviceToHost, c->stream));
    }
    for (nk_counter* c : g->group) {
        NK_CUDA(cudaSetDevice(c->cfg.device));
        NK_CUDA(cudaStreamSynchronize(c->stream));
    }
    return NK_OK;
}

int group_timings(nk_counter* g, nk_timings* out) {
    DeviceGuard dg;
    nk_counter* c0 = g->group[0];
    NK_CUDA(cudaSetDevice(c0->cfg.device));
    NK_TRY(resolve(c0));
    nk_timings t = c0->last;
    for (size_t r = 1; r < g->group.size(); ++r) {
        nk_counter* c = g->group[r];
        NK_CUDA(cudaSetDevice(c->cfg.device));
        NK_CUDA(cudaStreamSynchronize(c->stream));
        if (!c->stream_pe.mark0.empty() || c->stream_pe.copy0) {
            c->last.mark_ms = c->last.count_ms = 0.f;
            collect_timings(c, c->stream_pe);
            c->stream_pe = PhaseEvents{};
        }
        // the devices work side by side: a phase takes as long as its slowest member
        t.h2d_ms = std::max(t.h2d_ms, c->last.h2d_ms);
        t.mark_ms = std::max(t.mark_ms, c->last.mark_ms);
        t.count_ms = std::max(t.count_ms, c->last.count_ms);
        t.launches += c->last.launches;
        t.h2d_bytes += c->last.h2d_bytes;
        t.d2h_bytes += c->last.d2h_bytes;
    }
    *out = t;
    return NK_OK;
}

int group_synchronize(nk_counter* g) {
    DeviceGuard dg;
    for (nk_counter* c : g->group) NK_TRY(nk_synchronize(c));
    return NK_OK;
}

// ---- exact side tables of a group (src/spiking_hash.rs:157-172, 442-447, 467-473: ONE map over the whole input) ----
// GPU m has appended (word, neuron) of ITS windows.  1. every GPU builds the table of its own windows (nk_exact.cu) and
// a dense copy of it in neuron order; 2. GPU d copies, from every GPU, the records of the buckets that hold ITS neuron
// slice [d*P/n, (d+1)*P/n) (peer copies) and merges equal words: the table of its slice.  get_count asks the owner of the
// word's neuron; the table and kmer_per_neuron are the members' slices one after the other.
namespace {

template <typename F>
void on_every_member(int n, F f) {  // one host thread per GPU (the table builders synchronise their streams)
    std::vector<std::thread> th;
    for (int i = 1; i < n; ++i) th.emplace_back(f, i);
    f(0);
    for (std::thread& t : th) t.join();
}

void slice_bounds(const nk_counter* g, std::vector<unsigned long long>& b) {
    const unsigned long long n = g->group.size(), P = g->cfg.pool_size, per = (P + n - 1) / n;
    b.resize(n + 1);
    for (unsigned long long i = 0; i <= n; ++i) b[i] = std::min(P, i * per);
}

}  // namespace

int exact_finalize_group(nk_counter* g) {
    NvtxRange nvtx("nk:exact tables (multi-GPU group: per-GPU tables, then the slice merge over peer copies)");
    const int n = (int)g->group.size();
    const unsigned long long P = g->cfg.pool_size;
    std::vector<unsigned long long> bounds;
    slice_bounds(g, bounds);
    struct Export {
        nk::ExactSlot* dense = nullptr;
        std::vector<unsigned long long> first, last;
        cudaError_t e = cudaSuccess;
    };
    std::vector<Export> ex((size_t)n);
    on_every_member(n, [&](int m) {
        nk_counter* c = g->group[(size_t)m];
        Export& x = ex[(size_t)m];
        x.first.assign((size_t)n, 0);
        x.last.assign((size_t)n, 0);
        x.e = cudaSetDevice(c->cfg.device);
        if (x.e == cudaSuccess) x.e = nk::exact_finalize(c->xt, c->fm, P, nullptr, false, c->stream);
        if (x.e != cudaSuccess) return;
        const unsigned long long nk_ = c->xt.valid ? c->xt.n_keys : 0;
        if (nk_) x.e = cudaMalloc(&x.dense, nk_ * sizeof(nk::ExactSlot));
        if (x.e == cudaSuccess) x.e = nk::exact_dense_export(c->xt, x.dense, bounds.data(), n, x.first.data(), x.last.data(), c->stream);
    });
    cudaError_t bad = cudaSuccess;
    for (const Export& x : ex) if (x.e != cudaSuccess) bad = x.e;
    std::vector<cudaError_t> merged((size_t)n, cudaSuccess);
    if (bad == cudaSuccess) {
        on_every_member(n, [&](int d) {
            nk_counter* c = g->group[(size_t)d];
            cudaError_t& e = merged[(size_t)d];
            e = cudaSetDevice(c->cfg.device);
            if (e != cudaSuccess) return;
            unsigned long long total = 0;
            for (int m = 0; m < n; ++m) total += ex[(size_t)m].last[(size_t)d] - ex[(size_t)m].first[(size_t)d];
            nk::ExactSlot* cat = nullptr;
            if (total) e = cudaMalloc(&cat, total * sizeof(nk::ExactSlot));
            unsigned long long off = 0;
            for (int m = 0; m < n && e == cudaSuccess; ++m) {
                const unsigned long long a = ex[(size_t)m].first[(size_t)d], len = ex[(size_t)m].last[(size_t)d] - a;
                if (!len) continue;
                e = cudaMemcpyPeerAsync(cat + off, c->cfg.device, ex[(size_t)m].dense + a, g->group[(size_t)m]->cfg.device,
                                        len * sizeof(nk::ExactSlot), c->stream);
                off += len;
            }
            if (e == cudaSuccess)
                e = nk::exact_build_from_records(c->xs, P, cat, total, (unsigned)bounds[(size_t)d], (unsigned)bounds[(size_t)d + 1], c->stream);
            cudaStreamSynchronize(c->stream);
            cudaFree(cat);
        });
    }
    for (int m = 0; m < n; ++m) {
        cudaSetDevice(g->group[(size_t)m]->cfg.device);
        cudaFree(ex[(size_t)m].dense);
    }
    for (cudaError_t e : merged) if (e != cudaSuccess) bad = e;
    g->group_uniques_valid = false;
    if (bad == cudaErrorInvalidValue)
        return fail(NK_ERR_UNSUPPORTED, "exact counts: the words could not be partitioned into buckets that fit the on-chip tables");
    if (bad != cudaSuccess) return fail(NK_ERR_CUDA, "exact tables of the group: %s", cudaGetErrorString(bad));
    return NK_OK;
}

int group_enable_exact(nk_counter* g, int on) {
    if (g->group_streaming) return fail(NK_ERR_STATE, "nk_enable_exact_counts inside nk_stream_begin/end");
    DeviceGuard dg;
    for (nk_counter* c : g->group) NK_TRY(nk_enable_exact_counts(c, on));
    g->exact = on != 0;
    g->group_uniques_valid = false;
    return NK_OK;
}

int group_get_count(nk_counter* g, uint64_t kmer, uint32_t* count, int32_t* found) {
    if (!count || !found) return fail(NK_ERR_BAD_ARG, "null argument");
    if (!g->exact)
        return fail(NK_ERR_UNSUPPORTED, "exact k-mer table is off: call nk_enable_exact_counts(h, 1) before processing");
    DeviceGuard dg;
    const nk::FastMod& fm = g->group[0]->fm;
    const unsigned long long idx = nk::fastmod_u64(nk::siphash13_u64(kmer), fm);
    const unsigned long long n = g->group.size(), per = (g->cfg.pool_size + n - 1) / n;
    nk_counter* owner = g->group[(size_t)(idx / per)];
    return exact_lookup_in(owner, owner->xs, kmer, count, found);
}

int group_exact_table_size(nk_counter* g, uint64_t* n) {
    if (!n) return fail(NK_ERR_BAD_ARG, "null argument");
    if (!g->exact) return fail(NK_ERR_UNSUPPORTED, "exact k-mer table is off");
    *n = 0;
    for (nk_counter* c : g->group) *n += c->xs.valid ? c->xs.n_keys : 0;
    return NK_OK;
}

int group_copy_exact_table(nk_counter* g, uint64_t* keys, uint32_t* counts) {
    if (!g->exact) return fail(NK_ERR_UNSUPPORTED, "exact k-mer table is off");
    DeviceGuard dg;
    unsigned long long off = 0;
    for (nk_counter* c : g->group) {
        NK_TRY(exact_copy_table_of(c, c->xs, keys ? keys + off : nullptr, counts ? counts + off : nullptr));
        off += c->xs.valid ? c->xs.n_keys : 0;
    }
    return NK_OK;
}

int group_copy_uniques(nk_counter* g, uint32_t* out) {
    if (!out) return fail(NK_ERR_BAD_ARG, "null argument");
    if (!g->exact) return fail(NK_ERR_UNSUPPORTED, "exact k-mer table is off");
    DeviceGuard dg;
    std::vector<unsigned long long> bounds;
    slice_bounds(g, bounds);
    std::memset(out, 0, g->cfg.pool_size * sizeof(uint32_t));
    for (size_t d = 0; d < g->group.size(); ++d) {
        nk_counter* c = g->group[d];
        const unsigned long long lo = bounds[d], len = bounds[d + 1] - lo;
        if (!len || !c->xs.valid || !c->xs.uniques) continue;
        NK_CUDA(cudaSetDevice(c->cfg.device));
        NK_CUDA(cudaMemcpyAsync(out + lo, c->xs.uniques + lo, len * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
        NK_CUDA(cudaStreamSynchronize(c->stream));
    }
    return NK_OK;
}

}  // namespace nkd


extern "C" {

int nk_device_count(int32_t* n) {
    if (!n) return fail(NK_ERR_BAD_ARG, "null argument");
    int nd = 0;
    cudaError_t e = cudaGetDeviceCount(&nd);
    if (e != cudaSuccess) { cudaGetLastError(); nd = 0; }
    *n = nd;
    return NK_OK;
}

int nk_create_multi(const nk_config* cfg, const int32_t* devices, int32_t n_devices, nk_counter** out) {
    if (!cfg || !out) return fail(NK_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    if (n_devices < 1 || n_devices > nk::DIST_MAX_WORLD) return fail(NK_ERR_BAD_ARG, "n_devices must be in 1..%d", nk::DIST_MAX_WORLD);
    DeviceGuard dg;
    nk_counter* g = new nk_counter();
    g->cfg = *cfg;
    g->cfg.device = devices ? devices[0] : 0;
    auto bail = [&](int rc) {
        const std::string keep = g_err;
        if (g->group.empty()) delete g; else group_destroy(g);
        g_err = keep;
        return rc;
    };
    std::vector<nk_counter*> members;
    for (int r = 0; r < n_devices; ++r) {
        nk_config c = *cfg;
        c.device = devices ? devices[r] : r;
        nk_counter* m = nullptr;
        const int rc = nk_create(&c, &m);
        if (rc != NK_OK) {
            for (nk_counter* x : members) nk_destroy(x);
            const std::string keep = g_err;
            delete g;
            g_err = keep;
            return rc;
        }
        members.push_back(m);
    }
    g->group = members;
    // peer access between every pair of distinct devices (the slice kernels load the peers' counts, store packs)
    for (int r = 0; r < n_devices; ++r)
        for (int q = 0; q < n_devices; ++q) {
            const int dr = members[(size_t)r]->cfg.device, dq = members[(size_t)q]->cfg.device;
            if (dr == dq) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, dr, dq) != cudaSuccess || !can)
                return bail(fail(NK_ERR_UNSUPPORTED, "device %d cannot access device %d's memory (no NVLink / PCIe peer path)", dr, dq));
            if (cudaSetDevice(dr) != cudaSuccess) return bail(fail(NK_ERR_CUDA, "cudaSetDevice(%d)", dr));
            const cudaError_t e = cudaDeviceEnablePeerAccess(dq, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                return bail(fail(NK_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", dr, dq, cudaGetErrorString(e)));
            cudaGetLastError();
        }
    std::vector<void*> raw((size_t)n_devices);
    for (int r = 0; r < n_devices; ++r) raw[(size_t)r] = members[(size_t)r]->acc;
    for (int r = 0; r < n_devices; ++r) {
        const int rc = nk_dist_setup(members[(size_t)r], r, n_devices, nullptr, raw.data());
        if (rc != NK_OK) return bail(rc);
    }
    if (cudaSetDevice(members[0]->cfg.device) != cudaSuccess) return bail(fail(NK_ERR_CUDA, "cudaSetDevice"));
    if (cudaMalloc(&g->m_gathered, (size_t)n_devices * nk::PACK_MAX_U64 * sizeof(unsigned long long)) != cudaSuccess)
        return bail(fail(NK_ERR_OOM, "cudaMalloc(result packs)"));
    g->ev_counted.resize((size_t)n_devices);
    g->ev_posted.resize((size_t)n_devices);
    for (int r = 0; r < n_devices; ++r) {
        if (cudaSetDevice(members[(size_t)r]->cfg.device) != cudaSuccess ||
            cudaEventCreateWithFlags(&g->ev_counted[(size_t)r], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&g->ev_posted[(size_t)r], cudaEventDisableTiming) != cudaSuccess)
            return bail(fail(NK_ERR_CUDA, "cudaEventCreate"));
    }
    cudaSetDevice(members[0]->cfg.device);
    if (cudaEventCreateWithFlags(&g->ev_leader, cudaEventDisableTiming) != cudaSuccess) return bail(fail(NK_ERR_CUDA, "cudaEventCreate"));
    *out = g;
    return NK_OK;
}

// Host-only: the shard plan of a batch (what nk_create_multi handles do with every push), so that the
// partitioning rule can be checked without a device: member `rank` of `world` reads the bases from *start on
// and counts the pieces piece_offsets[0..*n_pieces] (relative to *start).
int nk_debug_shard(const uint64_t* offsets, uint64_t nseq, uint32_t k, int32_t world, int32_t rank, uint64_t* start,
                   uint64_t* piece_offsets, uint64_t* n_pieces) {
    if (!offsets || !start || !piece_offsets || !n_pieces) return fail(NK_ERR_BAD_ARG, "null argument");
    if (k < 1 || k > 32 || world < 1 || rank < 0 || rank >= world) return fail(NK_ERR_BAD_ARG, "bad k / world / rank");
    std::vector<uint64_t> cut, po;
    plan_cuts(offsets[nseq], world, cut);
    shard_pieces(offsets, nseq, cut[(size_t)rank], cut[(size_t)rank + 1], k, po);
    *start = cut[(size_t)rank];
    *n_pieces = po.size() - 1;
    std::memcpy(piece_offsets, po.data(), po.size() * sizeof(uint64_t));
    return NK_OK;
}

int nk_group_size(const nk_counter* h, int32_t* n) {
    if (!h || !n) return fail(NK_ERR_BAD_ARG, "null argument");
    *n = h->group.empty() ? 1 : (int32_t)h->group.size();
    return NK_OK;
}

}  // extern "C"
