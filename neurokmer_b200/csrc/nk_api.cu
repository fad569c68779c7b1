// nk_api.cu — the C ABI of libneurokmer (see include/neurokmer.h) on top of the
// sm_100a kernels.  Host-side orchestration only: staging, stream/event plumbing,
// the handle that mirrors `SpikingKmerCounter` (reference src/spiking_hash.rs:16-37).
//
// There is no CPU compute path in this file: every count, hash, LIF tick and top-N
// selection is a kernel launch; without an sm_100 device nk_create fails.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "nk_internal.h"

namespace nkd {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}


int get_event(nk_counter* h, cudaEvent_t* out) {
    if (h->ev_used == h->evpool.size()) {
        cudaEvent_t e;
        NK_CUDA(cudaEventCreate(&e));
        h->evpool.push_back(e);
    }
    *out = h->evpool[h->ev_used++];
    return NK_OK;
}

int ensure_devbuf(DevBuf& b, unsigned long long nbytes) {
    const unsigned long long need = nk::count_padded_bases(nbytes);
    if (need > b.bases_cap) {
        if (b.bases) cudaFree(b.bases);
        b.bases = nullptr;
        b.bases_cap = 0;
        NK_CUDA(cudaMalloc(&b.bases, need));
        b.bases_cap = need;
    }
    const unsigned long long words = nk::count_bitmap_words(nbytes);
    if (words > b.invalid_cap) {
        if (b.invalid) cudaFree(b.invalid);
        b.invalid = nullptr;
        b.invalid_cap = 0;
        NK_CUDA(cudaMalloc(&b.invalid, words * sizeof(unsigned int)));
        b.invalid_cap = words;
    }
    if (!b.copy_done) NK_CUDA(cudaEventCreateWithFlags(&b.copy_done, cudaEventDisableTiming));
    if (!b.compute_done) NK_CUDA(cudaEventCreateWithFlags(&b.compute_done, cudaEventDisableTiming));
    return NK_OK;
}

// only the invalid-start bitmap (+ events) of a chunk: the zero-copy paths read the bases from host memory
int ensure_bitmap_only(DevBuf& b, unsigned long long nbytes) {
    const unsigned long long words = nk::count_bitmap_words(nbytes);
    if (words > b.invalid_cap) {
        if (b.invalid) cudaFree(b.invalid);
        b.invalid = nullptr;
        b.invalid_cap = 0;
        NK_CUDA(cudaMalloc(&b.invalid, words * sizeof(unsigned int)));
        b.invalid_cap = words;
    }
    if (!b.copy_done) NK_CUDA(cudaEventCreateWithFlags(&b.copy_done, cudaEventDisableTiming));
    if (!b.compute_done) NK_CUDA(cudaEventCreateWithFlags(&b.compute_done, cudaEventDisableTiming));
    return NK_OK;
}

// device arrays of a pre-packed chunk of `nbytes` window starts (+ the invalid-start bitmap and events)
int ensure_devbuf_packed(DevBuf& b, unsigned long long nbytes, bool want_other) {
    const unsigned long long need_c = nk::count_padded_codes(nbytes);
    if (need_c > b.codes_cap) {
        if (b.codes) cudaFree(b.codes);
        b.codes = nullptr;
        b.codes_cap = 0;
        NK_CUDA(cudaMalloc(&b.codes, need_c));
        b.codes_cap = need_c;
    }
    const unsigned long long need_o = nk::count_padded_other(nbytes);
    if (want_other && need_o > b.other_cap) {
        if (b.other) cudaFree(b.other);
        b.other = nullptr;
        b.other_cap = 0;
        NK_CUDA(cudaMalloc(&b.other, need_o));
        b.other_cap = need_o;
    }
    const unsigned long long words = nk::count_bitmap_words(nbytes);
    if (words > b.invalid_cap) {
        if (b.invalid) cudaFree(b.invalid);
        b.invalid = nullptr;
        b.invalid_cap = 0;
        NK_CUDA(cudaMalloc(&b.invalid, words * sizeof(unsigned int)));
        b.invalid_cap = words;
    }
    if (!b.copy_done) NK_CUDA(cudaEventCreateWithFlags(&b.copy_done, cudaEventDisableTiming));
    if (!b.compute_done) NK_CUDA(cudaEventCreateWithFlags(&b.compute_done, cudaEventDisableTiming));
    return NK_OK;
}

int ensure_offsets(unsigned long long** p, unsigned long long* cap, unsigned long long n) {
    if (n > *cap) {
        if (*p) cudaFree(*p);
        *p = nullptr;
        *cap = 0;
        const unsigned long long want = n + n / 4 + 16;
        NK_CUDA(cudaMalloc(p, want * sizeof(unsigned long long)));
        *cap = want;
    }
    return NK_OK;
}

int materialize_zero(nk_counter* h) {
    if (!h->lazy_zero) return NK_OK;
    const unsigned long long P = h->cfg.pool_size;
    NK_CUDA(cudaMemsetAsync(h->currents, 0, P * sizeof(unsigned long long), h->stream));
    NK_CUDA(cudaMemsetAsync(h->v, 0, P * sizeof(float), h->stream));
    NK_CUDA(cudaMemsetAsync(h->r, 0, P * sizeof(unsigned int), h->stream));
    NK_CUDA(cudaMemsetAsync(h->spikes, 0, P * sizeof(unsigned long long), h->stream));
    h->lazy_zero = false;
    return NK_OK;
}

// fold acc into currents if a further `incoming` windows could overflow a u32 accumulator
int fold_now(nk_counter* h) {
    if (!h->acc_dirty) return NK_OK;
    NK_TRY(materialize_zero(h));
    NK_CUDA(nk::launch_fold(h->acc, h->currents, h->cfg.pool_size, h->currents_valid_overwrite, h->stream));
    ++h->last.launches;
    h->currents_valid_overwrite = false;
    h->acc_dirty = false;
    h->acc_kmers = 0;
    return NK_OK;
}

// the device k-mer counter of the call (scalars[2]) starts at zero; the fused post kernel leaves it zeroed
int zero_kmers(nk_counter* h) {
    if (!h->kmers_clean) NK_CUDA(cudaMemsetAsync(h->scalars + 2, 0, sizeof(unsigned long long), h->stream));
    h->kmers_clean = true;
    return NK_OK;
}

unsigned char* own_mail(nk_counter* h) {
    return reinterpret_cast<unsigned char*>(h->acc) + nk::dist_mail_offset(h->cfg.pool_size);
}

// sharded-pool handles: acc -> spill (u64, visible to the peers), and raise this rank's "spilled" flag
int spill_now(nk_counter* h) {
    if (!h->acc_dirty) return NK_OK;
    NK_CUDA(nk::launch_fold(h->acc, h->spill, h->cfg.pool_size, /*overwrite=*/!h->spill_dirty, h->stream));
    ++h->last.launches;
    if (!h->spill_dirty) NK_CUDA(cudaMemsetAsync(nk::dist_mail_flags(own_mail(h), 2), 1, 1, h->stream));  // u64 flag = 1
    h->spill_dirty = true;
    h->acc_dirty = false;
    h->acc_kmers = 0;
    return NK_OK;
}

// a stream that spilled ends on the NON-sharded path after all (all-reduce flow, plain stream_end): the
// spilled counts join `currents` like a fold would have put them there
int unspill(nk_counter* h) {
    if (!h->spill_dirty) return NK_OK;
    NK_TRY(materialize_zero(h));
    NK_CUDA(nk::launch_fold64(h->spill, h->currents, h->cfg.pool_size, h->currents_valid_overwrite, h->stream));
    ++h->last.launches;
    NK_CUDA(cudaMemsetAsync(nk::dist_mail_flags(own_mail(h), 2), 0, 8, h->stream));
    h->currents_valid_overwrite = false;
    h->spill_dirty = false;
    return NK_OK;
}

// mark + count one device-resident chunk: window starts [origin, origin+nstarts) of the
// concatenated batch, whose bytes live at `b.bases` (chunk-relative) and whose offsets
// (batch-absolute) live at d_offsets.
int prelaunch_table(nk_counter* h);

int count_chunk(nk_counter* h, DevBuf& b, const unsigned long long* d_offsets, unsigned long long seq_lo,
                unsigned long long seq_hi, unsigned long long origin, unsigned long long nstarts,
                unsigned long long max_windows, PhaseEvents* pe, bool packed) {
    if (nstarts == 0) return NK_OK;
    NvtxRange nvtx("nk:count (mark + windowing/SipHash/mod/RED)");
    if (h->acc_kmers + max_windows > h->fold_limit) NK_TRY(h->dist_world > 0 && h->streaming ? spill_now(h) : fold_now(h));
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    // phase timing covers at most kMaxTimedChunks chunks of a job (bounds the event pool of very long
    // streams; nk_timings.kmers and all results are unaffected)
    if (pe && pe->mark0.size() >= kMaxTimedChunks) pe = nullptr;
    if (pe) {
        NK_TRY(get_event(h, &e0));
        NK_TRY(get_event(h, &e2));
        NK_CUDA(cudaEventRecord(e0, h->stream));
    }
    // short-read batches (mean sequence length < 2 KiB): compact the valid window starts first (mode 3).
    // Long sequences: no invalid-start bitmap at all — every tile looks the few sequence ends that reach into
    // it up in the offsets (mode 5): no memset, no marking kernels, 1/8 B per base less HBM traffic.
    const unsigned long long nseq_here = seq_hi - seq_lo;
    const unsigned long long nseq_heur = h->nseq_hint ? h->nseq_hint : nseq_here;  // (the overlapped file path passes every record so far)
    const bool short_reads = nseq_here > 0 && nstarts / nseq_heur < 2048 && h->cfg.k > 1;
    const char* bm_env = getenv("NK_BITMAP");  // NK_BITMAP=1: the bitmap path for long sequences too (A/B runs, tests)
    const bool force_bitmap = bm_env && atoi(bm_env) != 0;
    const int mode = h->exact ? 2 : (short_reads ? 3 : (force_bitmap || nseq_here == 0 ? 0 : 5));
    // (an event record costs ~2.7 us of stream time: a phase with nothing in it shares its neighbour's event)
    e1 = e0;
    if (mode != 5) {
        NK_CUDA(nk::launch_mark_invalid(b.invalid, d_offsets, seq_lo, seq_hi, origin, nstarts, h->cfg.k,
                                        h->scalars + 2, h->stream, &h->last.launches));
        if (pe) {
            NK_TRY(get_event(h, &e1));
            NK_CUDA(cudaEventRecord(e1, h->stream));
        }
    }
    nk::CountParams p{};
    p.bases = packed ? b.codes : b.bases;
    p.other = packed && b.has_other ? b.other : nullptr;
    p.packed = packed ? 1 : 0;
    p.bases_bytes = b.bases_bytes;
    p.other_bytes = b.other_bytes;
    p.invalid = b.invalid;
    p.acc = h->acc;
    p.tile_counter = h->tile_counter;
    p.ntiles = nk::count_ntiles(nstarts);
    p.fm = h->fm;
    p.rm = nk::make_rotmul();
    p.k = h->cfg.k;
    if (h->exact) {
        NK_CUDA(nk::exact_reserve_words(h->xt, nstarts, h->stream));
        p.words = h->xt.words;
        p.widx = h->xt.widx;
        p.words_cursor = h->xt.cursor;
    }
    p.offsets = d_offsets + seq_lo;
    p.nseq = nseq_here;
    p.origin = origin;
    p.nstarts = nstarts;
    p.kmers_out = h->scalars + 2;
    NK_CUDA(nk::launch_count(p, h->cfg.use_canonical != 0, mode, h->stream));
    ++h->last.launches;
    if (pe) {
        NK_CUDA(cudaEventRecord(e2, h->stream));
        pe->mark0.push_back(e0);
        pe->count0.push_back(e1);
        pe->count1.push_back(e2);
    }
    h->acc_dirty = true;
    h->kmers_clean = false;
    h->acc_kmers += max_windows;
    // queued BEHIND the count kernel's persistent grid: the table build (side stream, 4 CTAs) gets
    // its SM slots when the first count CTAs retire, i.e. it runs in the count kernel's tail
    NK_TRY(prelaunch_table(h));
    return NK_OK;
}

int validate_batch(const nk_counter* h, const uint8_t* bases, const uint64_t* offsets, uint64_t nseq) {
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (nseq > 0 && !offsets) return fail(NK_ERR_BAD_ARG, "null offsets");
    if (nseq == 0) return NK_OK;
    if (offsets[0] != 0) return fail(NK_ERR_BAD_ARG, "offsets[0] must be 0");
    if (offsets[nseq] > 0 && !bases) return fail(NK_ERR_BAD_ARG, "null bases");
    return NK_OK;
}

int uniques_host_batch(nk_counter* h, const uint8_t* bases, const uint32_t* codes, const uint32_t* other,
                       const uint64_t* offsets, uint64_t nseq);

unsigned long long packed_chunk_bases();
int ensure_devbuf_packed(DevBuf& b, unsigned long long nbases, bool with_other);

// host batch -> chunked H2D (copy stream) overlapped with mark+count (compute stream)
int count_host_batch(nk_counter* h, const uint8_t* bases, const uint64_t* offsets, uint64_t nseq, PhaseEvents* pe,
                     int sync) {
    const bool wait_copies = sync != kPushFileDriver;
    h->last_push_zc = false;
    if (nseq == 0) return NK_OK;
    if (h->uniques_open) return uniques_host_batch(h, bases, nullptr, nullptr, offsets, nseq);  // second read of a file
    for (uint64_t s = 0; s < nseq; ++s)
        if (offsets[s + 1] < offsets[s]) return fail(NK_ERR_BAD_ARG, "offsets must be non-decreasing (at %llu)", (unsigned long long)s);
    const unsigned long long nbytes = offsets[nseq];
    if (nbytes == 0) return NK_OK;
    const int ob = h->cur_off;
    h->cur_off ^= 1;
    if (h->offsets_done[ob]) {
        // the kernels of the batch before last read this offsets buffer
        NK_CUDA(cudaStreamWaitEvent(h->copy_stream, h->offsets_done[ob], 0));
        if (nseq + 1 > h->offsets_cap2[ob]) NK_CUDA(cudaEventSynchronize(h->offsets_done[ob]));  // about to free it
    } else {
        NK_CUDA(cudaEventCreateWithFlags(&h->offsets_done[ob], cudaEventDisableTiming));
    }
    NK_TRY(ensure_offsets(&h->d_offsets2[ob], &h->offsets_cap2[ob], nseq + 1));
    unsigned long long* const d_offsets = h->d_offsets2[ob];
    if (pe && !pe->copy0) { NK_TRY(get_event(h, &pe->copy0)); NK_CUDA(cudaEventRecord(pe->copy0, h->copy_stream)); }
    NK_CUDA(cudaMemcpyAsync(d_offsets, offsets, (nseq + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, h->copy_stream));
    h->last.h2d_bytes += (nseq + 1) * sizeof(uint64_t) + nbytes;

    // Zero-copy body: when the batch lives in pinned, device-mapped host memory (nk_host_alloc, cudaHostAlloc /
    // cudaHostRegister) the count kernel's TMA bulk loads read whole tiles straight across PCIe — one launch
    // per < 2^32 starts, no staging copy, no per-chunk pipeline (the last tile's copy is clamped to the end of the
    // array, so nothing at all is staged).  PCIe binds either
    // way for ASCII (113 MB: 2.42 ms in place vs 2.46 ms staged, and a tighter spread); NK_ZEROCOPY=0 forces the
    // staged pipeline, which pageable memory always takes.
    unsigned long long zc_body = 0;
    const char* zc_env = getenv("NK_ZEROCOPY");
    cudaPointerAttributes at{};
    const bool at_ok = cudaPointerGetAttributes(&at, bases) == cudaSuccess;
    if (!at_ok) cudaGetLastError();
    // plain malloc / mmap memory (what the reference's &[Vec<u8>] is), and enough of it to be worth a thread pool
    const char* sp_env = getenv("NK_STAGE_POOL");
    const bool pool_on = wait_copies && at_ok && nbytes >= (8ull << 20) && (!sp_env || atoi(sp_env) != 0);
    const bool pageable_pool = pool_on && at.type == cudaMemoryTypeUnregistered;
    // Second form of the pool's work (default where the host has the cores for it): the threads PACK their pieces (2 bits
    // per base + `other` bits, the host packer's AVX-512 / AVX2 bodies) into their pinned slots instead of copying them,
    // and the chunks are counted by the pre-packed kernel: a host thread writes 3/8 B per base instead of 1, PCIe carries
    // 3/8 of the bytes.  With twelve workers 113 MB take 1.93 ms end to end — less than the 2.37 ms the same bytes need
    // when the count kernel reads them in place from PINNED memory (the link carries them at 47 GB/s) — so pinned batches
    // go this way too where the process has the host to itself (one GPU in the box).  It pays from ten workers on (a
    // process that shares the host with seven others keeps the plain copy / the in-place read).  NK_STAGE_PACK=0 / 1
    // forces either.
    const char* pk_env = getenv("NK_STAGE_PACK");
    const bool pack_stage = pool_on && (at.type == cudaMemoryTypeUnregistered || at.type == cudaMemoryTypeHost) &&
                            (pk_env ? atoi(pk_env) != 0 : stage_pack_worthwhile(h, at.type == cudaMemoryTypeHost));
    // (the file driver double-buffers its own pinned batches and must not block on the kernels: wait_copies == false)
    if (wait_copies && !pack_stage && (!zc_env || atoi(zc_env) != 0) && nbytes >= 4 * (unsigned long long)nk::COUNT_TILE &&
        ((uintptr_t)bases & 15) == 0) {
        if (at_ok && at.type == cudaMemoryTypeHost && at.devicePointer) {
            zc_body = nbytes;  // the whole batch: the last tile's copy is clamped to the end of the array
            const unsigned long long slice = (0xFFFFFFFFull / nk::COUNT_TILE - 1) * nk::COUNT_TILE;
            NK_TRY(ensure_bitmap_only(h->zc, std::min(zc_body, slice)));
            NK_CUDA(cudaEventRecord(h->zc.copy_done, h->copy_stream));  // the offsets are on the device
            NK_CUDA(cudaStreamWaitEvent(h->stream, h->zc.copy_done, 0));
            for (unsigned long long c0 = 0; c0 < zc_body; c0 += slice) {  // < 2^32 window starts per launch
                const unsigned long long n = std::min(slice, zc_body - c0);
                DevBuf view = h->zc;
                view.bases = static_cast<unsigned char*>(at.devicePointer) + c0;
                view.bases_bytes = (nbytes - c0 + 15) / 16 * 16;
                NK_TRY(count_chunk(h, view, d_offsets, 0, nseq, c0, n, n, pe, false));
            }
        } else {
            cudaGetLastError();
        }
    }

    // chunk plan: fixed 32 MiB granules.  Measured alternatives at 113 MB (end to end, B200, PCIe Gen5):
    // fixed 32 MiB 2.53 ms; geometric tail down to 2 MiB 2.68 ms; 3 x 36 MiB + 4 MiB tail 2.60 ms —
    // fewer, equal copies win over a shorter exposed tail.
    // Pinned sources that are not read in place (NK_ZEROCOPY=0, unaligned pointers): DMA copies in 16 MiB chunks, one
    // count launch per chunk.  Measured end to end at 113 MB (tools/h2d_sweep.py, profiles/r02_bench.md): in place 2.35 ms;
    // chunks of 2 / 4 / 8 / 16 / 32 MiB 3.48 / 2.80 / 2.42 / 2.33 / 2.40 ms — below 8 MiB the host's per-chunk
    // submission cost (copy, events, launch) shows, at 32 MiB the last chunk's kernel does.
    unsigned long long chunk_bytes = kChunkBytes;
    if (at_ok && at.type == cudaMemoryTypeHost) {
        unsigned long long mb = 16;
        if (const char* e = getenv("NK_H2D_CHUNK_MB")) { const unsigned long long t = strtoull(e, nullptr, 10); if (t >= 1 && t <= 32) mb = t; }
        chunk_bytes = mb << 20;
    }
    // experiment knob: NK_H2D_PLAN="32,32,24,16,8,2" = chunk sizes in MiB for a pinned source, the last one repeats
    std::vector<unsigned long long> plan;
    if (at_ok && at.type == cudaMemoryTypeHost) {
        if (const char* e = getenv("NK_H2D_PLAN")) {
            for (const char* q = e; *q;) {
                char* end = nullptr;
                const unsigned long long t = strtoull(q, &end, 10);
                if (end == q) break;
                if (t >= 1 && t <= 32) plan.push_back(t << 20);
                q = *end ? end + 1 : end;
            }
        }
    }
    size_t plan_i = 0;
    if (pack_stage) {
        chunk_bytes = packed_chunk_bases();
        h->last.h2d_bytes -= nbytes;   // (counted above as ASCII)
    }
    for (unsigned long long c0 = zc_body, c1 = 0; c0 < nbytes; c0 = c1) {
        if (!plan.empty()) { chunk_bytes = plan[std::min(plan_i, plan.size() - 1)]; ++plan_i; }
        c1 = std::min(c0 + chunk_bytes, nbytes);
        const unsigned long long copy_len = std::min(c1 + nk::COUNT_HALO, nbytes) - c0;
        DevBuf& b = h->buf[h->cur_buf];
        h->cur_buf ^= 1;
        const bool had = b.compute_done != nullptr;
        if (pack_stage) {
            // bases of this chunk plus the halo its last tiles read (64 bases of codes, 128 of `other` bits)
            const unsigned long long n_pack = std::min(c1 + 128, nbytes) - c0;
            NK_TRY(ensure_devbuf_packed(b, std::min(chunk_bytes, nbytes), true));
            NK_TRY(stage_pack_to_device(h, bases + c0, n_pack, b.codes, b.other, had ? b.compute_done : nullptr, h->copy_stream));
            b.has_other = true;
            h->last.h2d_bytes += (n_pack + 15) / 16 * 4 + (n_pack + 31) / 32 * 4;
            NK_CUDA(cudaEventRecord(b.copy_done, h->copy_stream));
            NK_CUDA(cudaStreamWaitEvent(h->stream, b.copy_done, 0));
            const uint64_t* first = std::upper_bound(offsets + 1, offsets + nseq + 1, (uint64_t)c0);
            const unsigned long long seq_lo = (unsigned long long)(first - (offsets + 1));
            const uint64_t* last = std::lower_bound(offsets, offsets + nseq, (uint64_t)c1);
            const unsigned long long seq_hi = (unsigned long long)(last - offsets);
            NK_TRY(count_chunk(h, b, d_offsets, seq_lo, seq_hi, c0, c1 - c0, c1 - c0, pe, true));
            NK_CUDA(cudaEventRecord(b.compute_done, h->stream));
            continue;
        }
        NK_TRY(ensure_devbuf(b, std::min(kChunkBytes, nbytes)));
        if (pageable_pool) {
            // pageable memory: the pool of host threads copies 2 MiB pieces into its pinned slots and issues their H2D
            // copies on its own streams (one cudaMemcpy from pageable memory runs at ~11 GB/s); returns once the source
            // bytes of this chunk have been consumed, with the copies in flight behind the previous user of `b`
            NK_TRY(stage_to_device(h, bases + c0, -1, 0, copy_len, b.bases, had ? b.compute_done : nullptr, h->copy_stream));
        } else {
            if (had) NK_CUDA(cudaStreamWaitEvent(h->copy_stream, b.compute_done, 0));
            NK_CUDA(cudaMemcpyAsync(b.bases, bases + c0, copy_len, cudaMemcpyHostToDevice, h->copy_stream));
        }
        NK_CUDA(cudaEventRecord(b.copy_done, h->copy_stream));
        NK_CUDA(cudaStreamWaitEvent(h->stream, b.copy_done, 0));
        // sequences that overlap [c0, c1): first with end > c0 ... first with start >= c1
        const uint64_t* first = std::upper_bound(offsets + 1, offsets + nseq + 1, (uint64_t)c0);
        const unsigned long long seq_lo = (unsigned long long)(first - (offsets + 1));
        const uint64_t* last = std::lower_bound(offsets, offsets + nseq, (uint64_t)c1);
        const unsigned long long seq_hi = (unsigned long long)(last - offsets);
        NK_TRY(count_chunk(h, b, d_offsets, seq_lo, seq_hi, c0, c1 - c0, c1 - c0, pe, false));
        NK_CUDA(cudaEventRecord(b.compute_done, h->stream));
    }
    NK_CUDA(cudaEventRecord(h->offsets_done[ob], h->stream));
    if (pe) {
        if (!pe->copy1) NK_TRY(get_event(h, &pe->copy1));
        NK_CUDA(cudaEventRecord(pe->copy1, h->copy_stream));
    }
    // the caller's buffers must be reusable on return: wait for the copies (not the kernels).
    // The file driver owns its pinned batches and double-buffers them instead (wait_copies = false).
    h->last_push_zc = zc_body != 0;
    if (sync == kPushSync) {
        NK_CUDA(cudaStreamSynchronize(h->copy_stream));
        if (zc_body) NK_CUDA(cudaStreamSynchronize(h->stream));  // the kernel itself read the caller's buffer
    }
    return NK_OK;
}

// pre-packed host batch -> chunked H2D of the code words (and `other` bits) overlapped with mark+count.
// Same pipeline as count_host_batch with 1/4 (+1/8) of the bytes on PCIe.
unsigned long long packed_chunk_bases() {
    unsigned long long mb = 32;  // B200 / PCIe Gen5, 113 Mbase: 16 -> 1.55 ms, 32 -> 1.44 ms, 64 -> 1.53 ms (profiles/r01_bench.md)
    if (const char* e = getenv("NK_PACKED_CHUNK_MBASES")) {
        const unsigned long long t = strtoull(e, nullptr, 10);
        if (t >= 1 && t <= 1024) mb = t;
    }
    return mb << 20;  // multiple of COUNT_TILE, of 16 (code words) and of 32 (`other` words)
}

int count_host_batch_packed(nk_counter* h, const uint32_t* codes, const uint32_t* other, const uint64_t* offsets,
                            uint64_t nseq, PhaseEvents* pe, int sync) {
    h->last_push_zc = false;
    if (nseq == 0) return NK_OK;
    for (uint64_t s = 0; s < nseq; ++s)
        if (offsets[s + 1] < offsets[s]) return fail(NK_ERR_BAD_ARG, "offsets must be non-decreasing (at %llu)", (unsigned long long)s);
    const unsigned long long nbases = offsets[nseq];
    if (nbases == 0) return NK_OK;
    if (!codes) return fail(NK_ERR_BAD_ARG, "null codes");
    const int ob = h->cur_off;
    h->cur_off ^= 1;
    if (h->offsets_done[ob]) {
        NK_CUDA(cudaStreamWaitEvent(h->copy_stream, h->offsets_done[ob], 0));
        if (nseq + 1 > h->offsets_cap2[ob]) NK_CUDA(cudaEventSynchronize(h->offsets_done[ob]));
    } else {
        NK_CUDA(cudaEventCreateWithFlags(&h->offsets_done[ob], cudaEventDisableTiming));
    }
    NK_TRY(ensure_offsets(&h->d_offsets2[ob], &h->offsets_cap2[ob], nseq + 1));
    unsigned long long* const d_offsets = h->d_offsets2[ob];
    if (pe && !pe->copy0) { NK_TRY(get_event(h, &pe->copy0)); NK_CUDA(cudaEventRecord(pe->copy0, h->copy_stream)); }
    NK_CUDA(cudaMemcpyAsync(d_offsets, offsets, (nseq + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, h->copy_stream));
    h->last.h2d_bytes += (nseq + 1) * sizeof(uint64_t);

    const unsigned long long chunk = packed_chunk_bases();
    const unsigned char* const hc = reinterpret_cast<const unsigned char*>(codes);
    const unsigned char* const ho = reinterpret_cast<const unsigned char*>(other);

    // Zero-copy body: when the packed arrays live in pinned, device-mapped host memory (nk_host_alloc,
    // cudaHostAlloc/Register) the count kernel's TMA bulk loads read the tiles straight across PCIe —
    // one launch, no staging copy, no per-chunk pipeline bubbles (3/8 B per base keeps PCIe at the
    // kernel's own pace: 113 Mbase end to end 1.44 ms staged -> 1.11 ms, profiles/r01_bench.md).  The last tile's copies
    // are clamped to the ends of the host arrays, so the whole batch is read this way.  NK_ZEROCOPY=0
    // forces the staged path.
    unsigned long long zc_body = 0;
    const char* zc_env = getenv("NK_ZEROCOPY");
    const bool zc_on = !zc_env || atoi(zc_env) != 0;
    if (zc_on && nbases >= 4ull * nk::COUNT_TILE + 128 && ((uintptr_t)codes & 15) == 0 && ((uintptr_t)other & 15) == 0) {
        cudaPointerAttributes ac{}, ao{};
        bool ok = cudaPointerGetAttributes(&ac, codes) == cudaSuccess && ac.type == cudaMemoryTypeHost && ac.devicePointer;
        if (ok && other) ok = cudaPointerGetAttributes(&ao, other) == cudaSuccess && ao.type == cudaMemoryTypeHost && ao.devicePointer;
        if (ok) {
            zc_body = nbases;  // the whole batch: the last tile's copies are clamped to the ends of the arrays
            const unsigned long long slice = (0xFFFFFFFFull / nk::COUNT_TILE - 1) * nk::COUNT_TILE;
            NK_TRY(ensure_bitmap_only(h->zc, std::min(zc_body, slice)));
            NK_CUDA(cudaEventRecord(h->zc.copy_done, h->copy_stream));  // the offsets are on the device
            NK_CUDA(cudaStreamWaitEvent(h->stream, h->zc.copy_done, 0));
            for (unsigned long long c0 = 0; c0 < zc_body; c0 += slice) {
                const unsigned long long n = std::min(slice, zc_body - c0);
                DevBuf view = h->zc;
                view.codes = static_cast<unsigned char*>(ac.devicePointer) + c0 / 4;
                view.other = other ? static_cast<unsigned char*>(ao.devicePointer) + c0 / 8 : nullptr;
                view.has_other = other != nullptr;
                view.bases_bytes = ((nbases + 15) / 16 * 4 - c0 / 4 + 15) / 16 * 16;
                view.other_bytes = ((nbases + 31) / 32 * 4 - c0 / 8 + 15) / 16 * 16;
                NK_TRY(count_chunk(h, view, d_offsets, 0, nseq, c0, n, n, pe, true));
            }
            h->last.h2d_bytes += (nbases + 15) / 16 * 4 + (other ? (nbases + 31) / 32 * 4 : 0);
        } else {
            cudaGetLastError();  // pageable memory: not an error, take the staged path
        }
    }

    for (unsigned long long c0 = zc_body, c1 = 0; c0 < nbases; c0 = c1) {
        c1 = std::min(c0 + chunk, nbases);
        // whole words of the host arrays, including the halo the tiles of this chunk read past c1
        const unsigned long long code_bytes = (std::min(c1 + 64, nbases) - c0 + 15) / 16 * 4;
        const unsigned long long other_bytes = (std::min(c1 + 128, nbases) - c0 + 31) / 32 * 4;
        DevBuf& b = h->buf[h->cur_buf];
        h->cur_buf ^= 1;
        const bool had = b.compute_done != nullptr;
        NK_TRY(ensure_devbuf_packed(b, std::min(chunk, nbases), other != nullptr));
        if (had) NK_CUDA(cudaStreamWaitEvent(h->copy_stream, b.compute_done, 0));
        NK_CUDA(cudaMemcpyAsync(b.codes, hc + c0 / 4, code_bytes, cudaMemcpyHostToDevice, h->copy_stream));
        h->last.h2d_bytes += code_bytes;
        b.has_other = other != nullptr;
        if (other) {
            NK_CUDA(cudaMemcpyAsync(b.other, ho + c0 / 8, other_bytes, cudaMemcpyHostToDevice, h->copy_stream));
            h->last.h2d_bytes += other_bytes;
        }
        NK_CUDA(cudaEventRecord(b.copy_done, h->copy_stream));
        NK_CUDA(cudaStreamWaitEvent(h->stream, b.copy_done, 0));
        const uint64_t* first = std::upper_bound(offsets + 1, offsets + nseq + 1, (uint64_t)c0);
        const unsigned long long seq_lo = (unsigned long long)(first - (offsets + 1));
        const uint64_t* last = std::lower_bound(offsets, offsets + nseq, (uint64_t)c1);
        const unsigned long long seq_hi = (unsigned long long)(last - offsets);
        NK_TRY(count_chunk(h, b, d_offsets, seq_lo, seq_hi, c0, c1 - c0, c1 - c0, pe, true));
        NK_CUDA(cudaEventRecord(b.compute_done, h->stream));
    }
    NK_CUDA(cudaEventRecord(h->offsets_done[ob], h->stream));
    if (pe) {
        if (!pe->copy1) NK_TRY(get_event(h, &pe->copy1));
        NK_CUDA(cudaEventRecord(pe->copy1, h->copy_stream));
    }
    h->last_push_zc = zc_body != 0;
    if (sync == kPushSync) {
        NK_CUDA(cudaStreamSynchronize(h->copy_stream));
        if (zc_body) NK_CUDA(cudaStreamSynchronize(h->stream));  // the kernel itself read the caller's arrays
    }
    return NK_OK;
}

// ---- uniques pass (second pass over the input) -------------------------------------------------------
// One device-resident chunk: mark the invalid starts, then run the count kernel in mode 4 (no pool update;
// words whose neuron is in the filter are appended).  If the word array was too small the cursor says by
// how much: grow and run the chunk again (the appended prefix of the failed run is simply overwritten).
int uniques_chunk(nk_counter* h, DevBuf& b, const unsigned long long* d_offsets, unsigned long long seq_lo,
                  unsigned long long seq_hi, unsigned long long origin, unsigned long long nstarts, bool packed) {
    if (nstarts == 0) return NK_OK;
    NK_CUDA(nk::launch_mark_invalid(b.invalid, d_offsets, seq_lo, seq_hi, origin, nstarts, h->cfg.k, h->scalars + 3,
                                    h->stream, nullptr));
    for (int attempt = 0; attempt < 3; ++attempt) {
        nk::CountParams p{};
        p.bases = packed ? b.codes : b.bases;
        p.other = packed && b.has_other ? b.other : nullptr;
        p.packed = packed ? 1 : 0;
        p.invalid = b.invalid;
        p.acc = h->acc;
        p.tile_counter = h->tile_counter;
        p.ntiles = nk::count_ntiles(nstarts);
        p.fm = h->fm;
        p.rm = nk::make_rotmul();
        p.k = h->cfg.k;
        p.filter = h->d_filter;
        p.words = h->ut.words;
        p.widx = h->ut.widx;
        p.words_cursor = h->ut.cursor;
        p.words_cap = h->ut.words_cap;
        NK_CUDA(nk::launch_count(p, h->cfg.use_canonical != 0, 4, h->stream));
        unsigned long long cursor = 0;
        NK_CUDA(cudaMemcpyAsync(&cursor, h->ut.cursor, sizeof cursor, cudaMemcpyDeviceToHost, h->stream));
        NK_CUDA(cudaStreamSynchronize(h->stream));
        if (cursor <= h->ut.words_cap) {
            h->ut_count = cursor;
            return NK_OK;
        }
        if (cursor > 0x7FFFFFF0ull)
            return fail(NK_ERR_UNSUPPORTED, "uniques pass: more than 2^31 windows map to the requested rows "
                        "(tiny pool?); use nk_enable_exact_counts on smaller batches");
        // too small: keep what earlier chunks appended, rewind the cursor, grow, run the chunk again
        NK_CUDA(nk::exact_grow_words(h->ut, cursor + cursor / 4 + 4096, h->ut_count, h->stream));
        NK_CUDA(cudaMemcpyAsync(h->ut.cursor, &h->ut_count, sizeof(unsigned long long), cudaMemcpyHostToDevice, h->stream));
        NK_CUDA(cudaStreamSynchronize(h->stream));
    }
    return fail(NK_ERR_CUDA, "uniques pass: the word array kept overflowing");
}

// a host batch (ASCII when bases != null, else packed) through the uniques pass, synchronously
int uniques_host_batch(nk_counter* h, const uint8_t* bases, const uint32_t* codes, const uint32_t* other,
                       const uint64_t* offsets, uint64_t nseq) {
    if (nseq == 0) return NK_OK;
    for (uint64_t s = 0; s < nseq; ++s)
        if (offsets[s + 1] < offsets[s]) return fail(NK_ERR_BAD_ARG, "offsets must be non-decreasing (at %llu)", (unsigned long long)s);
    const unsigned long long n = offsets[nseq];
    if (n == 0) return NK_OK;
    const bool packed = bases == nullptr;
    NK_CUDA(cudaStreamSynchronize(h->stream));
    NK_CUDA(cudaStreamSynchronize(h->copy_stream));
    NK_TRY(ensure_offsets(&h->d_offsets2[0], &h->offsets_cap2[0], nseq + 1));
    NK_CUDA(cudaMemcpyAsync(h->d_offsets2[0], offsets, (nseq + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    const unsigned long long chunk = packed ? packed_chunk_bases() : kChunkBytes;
    DevBuf& b = h->buf[0];
    for (unsigned long long c0 = 0, c1 = 0; c0 < n; c0 = c1) {
        c1 = std::min(c0 + chunk, n);
        if (packed) {
            NK_TRY(ensure_devbuf_packed(b, std::min(chunk, n), other != nullptr));
            const unsigned long long code_bytes = (std::min(c1 + 64, n) - c0 + 15) / 16 * 4;
            const unsigned long long other_bytes = (std::min(c1 + 128, n) - c0 + 31) / 32 * 4;
            NK_CUDA(cudaMemcpyAsync(b.codes, reinterpret_cast<const unsigned char*>(codes) + c0 / 4, code_bytes,
                                    cudaMemcpyHostToDevice, h->stream));
            b.has_other = other != nullptr;
            if (other)
                NK_CUDA(cudaMemcpyAsync(b.other, reinterpret_cast<const unsigned char*>(other) + c0 / 8, other_bytes,
                                        cudaMemcpyHostToDevice, h->stream));
        } else {
            NK_TRY(ensure_devbuf(b, std::min((unsigned long long)kChunkBytes, n)));
            const unsigned long long copy_len = std::min(c1 + nk::COUNT_HALO, n) - c0;
            NK_CUDA(cudaMemcpyAsync(b.bases, bases + c0, copy_len, cudaMemcpyHostToDevice, h->stream));
        }
        const uint64_t* first = std::upper_bound(offsets + 1, offsets + nseq + 1, (uint64_t)c0);
        const unsigned long long seq_lo = (unsigned long long)(first - (offsets + 1));
        const uint64_t* last = std::lower_bound(offsets, offsets + nseq, (uint64_t)c1);
        const unsigned long long seq_hi = (unsigned long long)(last - offsets);
        NK_TRY(uniques_chunk(h, b, h->d_offsets2[0], seq_lo, seq_hi, c0, c1 - c0, packed));  // synchronises
    }
    return NK_OK;
}

unsigned long long saturation_count(const nk_config& c) {
    // smallest count whose input current f32(f64(count)/f64(steps)) reaches the threshold
    if (!(c.threshold == c.threshold)) return ~0ull;
    auto cur = [&](unsigned long long n) { return (float)((double)n / (double)c.steps); };
    if (cur(0) >= c.threshold) return 0;
    unsigned long long lo = 0, hi = 1ull << 62;  // cur(lo) < thr
    if (!(cur(hi) >= c.threshold)) return ~0ull;
    while (hi - lo > 1) {
        const unsigned long long mid = lo + (hi - lo) / 2;
        if (cur(mid) >= c.threshold) hi = mid; else lo = mid;
    }
    return hi;
}

// (Re)build the per-count LIF table if the LIF parameters changed; `on` = stream to build on.
int ensure_table(nk_counter* h, unsigned long long table_n, cudaStream_t on) {
    const nk_config& a = h->cfg; const nk_config& b = h->table_cfg;
    const bool same = h->table_valid && h->table_n == table_n && a.steps == b.steps && a.threshold == b.threshold &&
                      a.leak == b.leak && a.refractory == b.refractory;
    if (same) return NK_OK;
    if (table_n > h->table_cap) {
        NK_CUDA(cudaStreamSynchronize(h->stream));
        NK_CUDA(cudaStreamSynchronize(h->copy_stream));
        cudaFree(h->table.spikes); cudaFree(h->table.v); cudaFree(h->table.r);
        h->table = nk::LifTable{};
        h->table_cap = 0;
        NK_CUDA(cudaMalloc(&h->table.spikes, table_n * sizeof(unsigned int)));
        NK_CUDA(cudaMalloc(&h->table.v, table_n * sizeof(float)));
        NK_CUDA(cudaMalloc(&h->table.r, table_n * sizeof(unsigned int)));
        h->table_cap = table_n;
    }
    nk::LifParams p{};
    p.steps = h->cfg.steps; p.thr = h->cfg.threshold; p.leak = h->cfg.leak; p.period = h->cfg.refractory;
    NK_CUDA(nk::launch_lif_table_build(p, h->table, table_n, on));
    ++h->last.launches;
    h->table_valid = true;
    h->table_cfg = h->cfg;
    h->table_n = table_n;
    if (on != h->stream) {
        if (!h->table_ready) NK_CUDA(cudaEventCreateWithFlags(&h->table_ready, cudaEventDisableTiming));
        NK_CUDA(cudaEventRecord(h->table_ready, on));
        h->table_inflight = true;
    }
    return NK_OK;
}

bool table_path_ok(const nk_counter* h, unsigned long long* table_n) {
    if (!h->fresh || h->force_direct != 0 || h->cfg.steps == 0 || !std::isfinite(h->cfg.threshold) || !std::isfinite(h->cfg.leak)) return false;
    const unsigned long long sat = saturation_count(h->cfg);
    if (sat >= (1ull << 20)) return false;
    *table_n = sat + 1;
    return true;
}

// At the start of a job on a fresh counter: the table depends on the LIF parameters only, so it
// is built on the side stream while the count kernel runs (the post kernel waits for its event).
int prelaunch_table(nk_counter* h) {
    unsigned long long table_n = 0;
    if (!table_path_ok(h, &table_n)) return NK_OK;
    return ensure_table(h, table_n, h->copy_stream);
}

// LIF over this call's totals; skip_zero: in-memory driver (:187-200) vs SIMD driver (:544-659).
// If the u32 batch accumulators still hold counts they are folded into `currents` by the LIF
// kernel itself (fold_mode), otherwise the stored currents are used as they are.
int simulate(nk_counter* h, bool skip_zero, bool with_topn) {
    NvtxRange nvtx("nk:post (fold + LIF + top-N)");
    h->last.lif_path = 0;
    h->top_cache_valid = false;
    int fold_mode = 0;
    if (h->acc_dirty) fold_mode = h->currents_valid_overwrite ? 2 : 1;
    if (h->cfg.steps == 0 || h->cfg.pool_size == 0) return fold_now(h);
    nk::LifParams p{};
    p.currents = h->currents;
    p.acc = h->acc;
    p.fold_mode = fold_mode;
    p.v = h->v;
    p.r = h->r;
    p.spikes = h->spikes;
    p.total_new = h->scalars + 0;
    p.max_spikes = h->scalars + 1;
    p.pool = h->cfg.pool_size;
    p.steps = h->cfg.steps;
    p.thr = h->cfg.threshold;
    p.leak = h->cfg.leak;
    p.period = h->cfg.refractory;
    p.skip_zero = skip_zero ? 1 : 0;
    if (!h->fired_clean) NK_CUDA(cudaMemsetAsync(h->scalars, 0, sizeof(unsigned long long), h->stream));
    h->fired_clean = false;
    unsigned long long table_n = 0;
    const bool use_table = table_path_ok(h, &table_n);
    if (h->lazy_zero) {
        if (use_table && fold_mode != 0) {
            p.zero_state = 1;   // the kernel writes currents, v, r, spikes of EVERY neuron without reading them
            p.fold_mode = 2;
        } else {
            NK_TRY(materialize_zero(h));
        }
    }
    if (use_table) {
        NK_TRY(ensure_table(h, table_n, h->stream));  // usually already built by prelaunch_table
        if (h->table_inflight) {
            NK_CUDA(cudaStreamWaitEvent(h->stream, h->table_ready, 0));
            h->table_inflight = false;
        }
        const unsigned long long per_call_b = (h->cfg.steps + h->cfg.refractory) / ((unsigned long long)h->cfg.refractory + 1ull);
        const unsigned long long bound = (h->spike_bound + per_call_b < h->spike_bound) ? ~0ull : h->spike_bound + per_call_b;
        const unsigned long long n_top = std::min<unsigned long long>(h->topn_hint, h->cfg.pool_size);
        if (with_topn && !h->exact && n_top >= 1 && n_top <= 2048) {
            // everything after the count kernel in ONE cooperative launch (nk_post.cu)
            nk::PostParams q{};
            q.lif = p;
            q.table = h->table;
            q.table_n = table_n;
            q.n = n_top;
            int bits = 0;
            while (bits < 64 && (bound >> bits)) ++bits;
            q.passes = std::max(1, (bits + 7) / 8);
            q.single_pass = bound < (unsigned long long)nk::POST_EXACT_BINS ? 1 : 0;
            q.ctrl = h->post_zero;
            q.hist = reinterpret_cast<unsigned int*>(h->post_zero + 8);
            q.block_ties = q.hist + 8 * 256;
            q.seg_counts = h->topn.block_counts;
            q.out_idx = h->topn.out_idx;
            q.out_spikes = h->topn.out_spikes;
            // single GPU: the last block writes the result pack (<= 0.8 KB) straight into the pinned, device-mapped
            // host buffer — no D2H copy operation behind the kernel (~4 us of stream time)
            q.pack = h->h_pack_dev ? h->h_pack_dev : h->d_pack;
            h->pack_direct = h->h_pack_dev != nullptr;
            q.kmers = h->scalars + 2;
            q.trace = getenv("NK_POST_TRACE") ? 1 : 0;
            // (no memsets: the kernel leaves its scratch, the spike counter and the k-mer counter zeroed)
            NK_CUDA(nk::launch_post(q, h->post_grid, h->stream));
            h->fired_clean = h->kmers_clean = true;
            ++h->last.launches;
            h->last.lif_path = 3;
            h->top_cached_n = n_top;
            h->top_cache_valid = true;
            h->pending_pack = true;
        } else {
            NK_CUDA(nk::launch_lif_table_apply(p, h->table, table_n, h->stream));
            ++h->last.launches;
            h->last.lif_path = 2;
        }
    } else {
        // carried (or forced non-table) state: one simulation per distinct (state, count) key when the parameters
        // allow the saturation clamp (leak >= 0), else one per neuron
        const unsigned long long sat = saturation_count(h->cfg);
        const bool memo_ok = h->force_direct != 1 && std::isfinite(h->cfg.threshold) && std::isfinite(h->cfg.leak) &&
                             h->cfg.leak >= 0.0f && !std::signbit(h->cfg.leak) && sat < (1ull << 21) - 1 &&
                             (h->force_direct == 2 || h->cfg.pool_size >= (1ull << 18));  // small pools: direct is as fast
        if (memo_ok) {
            if (!h->memo.keys) {
                NK_CUDA(cudaMalloc(&h->memo.keys, nk::LIF_MEMO_SLOTS * sizeof(unsigned long long)));
                NK_CUDA(cudaMalloc(&h->memo.dense, nk::LIF_MEMO_SLOTS / 2 * sizeof(unsigned int)));
                NK_CUDA(cudaMalloc(&h->memo.res_v, nk::LIF_MEMO_SLOTS * sizeof(float)));
                NK_CUDA(cudaMalloc(&h->memo.res_r, nk::LIF_MEMO_SLOTS * sizeof(unsigned int)));
                NK_CUDA(cudaMalloc(&h->memo.res_f, nk::LIF_MEMO_SLOTS * sizeof(unsigned int)));
                NK_CUDA(cudaMalloc(&h->memo.slot_of, h->cfg.pool_size * sizeof(unsigned int)));
                NK_CUDA(cudaMalloc(&h->memo.ctrl, 4 * sizeof(unsigned long long)));
            }
            NK_CUDA(nk::launch_lif_memo(p, h->memo, sat, h->stream, &h->last.launches));
            h->last.lif_path = 5;
        } else {
            NK_CUDA(nk::launch_lif(p, h->stream));
            ++h->last.launches;
            h->last.lif_path = 1;
        }
    }
    if (p.zero_state) h->lazy_zero = false;
    if (fold_mode) {
        h->acc_dirty = false;
        h->acc_kmers = 0;
        h->currents_valid_overwrite = false;
    }
    h->fresh = false;
    const unsigned long long per_call = (h->cfg.steps + h->cfg.refractory) / ((unsigned long long)h->cfg.refractory + 1ull);
    h->spike_bound = (h->spike_bound + per_call < h->spike_bound) ? ~0ull : h->spike_bound + per_call;
    return NK_OK;
}

// `(cost * 1000.0) as u64` — src/models.rs:162, src/spiking_hash.rs:649.  Rust's float -> integer `as` cast
// truncates toward zero and saturates (NaN -> 0, negative -> 0, >= 2^64 -> u64::MAX); the same cast is
// undefined behaviour in C++ outside [0, 2^64), so the clamp is explicit.
unsigned long long cost_fixed(double spike_cost) {
    const double x = spike_cost * 1000.0;
    if (!(x > 0.0)) return 0ull;
    if (x >= 18446744073709551616.0) return ~0ull;
    return (unsigned long long)x;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    if (a && b) cudaEventElapsedTime(&ms, a, b);
    return ms;
}

void collect_timings(nk_counter* h, const PhaseEvents& pe) {
    float mark = 0.f, cnt = 0.f;
    for (size_t i = 0; i < pe.mark0.size(); ++i) {
        mark += ev_ms(pe.mark0[i], pe.count0[i]);
        cnt += ev_ms(pe.count0[i], pe.count1[i]);
    }
    if (pe.copy0) h->last.h2d_ms = ev_ms(pe.copy0, pe.copy1);
    h->last.mark_ms += mark;
    h->last.count_ms += cnt;
    if (pe.fold0) h->last.fold_ms = ev_ms(pe.fold0, pe.fold1);
    if (pe.fold1) h->last.lif_ms = ev_ms(pe.fold1, pe.lif1);
    if (pe.begin) h->last.total_ms = ev_ms(pe.begin, pe.end);
}

// enqueue the read-back of {new spikes, max spikes, kmers}; nothing waits here
int finish_call(nk_counter* h, bool had_lif, const PhaseEvents* pe) {
    if (h->pending_pack) {  // fused post kernel: scalars and the sorted top-N rows come back in one copy
        const size_t bytes = (nk::PACK_HDR + 2 * h->top_cached_n) * sizeof(unsigned long long);
        if (!h->pack_direct) NK_CUDA(cudaMemcpyAsync(h->h_pack, h->d_pack, bytes, cudaMemcpyDeviceToHost, h->stream));
        h->pack_direct = false;
        h->last.d2h_bytes += bytes;   // (written by the kernel itself when the pack lives in mapped host memory)
    } else {
        NK_CUDA(cudaMemcpyAsync(h->h_scalars, h->scalars, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
        h->last.d2h_bytes += 3 * sizeof(unsigned long long);
    }
    h->pending = true;
    h->pending_lif = had_lif;
    h->pending_timings = pe != nullptr;
    if (pe) h->pend_pe = *pe;
    return NK_OK;
}

// wait for the in-flight read-back (if any) and fold it into the EnergyTracker mirrors
int resolve(nk_counter* h) {
    if (!h->pending) return NK_OK;
    NK_CUDA(cudaSetDevice(h->cfg.device));
    NK_CUDA(cudaStreamSynchronize(h->stream));
    h->pending = false;
    const bool had_pack = h->pending_pack;
    if (h->pending_pack) {
        h->h_scalars[0] = h->h_pack[0];
        h->h_scalars[2] = h->h_pack[2];
        h->pending_pack = false;
        if (h->h_pack[1] != 0) {  // peer-signalled multi-GPU job: a rank's signal never arrived
            h->pending_lif = false;
            h->top_cache_valid = false;
            h->dist_failed = true;
            h->pend_pe = PhaseEvents{};
            return fail(NK_ERR_STATE, "multi-GPU job failed: timed out waiting for a peer rank (code %llu: 1 = counting-finished "
                        "signal, 2 = result pack)", h->h_pack[1]);
        }
    }
    if (h->pending_lif) {
        const unsigned long long fired = h->h_scalars[0];
        h->total_spikes += fired;
        // src/models.rs:162-163 / src/spiking_hash.rs:649-655
        h->energy_fixed += fired * cost_fixed(h->cfg.spike_cost);
    }
    h->last.kmers = h->h_scalars[2];
    if (had_pack) {
        // %globaltimer stamps of the fused kernel(s), nanoseconds on this GPU's clock (see nk_post.cu)
        const unsigned long long* t = h->h_pack + 4;
        auto ms = [](unsigned long long a, unsigned long long b) { return b > a ? (float)((double)(b - a) * 1e-6) : 0.f; };
        h->last.post_ms = ms(t[0], t[3]);
        if (getenv("NK_POST_TRACE") && !h->dist_job)
            fprintf(stderr, "[post trace] start->ready %.1f us, phase 1 + barrier %.1f, select %.1f, look-back %.1f, gather+ticket %.1f, final %.1f (last block %llu)\n",
                    ms(t[0], t[1]) * 1e3, ms(t[1], t[2]) * 1e3, ms(t[2], t[4]) * 1e3, ms(t[4], t[5]) * 1e3, ms(t[5], t[6]) * 1e3, ms(t[6], t[3]) * 1e3, t[7]);
        h->last.exch_reduce_ms = ms(t[1], t[2]);  // single GPU: the same phase on local memory
        if (h->dist_job) {
            h->last.exch_wait_ms = ms(t[0], t[1]) + ms(t[4], t[5]);
            h->last.merge_ms = ms(t[5], t[6]);
            // peer memory read by this rank's slice kernel (u32 counts of its slice on every OTHER rank) + the packs
            h->last.exch_bytes = (unsigned long long)(h->dist_world - 1) *
                                 (h->dist_len * 4ull + (nk::PACK_HDR + 2 * h->dist_n_each) * 8ull);
        }
        h->dist_job = false;
    }
    if (h->pending_timings) collect_timings(h, h->pend_pe);
    h->pend_pe = PhaseEvents{};
    return NK_OK;
}

void begin_call(nk_counter* h) {
    resolve(h);
    h->rows_valid = false;  // a new job: the rows fixed by nk_uniques_begin no longer describe the state
    h->ev_used = 0;
    const float topn = h->last.topn_ms;
    h->last = nk_timings{};
    h->last.topn_ms = topn;
}

int free_devbuf(DevBuf& b) {
    if (b.bases) cudaFree(b.bases);
    if (b.codes) cudaFree(b.codes);
    if (b.other) cudaFree(b.other);
    if (b.invalid) cudaFree(b.invalid);
    if (b.copy_done) cudaEventDestroy(b.copy_done);
    if (b.compute_done) cudaEventDestroy(b.compute_done);
    b = DevBuf{};
    return NK_OK;
}

int fold_and_simulate(nk_counter* h, bool skip_zero, PhaseEvents& pe, bool with_topn) {
    NK_TRY(get_event(h, &pe.fold0));
    NK_CUDA(cudaEventRecord(pe.fold0, h->stream));
    const uint64_t launches_before = h->last.launches;
    const bool zero_totals = !h->acc_dirty && h->currents_valid_overwrite;
    NK_TRY(unspill(h));
    if (zero_totals) {
        // nothing was counted by this call: totals are all zero (currents are OVERWRITTEN, :174-176)
        NK_TRY(materialize_zero(h));
        NK_CUDA(cudaMemsetAsync(h->currents, 0, h->cfg.pool_size * sizeof(unsigned long long), h->stream));
        h->currents_valid_overwrite = false;
    }
    if (h->last.launches == launches_before && !zero_totals) {
        pe.fold1 = pe.fold0;   // the fold is fused into the LIF kernel: nothing was enqueued for this phase
    } else {
        NK_TRY(get_event(h, &pe.fold1));
        NK_CUDA(cudaEventRecord(pe.fold1, h->stream));
    }
    NK_TRY(simulate(h, skip_zero, with_topn));  // folds acc -> currents inside the LIF kernel
    NK_TRY(get_event(h, &pe.lif1));
    NK_CUDA(cudaEventRecord(pe.lif1, h->stream));
    if (h->exact) {  // counts.clear() + refill, kmer_per_neuron rebuilt (:157-172, :426-427, :467-473)
        // (the stored currents are this call's per-neuron totals: the bucket sizes come from them)
        cudaError_t e = nk::exact_finalize(h->xt, h->fm, h->cfg.pool_size, h->currents, false, h->stream);
        if (e == cudaErrorInvalidValue) return fail(NK_ERR_UNSUPPORTED, "exact counts: the words could not be partitioned into buckets that fit the on-chip tables");
        NK_CUDA(e);
    }
    return NK_OK;
}

// get_count on one table of `h`'s device (the handle's own table, or a group member's slice table)
int exact_lookup_in(nk_counter* h, nk::ExactTable& t, uint64_t kmer, uint32_t* count, int32_t* found) {
    NK_CUDA(cudaSetDevice(h->cfg.device));
    NK_TRY(resolve(h));
    *count = 0;
    *found = 0;
    if (!t.valid || t.n_keys == 0) return NK_OK;
    NK_CUDA(nk::exact_lookup(t, h->fm, kmer, h->scalars + 4, h->stream));
    NK_CUDA(cudaMemcpyAsync(h->h_scalars + 4, h->scalars + 4, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    NK_CUDA(cudaStreamSynchronize(h->stream));
    *found = (int32_t)h->h_scalars[4];
    *count = (uint32_t)h->h_scalars[5];
    return NK_OK;
}

// one table's (key, count) rows to host arrays of t.n_keys entries (either may be null)
int exact_copy_table_of(nk_counter* h, nk::ExactTable& t, uint64_t* keys, uint32_t* counts) {
    NK_CUDA(cudaSetDevice(h->cfg.device));
    NK_TRY(resolve(h));
    const unsigned long long n = t.valid ? t.n_keys : 0;
    if (n == 0) return NK_OK;
    // the table lives bucket by bucket: compact it into dense device arrays, then copy those out
    unsigned long long* dk = nullptr;
    unsigned int* dc = nullptr;
    if (keys) NK_CUDA(cudaMalloc(&dk, n * 8));
    if (counts && cudaMalloc(&dc, n * 4) != cudaSuccess) { cudaFree(dk); return fail(NK_ERR_OOM, "cudaMalloc(exact table copy)"); }
    cudaError_t e = nk::exact_dense_copy(t, dk, dc, nullptr, h->stream);
    if (e == cudaSuccess && keys) e = cudaMemcpyAsync(keys, dk, n * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && counts) e = cudaMemcpyAsync(counts, dc, n * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dk);
    cudaFree(dc);
    NK_CUDA(e);
    return NK_OK;
}

}  // namespace nkd

using namespace nkd;

// ---------------------------------------------------------------------------
// Whole-file driver (nk_process_file): FastxReader -> pinned batch -> count_host_batch.
// Replaces the reference's producer thread + worker channels (src/spiking_hash.rs:285-422):
// the parser fills a pinned batch while the previous batch's kernels run; a sequence that
// does not fit is cut into pieces that overlap by k-1 bases (every window counted once).
// ---------------------------------------------------------------------------
namespace nk {

namespace {

// ---- parallel FASTA ingest -------------------------------------------------------------------
// Plain FASTA files are memory-mapped and cut into windows; a pool of host threads strips headers
// and line terminators of different windows concurrently into pinned buffers, and this thread
// pushes the finished windows IN FILE ORDER as batches (async H2D on the copy stream, kernels on
// the compute stream).  Sequence data before the first header of a window continues the record
// left open by the previous window: it is prefixed with that record's last k-1 bases (kept in
// `carry`), so every window of every record is counted exactly once.  FASTQ (whose record
// boundaries are ambiguous at an arbitrary line) and gzip input use the serial reader.
struct FaSlot {
    uint8_t* buf = nullptr;            // pinned; data starts at buf + 32 (headroom for the k-1 overlap)
    size_t fill = 0;
    std::vector<uint64_t> rec_starts;
    FastaWindowPlan plan;
    cudaEvent_t copied = nullptr;
    bool copy_inflight = false;
    int state = 0;                      // 0 free, 1 queued/parsing, 2 parsed
    // FASTQ windows: [a, b) is a whole number of records; offs = record ends; ok = no malformed record
    int kind = 0;                       // 0 FASTA window, 1 FASTQ window
    size_t a = 0, b = 0, cap = 0;
    std::vector<uint64_t> offs;
    bool ok = true;
};

struct FaPool {
    const uint8_t* file = nullptr;
    std::vector<FaSlot> slots;
    std::vector<int> queue;             // slot indices waiting for a worker
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    bool quit = false;
    std::vector<std::thread> threads;

    void worker() {
        for (;;) {
            int si;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return quit || !queue.empty(); });
                if (queue.empty()) return;
                si = queue.front();
                queue.erase(queue.begin());
            }
            FaSlot& sl = slots[si];
            if (sl.kind == 0) fasta_parse_window(file, sl.plan, sl.buf + 32, &sl.fill, &sl.rec_starts);
            else sl.ok = fastq_parse_window(file, sl.a, sl.b, sl.buf, sl.cap, &sl.fill, &sl.offs);
            {
                std::lock_guard<std::mutex> lk(mu);
                sl.state = 2;
            }
            cv_done.notify_all();
        }
    }
    void stop() {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
        }
        cv_work.notify_all();
        for (auto& t : threads) t.join();
        threads.clear();
    }
};

// returns NK_OK with *done = false when this path does not apply (caller falls back to the serial reader)
int count_fasta_parallel(nk_counter* h, const char* path, PhaseEvents& pe, std::string* err, bool* done) {
    *done = false;
    unsigned nthreads = std::thread::hardware_concurrency();
    if (const char* e = getenv("NK_FASTA_THREADS")) nthreads = (unsigned)atoi(e);
    if (nthreads > 8) nthreads = 8;
    size_t window = 2u << 20, min_size = 256u << 20;  // measured: no gain at 115 MB (pinning + thread start-up), 4-6x at 2 GB
    if (const char* e = getenv("NK_FASTA_WINDOW")) { window = (size_t)atoll(e); min_size = 2 * window; }
    if (nthreads < 2 || window < 64) return NK_OK;
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return NK_OK;
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < min_size) { ::close(fd); return NK_OK; }
    const size_t size = (size_t)st.st_size;
    void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (map == MAP_FAILED) return NK_OK;
    madvise(map, size, MADV_SEQUENTIAL);

    FaPool pool;
    pool.file = (const uint8_t*)map;
    const int nslots = (int)nthreads + 3;
    pool.slots.resize(nslots);
    int rc = NK_OK;
    for (auto& sl : pool.slots) {
        if (cudaMallocHost((void**)&sl.buf, window + kFastaSlack + 64) != cudaSuccess) { rc = NK_ERR_OOM; *err = "cudaMallocHost(FASTA window)"; break; }
        cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming);
    }
    if (rc == NK_OK) {
        for (unsigned t = 0; t < nthreads; ++t) pool.threads.emplace_back([&pool] { pool.worker(); });
        const unsigned k = h->cfg.k;
        uint8_t carry[32];
        size_t carry_len = 0;
        std::vector<uint64_t> offsets;
        size_t next_ws = 0;
        int next_state = FA_LINE_START;
        unsigned long long dispatched = 0, consumed = 0;
        while (rc == NK_OK && (next_ws < size || consumed < dispatched)) {
            // dispatch as many windows as there are free slots
            while (next_ws < size && dispatched - consumed < (unsigned long long)nslots) {
                FaSlot& sl = pool.slots[dispatched % nslots];
                if (sl.copy_inflight) { cudaEventSynchronize(sl.copied); sl.copy_inflight = false; }
                int st_next = FA_LINE_START;
                sl.plan = fasta_plan_window(pool.file, size, next_ws, window, next_state, &st_next);
                next_ws = sl.plan.we;
                next_state = st_next;
                {
                    std::lock_guard<std::mutex> lk(pool.mu);
                    sl.state = 1;
                    pool.queue.push_back((int)(dispatched % nslots));
                }
                pool.cv_work.notify_one();
                ++dispatched;
            }
            if (consumed == dispatched) break;
            // consume the next window in file order
            FaSlot& sl = pool.slots[consumed % nslots];
            {
                std::unique_lock<std::mutex> lk(pool.mu);
                pool.cv_done.wait(lk, [&] { return sl.state == 2; });
                sl.state = 0;
            }
            uint8_t* data = sl.buf + 32;
            const size_t ov = std::min<size_t>(carry_len, k - 1);
            memcpy(data - ov, carry + (carry_len - ov), ov);
            offsets.clear();
            offsets.push_back(0);
            for (uint64_t r : sl.rec_starts) offsets.push_back(ov + r);
            offsets.push_back(ov + sl.fill);
            // last k-1 bases of the record that is open at the end of this window
            const size_t open_from = sl.rec_starts.empty() ? 0 : (size_t)sl.rec_starts.back();
            const size_t open_len = sl.fill - open_from;
            if (sl.rec_starts.empty()) {
                // still the same record: tail of (carry + data)
                uint8_t tmp[64];
                size_t n = 0;
                const size_t keep_c = std::min<size_t>(carry_len, 31);
                memcpy(tmp, carry + (carry_len - keep_c), keep_c);
                n = keep_c;
                const size_t keep_d = std::min<size_t>(sl.fill, 31);
                memcpy(tmp + n, data + sl.fill - keep_d, keep_d);
                n += keep_d;
                carry_len = std::min<size_t>(n, 31);
                memmove(carry, tmp + (n - carry_len), carry_len);
            } else {
                carry_len = std::min<size_t>(open_len, 31);
                memcpy(carry, data + sl.fill - carry_len, carry_len);
            }
            if (ov + sl.fill > 0) {
                rc = count_host_batch(h, data - ov, offsets.data(), offsets.size() - 1, &pe, kPushFileDriver);
                if (rc != NK_OK) { *err = g_err; break; }
                cudaEventRecord(sl.copied, h->copy_stream);
                sl.copy_inflight = true;
            }
            ++consumed;
        }
    }
    pool.stop();
    cudaStreamSynchronize(h->copy_stream);
    for (auto& sl : pool.slots) {
        if (sl.buf) cudaFreeHost(sl.buf);
        if (sl.copied) cudaEventDestroy(sl.copied);
    }
    munmap(map, size);
    *done = rc == NK_OK;
    return rc;
}

// FASTQ twin of count_fasta_parallel (see nk_host.h for how record starts are found)
int count_fastq_parallel(nk_counter* h, const char* path, PhaseEvents& pe, std::string* err, bool* done) {
    *done = false;
    unsigned nthreads = std::thread::hardware_concurrency();
    if (const char* e = getenv("NK_FASTA_THREADS")) nthreads = (unsigned)atoi(e);
    if (nthreads > 8) nthreads = 8;
    size_t window = 4u << 20, min_size = 256u << 20;
    if (const char* e = getenv("NK_FASTA_WINDOW")) { window = (size_t)atoll(e); min_size = 2 * window; }
    if (nthreads < 2 || window < 64) return NK_OK;
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return NK_OK;
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < min_size) { ::close(fd); return NK_OK; }
    const size_t size = (size_t)st.st_size;
    void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (map == MAP_FAILED) return NK_OK;
    madvise(map, size, MADV_SEQUENTIAL);
    const uint8_t* file = (const uint8_t*)map;

    // pass 1: newlines per fixed range (parallel), prefix sum, first record start of every range
    const size_t nr = (size + window - 1) / window;
    std::vector<uint64_t> nl(nr + 1, 0);
    {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthreads; ++t)
            th.emplace_back([&, t] {
                for (size_t i = t; i < nr; i += nthreads)
                    nl[i + 1] = fastq_count_newlines(file, i * window, std::min(size, (i + 1) * window));
            });
        for (auto& x : th) x.join();
    }
    for (size_t i = 0; i < nr; ++i) nl[i + 1] += nl[i];
    std::vector<size_t> starts(nr + 1, size);
    size_t max_win = 0;
    for (size_t i = 0; i < nr; ++i) starts[i] = fastq_first_record_start(file, size, i * window, nl[i]);
    for (size_t i = 0; i < nr; ++i) max_win = std::max(max_win, starts[i + 1] - starts[i]);
    const size_t slack = std::max<size_t>(kFastaSlack, window / 4);
    if (max_win > window + slack) {  // reads longer than a window: leave the file to the serial reader
        munmap(map, size);
        return NK_OK;
    }

    FaPool pool;
    pool.file = file;
    const int nslots = (int)nthreads + 3;
    pool.slots.resize(nslots);
    int rc = NK_OK;
    const size_t cap = (window + slack) / 2 + 64;  // |seq| == |qual|, so the sequences are under half of a window
    for (auto& sl : pool.slots) {
        if (cudaMallocHost((void**)&sl.buf, cap) != cudaSuccess) { rc = NK_ERR_OOM; *err = "cudaMallocHost(FASTQ window)"; break; }
        cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming);
        sl.kind = 1;
        sl.cap = cap;
    }
    if (rc == NK_OK) {
        for (unsigned t = 0; t < nthreads; ++t) pool.threads.emplace_back([&pool] { pool.worker(); });
        size_t dispatched = 0, consumed = 0;
        bool stop = false;
        while (rc == NK_OK && (consumed < dispatched || (!stop && dispatched < nr))) {
            while (!stop && dispatched < nr && dispatched - consumed < (size_t)nslots) {
                FaSlot& sl = pool.slots[dispatched % nslots];
                if (sl.copy_inflight) { cudaEventSynchronize(sl.copied); sl.copy_inflight = false; }
                sl.a = starts[dispatched];
                sl.b = starts[dispatched + 1];
                {
                    std::lock_guard<std::mutex> lk(pool.mu);
                    sl.state = 1;
                    pool.queue.push_back((int)(dispatched % nslots));
                }
                pool.cv_work.notify_one();
                ++dispatched;
            }
            if (consumed == dispatched) break;
            FaSlot& sl = pool.slots[consumed % nslots];
            {
                std::unique_lock<std::mutex> lk(pool.mu);
                pool.cv_done.wait(lk, [&] { return sl.state == 2; });
                sl.state = 0;
            }
            if (!stop && sl.fill > 0 && sl.offs.size() > 1) {
                rc = count_host_batch(h, sl.buf, sl.offs.data(), sl.offs.size() - 1, &pe, kPushFileDriver);
                if (rc != NK_OK) { *err = g_err; break; }
                cudaEventRecord(sl.copied, h->copy_stream);
                sl.copy_inflight = true;
            }
            if (!sl.ok) stop = true;  // first malformed record: the iteration ends here (src/utils.rs:17-20)
            ++consumed;
        }
    }
    pool.stop();
    cudaStreamSynchronize(h->copy_stream);
    for (auto& sl : pool.slots) {
        if (sl.buf) cudaFreeHost(sl.buf);
        if (sl.copied) cudaEventDestroy(sl.copied);
    }
    munmap(map, size);
    *done = rc == NK_OK;
    return rc;
}

}  // namespace

// ingest_only: the second read of a file by the uniques pass — the batches are routed to
// uniques_host_batch (count_host_batch checks h->uniques_open), nothing is simulated or read back
// Plain FASTA / FASTQ files: raw bytes -> device (pool of host threads, pread + pinned slots + async H2D), records
// parsed THERE (nk_parse.cu), then the ordinary count kernels over the device-resident result.  *handled = false:
// the host reader below takes the file (compressed input, pipes, files larger than the device can hold).
static int process_file_device(nk_counter* h, const char* path, bool streaming, std::string* err, bool ingest_only, bool* handled) {
    *handled = false;
    cudaError_t ce = cudaSetDevice(h->cfg.device);
    if (ce != cudaSuccess) { *err = std::string("cudaSetDevice: ") + cudaGetErrorString(ce); return NK_ERR_CUDA; }
    const unsigned long long slice = (0xFFFFFFFFull / nk::COUNT_TILE - 1) * nk::COUNT_TILE;
    auto failed = [&](int rc) { *err = g_err; return rc; };
    if (ingest_only) {
        // the uniques pass: the parsed file of the counting pass is normally still resident
        struct stat st;
        bool same = h->fp_valid && h->fp_path == path && ::stat(path, &st) == 0 && (unsigned long long)st.st_size == h->fp_size &&
                    (unsigned long long)st.st_mtim.tv_sec * 1000000000ull + (unsigned long long)st.st_mtim.tv_nsec == h->fp_mtime_ns;
        unsigned long long nb = h->fp_nbases, nr = h->fp_nrec;
        if (!same) {
            bool fq = false;
            int rc = parse_file_on_device(h, path, handled, &fq, &nb, &nr, err, nullptr);
            if (rc != NK_OK || !*handled) return rc;
        }
        *handled = true;
        for (unsigned long long c0 = 0; c0 < nb && nr > 0; c0 += slice) {
            DevBuf view = h->staged;
            view.bases = h->staged.bases + c0;
            int rc = uniques_chunk(h, view, h->staged_offsets, 0, nr, c0, std::min(slice, nb - c0), false);
            if (rc != NK_OK) return failed(rc);
        }
        return NK_OK;
    }
    timespec tr0, tr1, tr2;
    clock_gettime(CLOCK_MONOTONIC, &tr0);
    PhaseEvents pe;
    int rc = NK_OK;
    bool fq = false;
    unsigned long long nb = 0, nr = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        begin_call(h);
        pe = PhaseEvents{};
        if ((rc = get_event(h, &pe.begin)) != NK_OK) return failed(rc);
        cudaEventRecord(pe.begin, h->stream);
        NK_TRY(zero_kmers(h));
        h->currents_valid_overwrite = true;
        // stage + parse + count (FASTA: the counting overlaps the read)
        rc = parse_file_on_device(h, path, handled, &fq, &nb, &nr, err, &pe);
        if (rc == NK_ERR_STATE && attempt == 0) {
            // the overlapped path met more records than it had room for: void the partial counts, take the two-phase path
            if (h->acc_dirty) { cudaMemsetAsync(h->acc, 0, h->cfg.pool_size * sizeof(unsigned int), h->stream); h->acc_dirty = false; h->acc_kmers = 0; }
            if (h->exact) nk::exact_clear(h->xt, h->cfg.pool_size, false, h->stream);
            h->kmers_clean = false;
            h->file_no_overlap = true;
            continue;
        }
        break;
    }
    h->file_no_overlap = false;
    if (rc != NK_OK || !*handled) return rc;
    clock_gettime(CLOCK_MONOTONIC, &tr1);
    if ((rc = fold_and_simulate(h, /*skip_zero=*/!streaming, pe, true)) != NK_OK) return failed(rc);
    if ((rc = get_event(h, &pe.end)) != NK_OK) return failed(rc);
    cudaEventRecord(pe.end, h->stream);
    if ((rc = finish_call(h, true, &pe)) != NK_OK) return failed(rc);
    if ((rc = resolve(h)) != NK_OK) return failed(rc);
    clock_gettime(CLOCK_MONOTONIC, &tr2);
    if (getenv("NK_FILE_TRACE"))
        fprintf(stderr, "[file trace] begin + stage + parse %.2f ms, count + post + read-back %.2f ms\n",
                (tr1.tv_sec - tr0.tv_sec) * 1e3 + (tr1.tv_nsec - tr0.tv_nsec) * 1e-6, (tr2.tv_sec - tr1.tv_sec) * 1e3 + (tr2.tv_nsec - tr1.tv_nsec) * 1e-6);
    return NK_OK;
}

int process_file(nk_counter* h, const char* path, bool streaming, std::string* err, bool ingest_only) {
    if (!is_group(h)) {
        bool handled = false;
        const int rc = process_file_device(h, path, streaming, err, ingest_only, &handled);
        if (rc != NK_OK || handled) return rc;
    }
    FastxReader rd;
    if (rd.open(path, err) != 0) return NK_ERR_IO;
    auto cuda_fail = [&](cudaError_t e, const char* what) {
        *err = std::string(what) + ": " + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? NK_ERR_OOM : NK_ERR_CUDA;
    };
    cudaError_t ce = cudaSetDevice(h->cfg.device);
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaSetDevice");
    const size_t cap = (size_t)kChunkBytes;
    // pinned, double-buffered host staging: the parser fills one batch while the other one's
    // H2D copy (copy stream) and kernels (compute stream) are in flight
    // (the two batches live in the handle: pinning 64 MB costs ~25 ms, and the uniques pass reads the file again)
    uint8_t* batches[2] = {nullptr, nullptr};
    cudaEvent_t copied[2] = {nullptr, nullptr};
    bool inflight[2] = {false, false};
    for (int i = 0; i < 2; ++i) {
        if (!h->file_batch[i]) {
            ce = cudaMallocHost((void**)&h->file_batch[i], cap);
            if (ce != cudaSuccess) { h->file_batch[i] = nullptr; return cuda_fail(ce, "cudaMallocHost(batch)"); }
        }
        batches[i] = h->file_batch[i];
        cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming);
    }
    int cur = 0;
    uint8_t* batch = batches[0];
    std::vector<uint64_t> offsets;
    offsets.reserve(1u << 20);
    offsets.push_back(0);
    size_t fill = 0;
    const unsigned k = h->cfg.k;

    // a multi-GPU group handle: every batch is sharded over the members (group_push waits for the members'
    // copies, and for the kernels that read the pinned batch in place, before the parser refills it)
    const bool grp = is_group(h);
    if (!ingest_only && !grp) begin_call(h);
    PhaseEvents pe;
    int rc = NK_OK;
    if (grp && !ingest_only && (rc = group_begin(h)) != NK_OK) { *err = g_err; for (int i = 0; i < 2; ++i) cudaEventDestroy(copied[i]); return rc; }
    auto flush = [&]() -> int {
        if (grp) {
            if (offsets.size() > 1 && fill > 0) {
                int r = ingest_only ? nk_uniques_push(h->group[0], batch, offsets.data(), offsets.size() - 1)
                                    : group_push(h, batch, nullptr, nullptr, offsets.data(), offsets.size() - 1);
                if (r != NK_OK) { *err = g_err; return r; }
            }
            offsets.clear();
            offsets.push_back(0);
            fill = 0;
            return NK_OK;
        }
        if (offsets.size() > 1 && fill > 0) {
            int r = count_host_batch(h, batch, offsets.data(), offsets.size() - 1, &pe, kPushFileDriver);
            if (r != NK_OK) { *err = g_err; return r; }
            cudaEventRecord(copied[cur], h->copy_stream);
            inflight[cur] = true;
            cur ^= 1;
            batch = batches[cur];
            if (inflight[cur]) { cudaEventSynchronize(copied[cur]); inflight[cur] = false; }
        }
        offsets.clear();
        offsets.push_back(0);
        fill = 0;
        return NK_OK;
    };
    do {
        if (!ingest_only && !grp) {
            if (get_event(h, &pe.begin) != NK_OK) { *err = g_err; rc = NK_ERR_CUDA; break; }
            cudaEventRecord(pe.begin, h->stream);
            zero_kmers(h);
            h->currents_valid_overwrite = true;
        }
        bool stop = false;
        if (!rd.is_compressed() && !grp) {
            bool handled = false;
            rc = rd.is_fastq() ? count_fastq_parallel(h, path, pe, err, &handled) : count_fasta_parallel(h, path, pe, err, &handled);
            if (rc != NK_OK) break;
            stop = handled;  // the whole file was ingested by the parallel path
        }
        while (!stop && rd.next_record()) {
            if (rd.is_fastq()) {
                // a FASTQ record is validated as a whole before it counts: keep it inside one batch
                size_t start = fill;
                bool done = false;
                for (;;) {
                    fill += rd.read_seq(batch + fill, cap - fill, &done);
                    if (done) break;
                    if (start == 0) { *err = "FASTQ record longer than the 32 MiB batch buffer"; rc = NK_ERR_UNSUPPORTED; break; }
                    // move the partial record to the front of a fresh batch
                    const size_t part = fill - start;
                    fill = start;
                    std::vector<uint8_t> tmp(batch + start, batch + start + part);
                    if ((rc = flush()) != NK_OK) break;
                    memcpy(batch, tmp.data(), part);
                    fill = part;
                    start = 0;
                }
                if (rc != NK_OK) break;
                if (!rd.finish_record(fill - start)) { fill = start; stop = true; break; }  // malformed: drop + stop
                offsets.push_back(fill);
            } else {
                bool done = false;
                while (!done) {
                    if (fill == cap) {
                        // close the piece, push, and restart with the last k-1 bases as overlap
                        offsets.push_back(fill);
                        uint8_t tail[32];
                        const size_t ov = std::min<size_t>(k - 1, fill - offsets[offsets.size() - 2]);
                        memcpy(tail, batch + fill - ov, ov);
                        if ((rc = flush()) != NK_OK) break;
                        memcpy(batch, tail, ov);
                        fill = ov;
                    }
                    fill += rd.read_seq(batch + fill, cap - fill, &done);
                }
                if (rc != NK_OK) break;
                offsets.push_back(fill);
            }
            if (fill == cap && (rc = flush()) != NK_OK) break;
        }
        if (rc != NK_OK) break;
        if ((rc = flush()) != NK_OK) break;
        if (ingest_only) break;
        if (grp) {
            if ((rc = group_end(h, /*skip_zero=*/!streaming)) != NK_OK) { *err = g_err; break; }
            uint64_t spikes = 0;
            if ((rc = nk_total_spikes(h, &spikes)) != NK_OK) *err = g_err;  // observes the result: surfaces device errors here
            break;
        }
        if ((rc = fold_and_simulate(h, /*skip_zero=*/!streaming, pe, true)) != NK_OK) { *err = g_err; break; }
        if (get_event(h, &pe.end) != NK_OK) { *err = g_err; rc = NK_ERR_CUDA; break; }
        cudaEventRecord(pe.end, h->stream);
        if ((rc = finish_call(h, true, &pe)) != NK_OK) { *err = g_err; break; }
        if ((rc = resolve(h)) != NK_OK) { *err = g_err; break; }
    } while (0);
    if (grp) {
        if (rc != NK_OK && !ingest_only) group_reset(h);
    } else {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamSynchronize(h->stream);
    }
    for (int i = 0; i < 2; ++i) cudaEventDestroy(copied[i]);
    return rc;
}

}  // namespace nk

extern "C" {

const char* nk_last_error(void) { return g_err.c_str(); }
const char* nk_version(void) { return "neurokmer-b200 0.1.0 (sm_100a)"; }

int nk_config_default(nk_config* cfg) {
    if (!cfg) return fail(NK_ERR_BAD_ARG, "null cfg");
    std::memset(cfg, 0, sizeof *cfg);
    cfg->k = 31;               // src/main.rs:14-15
    cfg->pool_size = 1000000;  // src/main.rs:17-18
    cfg->use_canonical = 0;    // src/main.rs:20-21
    cfg->threshold = 1.0f;     // src/main.rs:37
    cfg->leak = 0.95f;
    cfg->refractory = 2;
    cfg->spike_cost = 1.0;
    cfg->steps = 1000;         // src/spiking_hash.rs:70
    cfg->device = 0;
    return NK_OK;
}

int nk_create(const nk_config* cfg, nk_counter** out) {
    if (!cfg || !out) return fail(NK_ERR_BAD_ARG, "null argument");
    *out = nullptr;
    if (cfg->k < 1 || cfg->k > 32) return fail(NK_ERR_BAD_ARG, "k must be in [1,32], got %u", cfg->k);
    if (cfg->pool_size == 0) return fail(NK_ERR_BAD_ARG, "pool_size must be > 0");
    if (cfg->pool_size >= (1ull << 32)) return fail(NK_ERR_UNSUPPORTED, "pool_size must be < 2^32");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(NK_ERR_NO_DEVICE, "no CUDA device (%s); libneurokmer has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(NK_ERR_BAD_ARG, "device %d out of range (%d devices)", cfg->device, ndev);
    cudaDeviceProp prop{};
    NK_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(NK_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
    NK_CUDA(cudaSetDevice(cfg->device));

    nk_counter* h = new nk_counter();
    h->cfg = *cfg;
    h->fm = nk::make_fastmod(cfg->pool_size);
    if (const char* e = std::getenv("NK_MOD_TWO_STAGE")) {  // A/B switch: keep the two-stage remainder in the count kernel
        if (e[0] == '1') h->fm.kind = 0u;
    }
    auto bail = [&](int rc) { nk_destroy(h); return rc; };
#define NK_C(expr)                                                                                              \
    do {                                                                                                        \
        cudaError_t e_ = (expr);                                                                                \
        if (e_ != cudaSuccess)                                                                                  \
            return bail(fail(e_ == cudaErrorMemoryAllocation ? NK_ERR_OOM : NK_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_))); \
    } while (0)
    NK_C(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    NK_C(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    const unsigned long long P = cfg->pool_size;
    // the accumulators carry the multi-GPU mailbox at their tail: one IPC handle maps both into the peers
    NK_C(cudaMalloc(&h->acc, nk::dist_alloc_bytes(P)));
    NK_C(cudaMemset(reinterpret_cast<unsigned char*>(h->acc) + nk::dist_mail_offset(P), 0, nk::DIST_MAIL_BYTES));
    h->spill = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(h->acc) + nk::dist_spill_offset(P));
    NK_C(cudaMalloc(&h->currents, P * sizeof(unsigned long long)));
    NK_C(cudaMalloc(&h->v, P * sizeof(float)));
    NK_C(cudaMalloc(&h->r, P * sizeof(unsigned int)));
    NK_C(cudaMalloc(&h->spikes, P * sizeof(unsigned long long)));
    NK_C(cudaMalloc(&h->scalars, 8 * sizeof(unsigned long long)));
    NK_C(cudaMalloc(&h->tile_counter, 64));
    NK_C(cudaMemset(h->tile_counter, 0, 64));  // [0] tile cursor, [1] finished CTAs: the count kernel re-arms both itself
    NK_C(cudaMallocHost(&h->h_scalars, 8 * sizeof(unsigned long long)));
    NK_C(cudaMalloc(&h->topn.hist, 256 * sizeof(unsigned int)));
    NK_C(cudaMalloc(&h->topn.ctrl, 8 * sizeof(unsigned long long)));
    NK_C(cudaMalloc(&h->topn.block_counts, ((P + nk::POST_SEG_ITEMS - 1) / nk::POST_SEG_ITEMS + 1) * sizeof(unsigned int)));
    NK_C(nk::post_max_grid(cfg->device, &h->post_grid));
    NK_C(cudaMalloc(&h->post_zero, nk::POST_SCRATCH_BYTES));
    NK_C(cudaMemset(h->post_zero, 0, nk::POST_SCRATCH_BYTES));  // the fused kernel leaves it zeroed
    NK_C(cudaMalloc(&h->d_pack, nk::PACK_MAX_U64 * sizeof(unsigned long long)));
    NK_C(cudaMallocHost(&h->h_pack, nk::PACK_MAX_U64 * sizeof(unsigned long long)));
    {
        void* dp = nullptr;
        if (cudaHostGetDevicePointer(&dp, h->h_pack, 0) == cudaSuccess) h->h_pack_dev = static_cast<unsigned long long*>(dp);
        else cudaGetLastError();
    }
    NK_C(cudaMalloc(&h->topn.out_idx, 2048 * sizeof(unsigned long long)));
    NK_C(cudaMalloc(&h->topn.out_spikes, 2048 * sizeof(unsigned long long)));
    NK_C(cudaMallocHost(&h->h_top, 2 * 2048 * sizeof(unsigned long long)));
    h->topn_cap = 2048;
#undef NK_C
    if (cudaMemsetAsync(h->acc, 0, P * sizeof(unsigned int), h->stream) != cudaSuccess) return bail(fail(NK_ERR_CUDA, "memset acc"));
    int rc = nk_reset(h);
    if (rc != NK_OK) return bail(rc);
    *out = h;
    return NK_OK;
}

int nk_reset(nk_counter* h) {
    if (is_group(h)) return group_reset(h);
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    const unsigned long long P = h->cfg.pool_size;
    // acc is all-zero whenever no batch is in flight (every fold zeroes it): only a reset in the
    // middle of a stream has to clear it.  currents / v / r / spikes become lazily zero.
    if (h->acc_dirty || h->streaming) NK_CUDA(cudaMemsetAsync(h->acc, 0, P * sizeof(unsigned int), h->stream));
    if (h->spill_dirty) {
        NK_CUDA(cudaMemsetAsync(nk::dist_mail_flags(own_mail(h), 2), 0, 8, h->stream));
        h->spill_dirty = false;
    }
    h->lazy_zero = true;
    h->table_valid = false;  // a reset counter is a fresh counter: its first job rebuilds the LIF table
    h->top_cache_valid = false;
    h->pending_pack = false;
    if (!(h->fired_clean && h->kmers_clean)) NK_CUDA(cudaMemsetAsync(h->scalars, 0, 8 * sizeof(unsigned long long), h->stream));
    h->fired_clean = h->kmers_clean = true;
    if (h->pending) { NK_CUDA(cudaStreamSynchronize(h->stream)); h->pending = false; h->pend_pe = PhaseEvents{}; }
    if (h->exact) NK_CUDA(nk::exact_clear(h->xt, h->cfg.pool_size, true, h->stream));
    h->total_spikes = 0;
    h->energy_fixed = 0;
    h->spike_bound = 0;
    h->slice_only = false;
    h->fresh = true;
    h->streaming = false;
    h->acc_dirty = false;
    h->acc_kmers = 0;
    h->currents_valid_overwrite = true;
    return NK_OK;
}

int nk_destroy(nk_counter* h) {
    if (is_group(h)) return group_destroy(h);
    if (!h) return NK_OK;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    cudaFree(h->acc); cudaFree(h->currents); cudaFree(h->v); cudaFree(h->r); cudaFree(h->spikes);
    cudaFree(h->scalars); cudaFree(h->tile_counter);
    if (h->h_scalars) cudaFreeHost(h->h_scalars);
    cudaFree(h->table.spikes); cudaFree(h->table.v); cudaFree(h->table.r);
    if (h->table_ready) cudaEventDestroy(h->table_ready);
    cudaFree(h->topn.hist); cudaFree(h->topn.ctrl); cudaFree(h->topn.block_counts);
    cudaFree(h->topn.out_idx); cudaFree(h->topn.out_spikes);
    cudaFree(h->post_zero); cudaFree(h->d_pack);
    if (h->h_pack) cudaFreeHost(h->h_pack);
    if (h->h_top) cudaFreeHost(h->h_top);
    cudaFree(h->memo.keys); cudaFree(h->memo.dense); cudaFree(h->memo.res_v); cudaFree(h->memo.res_r);
    cudaFree(h->memo.res_f); cudaFree(h->memo.slot_of); cudaFree(h->memo.ctrl);
    for (int i = 0; i < 2; ++i) if (h->file_batch[i]) cudaFreeHost(h->file_batch[i]);
    nk::exact_free(h->ut);
    cudaFree(h->d_filter);
    cudaFree(h->d_rows);
    for (int r = 0; r < 16; ++r) if (h->dist_ipc_opened[r]) cudaIpcCloseMemHandle((void*)h->dist_peer[r]);
    cudaFree(h->d_merged);
    nk::exact_free(h->xt);
    nk::exact_free(h->xs);
    cudaFree(h->d_top_uniques);
    ingest_free(h);
    free_devbuf(h->buf[0]); free_devbuf(h->buf[1]); free_devbuf(h->staged); free_devbuf(h->zc);
    for (int i = 0; i < 2; ++i) { cudaFree(h->d_offsets2[i]); if (h->offsets_done[i]) cudaEventDestroy(h->offsets_done[i]); }
    cudaFree(h->staged_offsets);
    for (cudaEvent_t e : h->evpool) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    delete h;
    return NK_OK;
}

int nk_set_steps(nk_counter* h, uint64_t steps) {
    if (is_group(h)) return group_set_steps(h, steps);
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    h->cfg.steps = steps;
    return NK_OK;
}
int nk_get_steps(const nk_counter* h, uint64_t* steps) {
    if (!h || !steps) return fail(NK_ERR_BAD_ARG, "null argument");
    *steps = h->cfg.steps;
    return NK_OK;
}

int nk_process_batch(nk_counter* h, const uint8_t* bases, const uint64_t* offsets, uint64_t nseq) {
    if (is_group(h)) { NK_TRY(validate_batch(h, bases, offsets, nseq)); return group_process_batch(h, bases, nullptr, nullptr, offsets, nseq); }
    NK_TRY(validate_batch(h, bases, offsets, nseq));
    if (h->streaming) return fail(NK_ERR_STATE, "nk_process_batch inside nk_stream_begin/end");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    begin_call(h);
    PhaseEvents pe;
    NK_TRY(get_event(h, &pe.begin));
    NK_CUDA(cudaEventRecord(pe.begin, h->stream));
    NK_TRY(zero_kmers(h));
    h->currents_valid_overwrite = true;  // totals of THIS call overwrite the stored currents (:174-176)
    NK_TRY(count_host_batch(h, bases, offsets, nseq, &pe, kPushSync));
    NK_TRY(fold_and_simulate(h, /*skip_zero=*/true, pe, true));
    NK_TRY(get_event(h, &pe.end));
    NK_CUDA(cudaEventRecord(pe.end, h->stream));
    NK_TRY(finish_call(h, true, &pe));
    return NK_OK;
}

int nk_stream_begin(nk_counter* h) {
    if (is_group(h)) return group_begin(h);
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (h->streaming) return fail(NK_ERR_STATE, "nk_stream_begin called twice");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    begin_call(h);
    h->streaming = true;
    h->stream_pe = PhaseEvents{};
    h->currents_valid_overwrite = true;
    NK_TRY(zero_kmers(h));
    return NK_OK;
}

int nk_stream_push(nk_counter* h, const uint8_t* bases, const uint64_t* offsets, uint64_t nseq) {
    if (is_group(h)) { NK_TRY(validate_batch(h, bases, offsets, nseq)); return group_push(h, bases, nullptr, nullptr, offsets, nseq); }
    NK_TRY(validate_batch(h, bases, offsets, nseq));
    if (!h->streaming) return fail(NK_ERR_STATE, "nk_stream_push without nk_stream_begin");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    return count_host_batch(h, bases, offsets, nseq, &h->stream_pe, kPushSync);
}

// ---- pre-packed input ("nk2" layout, include/neurokmer.h) ------------------------------------------
static int validate_packed(const nk_counter* h, const uint32_t* codes, const uint64_t* offsets, uint64_t nseq) {
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (nseq > 0 && !offsets) return fail(NK_ERR_BAD_ARG, "null offsets");
    if (nseq == 0) return NK_OK;
    if (offsets[0] != 0) return fail(NK_ERR_BAD_ARG, "offsets[0] must be 0");
    if (offsets[nseq] > 0 && !codes) return fail(NK_ERR_BAD_ARG, "null codes");
    return NK_OK;
}

uint64_t nk_packed_code_words(uint64_t nbases) { return (nbases + 15) / 16; }
uint64_t nk_packed_other_words(uint64_t nbases) { return (nbases + 31) / 32; }

int nk_pack_bases(const uint8_t* bases, uint64_t nbases, uint32_t* codes, uint32_t* other, int threads, uint64_t* n_other) {
    if (nbases > 0 && (!bases || !codes)) return fail(NK_ERR_BAD_ARG, "null argument");
    const uint64_t n = nk::host_pack_bases(bases, nbases, codes, other, threads, 0);
    if (n_other) *n_other = n;
    return NK_OK;
}

int nk_debug_pack_body(const uint8_t* bases, uint64_t nbases, uint32_t* codes, uint32_t* other, int body, uint64_t* n_other) {
    if (body < 1 || body > 3) return fail(NK_ERR_BAD_ARG, "body must be 1 (portable), 2 (AVX2) or 3 (AVX-512BW)");
    if (!nk::host_pack_body_available(body)) return fail(NK_ERR_UNSUPPORTED, "this CPU lacks the instruction set of body %d", body);
    if (nbases > 0 && (!bases || !codes)) return fail(NK_ERR_BAD_ARG, "null argument");
    const uint64_t n = nk::host_pack_bases(bases, nbases, codes, other, 1, body);
    if (n_other) *n_other = n;
    return NK_OK;
}

int nk_process_batch_packed(nk_counter* h, const uint32_t* codes, const uint32_t* other, const uint64_t* offsets,
                            uint64_t nseq) {
    if (is_group(h)) { NK_TRY(validate_packed(h, codes, offsets, nseq)); return group_process_batch(h, nullptr, codes, other, offsets, nseq); }
    NK_TRY(validate_packed(h, codes, offsets, nseq));
    if (h->streaming) return fail(NK_ERR_STATE, "nk_process_batch_packed inside nk_stream_begin/end");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    begin_call(h);
    PhaseEvents pe;
    NK_TRY(get_event(h, &pe.begin));
    NK_CUDA(cudaEventRecord(pe.begin, h->stream));
    NK_TRY(zero_kmers(h));
    h->currents_valid_overwrite = true;
    NK_TRY(count_host_batch_packed(h, codes, other, offsets, nseq, &pe, kPushSync));
    NK_TRY(fold_and_simulate(h, /*skip_zero=*/true, pe, true));
    NK_TRY(get_event(h, &pe.end));
    NK_CUDA(cudaEventRecord(pe.end, h->stream));
    NK_TRY(finish_call(h, true, &pe));
    return NK_OK;
}

int nk_stream_push_packed(nk_counter* h, const uint32_t* codes, const uint32_t* other, const uint64_t* offsets,
                          uint64_t nseq) {
    if (is_group(h)) { NK_TRY(validate_packed(h, codes, offsets, nseq)); return group_push(h, nullptr, codes, other, offsets, nseq); }
    NK_TRY(validate_packed(h, codes, offsets, nseq));
    if (!h->streaming) return fail(NK_ERR_STATE, "nk_stream_push_packed without nk_stream_begin");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    return count_host_batch_packed(h, codes, other, offsets, nseq, &h->stream_pe, kPushSync);
}

int nk_stream_accumulated(nk_counter* h, void** dev_currents) {
    if (is_group(h)) return group_unsupported("nk_stream_accumulated");
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (!h->streaming) return fail(NK_ERR_STATE, "nk_stream_accumulated without nk_stream_begin");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    NK_TRY(materialize_zero(h));
    NK_TRY(unspill(h));
    if (!h->acc_dirty && h->currents_valid_overwrite) {
        NK_CUDA(cudaMemsetAsync(h->currents, 0, h->cfg.pool_size * sizeof(unsigned long long), h->stream));
        h->currents_valid_overwrite = false;
    }
    NK_TRY(fold_now(h));
    // no host synchronisation: the caller's collective must be enqueued on nk_cuda_stream()
    // (or after nk_synchronize()), which orders it after the fold
    if (dev_currents) *dev_currents = h->currents;
    return NK_OK;
}

int nk_stream_finish(nk_counter* h) {
    if (is_group(h)) return group_end(h, /*skip_zero=*/false);
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (!h->streaming) return fail(NK_ERR_STATE, "nk_stream_finish without nk_stream_begin");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    PhaseEvents pe = h->stream_pe;
    NK_TRY(fold_and_simulate(h, /*skip_zero=*/false, pe, true));
    NK_TRY(finish_call(h, true, &pe));
    h->streaming = false;
    return NK_OK;
}

int nk_stream_end(nk_counter* h) { return nk_stream_finish(h); }

int nk_simulate(nk_counter* h) {
    if (is_group(h)) return group_simulate(h);
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (h->streaming) return fail(NK_ERR_STATE, "nk_simulate inside nk_stream_begin/end");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    begin_call(h);
    cudaEvent_t a, b;
    NK_TRY(get_event(h, &a));
    NK_TRY(get_event(h, &b));
    NK_CUDA(cudaEventRecord(a, h->stream));
    NK_TRY(simulate(h, /*skip_zero=*/false, false));
    NK_CUDA(cudaEventRecord(b, h->stream));
    PhaseEvents pe;
    pe.fold1 = a; pe.lif1 = b; pe.begin = a; pe.end = b;
    NK_TRY(finish_call(h, true, &pe));
    return NK_OK;
}

int nk_process_sequence(nk_counter* h, const uint8_t* seq, uint64_t len) {
    if (is_group(h)) return group_unsupported("nk_process_sequence (sequential per-sequence API: replicas only)");
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (len > 0 && !seq) return fail(NK_ERR_BAD_ARG, "null seq");
    if (h->streaming) return fail(NK_ERR_STATE, "nk_process_sequence inside nk_stream_begin/end");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    begin_call(h);
    if (len < h->cfg.k) return NK_OK;  // src/spiking_hash.rs:206-208
    // process_sequence ADDS to neuron_currents (fetch_add, :225,241,258) and zeroes them afterwards
    const uint64_t offs[2] = {0, len};
    NK_TRY(zero_kmers(h));
    h->currents_valid_overwrite = false;
    NK_TRY(materialize_zero(h));
    NK_TRY(count_host_batch(h, seq, offs, 1, nullptr, kPushSync));
    NK_TRY(fold_now(h));
    nk::LifParams p{};
    p.currents = h->currents; p.v = h->v; p.r = h->r; p.spikes = h->spikes;
    p.total_new = h->scalars + 0; p.max_spikes = h->scalars + 1;
    p.pool = h->cfg.pool_size; p.steps = 1; p.thr = h->cfg.threshold; p.leak = h->cfg.leak;
    p.period = h->cfg.refractory; p.skip_zero = 1;
    NK_CUDA(cudaMemsetAsync(h->scalars, 0, sizeof(unsigned long long), h->stream));
    h->fired_clean = false;
    NK_CUDA(nk::launch_lif_single_tick(p, h->currents, h->stream));
    ++h->last.launches;
    h->fresh = false;
    h->top_cache_valid = false;
    if (h->exact) NK_CUDA(nk::exact_finalize(h->xt, h->fm, h->cfg.pool_size, nullptr, true, h->stream));
    h->spike_bound = h->spike_bound + 1 ? h->spike_bound + 1 : h->spike_bound;
    return finish_call(h, true, nullptr);
}

int nk_top_n(nk_counter* h, uint64_t top_n, nk_top_entry* out, uint64_t* n_out) {
    if (is_group(h)) return group_top_n(h, top_n, out, n_out);
    if (!h || !n_out) return fail(NK_ERR_BAD_ARG, "null argument");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    uint64_t n = std::min<uint64_t>(top_n, h->cfg.pool_size);
    *n_out = 0;
    if (n == 0) return NK_OK;
    if (!out) return fail(NK_ERR_BAD_ARG, "null out");
    if (n > nk::TOPN_MAX_N) return fail(NK_ERR_UNSUPPORTED, "top_n > %llu not supported", nk::TOPN_MAX_N);
    if (n <= 2048 && n > h->topn_hint) h->topn_hint = n;  // later jobs' fused kernel computes at least this many rows
    if (h->top_cache_valid && n <= h->top_cached_n && !h->exact) {
        // rows already computed by the fused post kernel of the last job (sorted: a prefix is the top-n)
        NK_TRY(resolve(h));
        const unsigned long long cn = h->top_cached_n;
        for (uint64_t i = 0; i < n; ++i) {
            out[i].idx = h->h_pack[nk::PACK_HDR + i];
            out[i].spikes = h->h_pack[nk::PACK_HDR + cn + i];
            out[i].uniques = (h->rows_valid && i < h->row_idx.size() && h->row_idx[i] == out[i].idx) ? h->row_uniques[i]
                                                                                                     : NK_UNIQUES_NOT_COMPUTED;
            out[i]._pad = 0;
        }
        *n_out = n;
        h->last.topn_ms = 0.f;
        h->last.topn_launches = 0;
        return NK_OK;
    }
    unsigned long long cap = 1;
    while (cap < n) cap <<= 1;
    if (cap > h->topn_cap) {
        cudaFree(h->topn.out_idx); cudaFree(h->topn.out_spikes);
        if (h->h_top) cudaFreeHost(h->h_top);
        h->topn.out_idx = h->topn.out_spikes = nullptr; h->h_top = nullptr; h->topn_cap = 0;
        NK_CUDA(cudaMalloc(&h->topn.out_idx, cap * sizeof(unsigned long long)));
        NK_CUDA(cudaMalloc(&h->topn.out_spikes, cap * sizeof(unsigned long long)));
        NK_CUDA(cudaMallocHost(&h->h_top, 2 * cap * sizeof(unsigned long long)));
        h->topn_cap = cap;
    }
    if (h->slice_only)
        return fail(NK_ERR_STATE, "sharded-pool mode computed %llu rows; ask for more BEFORE the job (the hint is now %llu)",
                    h->top_cached_n, h->topn_hint);
    NK_TRY(materialize_zero(h));
    // the host-side bound on the largest cumulative spike count picks the radix passes: no round trip
    cudaEvent_t a, b;
    if (!h->pending) h->ev_used = 0;
    NK_TRY(get_event(h, &a));
    NK_TRY(get_event(h, &b));
    NK_CUDA(cudaEventRecord(a, h->stream));
    uint64_t launches = 0;
    NK_CUDA(nk::launch_topn(h->spikes, h->cfg.pool_size, n, h->spike_bound, h->topn, h->stream, &launches));
    NK_CUDA(cudaEventRecord(b, h->stream));
    std::vector<unsigned int> h_uni;
    const bool have_uniques = h->exact && h->xt.valid && h->xt.uniques;
    if (have_uniques) {
        cudaFree(h->d_top_uniques);
        h->d_top_uniques = nullptr;
        NK_CUDA(cudaMalloc(&h->d_top_uniques, n * sizeof(unsigned int)));
        NK_CUDA(nk::exact_gather_uniques(h->xt, h->topn.out_idx, n, h->d_top_uniques, h->stream));
        h_uni.resize(n);
        NK_CUDA(cudaMemcpyAsync(h_uni.data(), h->d_top_uniques, n * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    }
    unsigned long long* hi = reinterpret_cast<unsigned long long*>(h->h_top);
    unsigned long long* hs = hi + h->topn_cap;
    NK_CUDA(cudaMemcpyAsync(hi, h->topn.out_idx, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    NK_CUDA(cudaMemcpyAsync(hs, h->topn.out_spikes, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    NK_CUDA(cudaStreamSynchronize(h->stream));
    for (uint64_t i = 0; i < n; ++i) {
        out[i].idx = hi[i];
        out[i].spikes = hs[i];
        out[i].uniques = have_uniques ? h_uni[i]
                         : (h->rows_valid && i < h->row_idx.size() && h->row_idx[i] == hi[i]) ? h->row_uniques[i]
                                                                                             : NK_UNIQUES_NOT_COMPUTED;
        out[i]._pad = 0;
    }
    *n_out = n;
    NK_TRY(resolve(h));  // the synchronisation above also completed any in-flight read-back
    h->last.d2h_bytes += 2 * n * sizeof(unsigned long long);
    h->last.topn_ms = ev_ms(a, b);
    h->last.topn_launches = launches;
    return NK_OK;
}

int nk_total_spikes(const nk_counter* h, uint64_t* out) {
    if (is_group(h)) return nk_total_spikes(h->group[0], out);
    if (!h || !out) return fail(NK_ERR_BAD_ARG, "null argument");
    NK_TRY(resolve(const_cast<nk_counter*>(h)));
    *out = h->total_spikes;
    return NK_OK;
}
int nk_energy_used(const nk_counter* h, double* out) {
    if (is_group(h)) return nk_energy_used(h->group[0], out);
    if (!h || !out) return fail(NK_ERR_BAD_ARG, "null argument");
    NK_TRY(resolve(const_cast<nk_counter*>(h)));
    *out = (double)h->energy_fixed / 1000.0;  // src/models.rs:170-172
    return NK_OK;
}
int nk_enable_exact_counts(nk_counter* h, int on) {
    if (is_group(h)) return group_enable_exact(h, on);
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (h->streaming) return fail(NK_ERR_STATE, "nk_enable_exact_counts inside nk_stream_begin/end");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    NK_TRY(resolve(h));
    h->exact = on != 0;
    if (!h->exact) { NK_CUDA(cudaStreamSynchronize(h->stream)); nk::exact_free(h->xt); nk::exact_free(h->xs); }
    return NK_OK;
}

int nk_get_count(nk_counter* h, uint64_t kmer, uint32_t* count, int32_t* found) {
    if (is_group(h)) return group_get_count(h, kmer, count, found);
    if (!h || !count || !found) return fail(NK_ERR_BAD_ARG, "null argument");
    if (!h->exact)
        return fail(NK_ERR_UNSUPPORTED, "exact k-mer table is off: call nk_enable_exact_counts(h, 1) before processing");
    return exact_lookup_in(h, h->xt, kmer, count, found);
}

int nk_exact_table_size(nk_counter* h, uint64_t* n) {
    if (is_group(h)) return group_exact_table_size(h, n);
    if (!h || !n) return fail(NK_ERR_BAD_ARG, "null argument");
    if (!h->exact) return fail(NK_ERR_UNSUPPORTED, "exact k-mer table is off");
    NK_TRY(resolve(h));
    *n = h->xt.valid ? h->xt.n_keys : 0;
    return NK_OK;
}

int nk_copy_exact_table(nk_counter* h, uint64_t* keys, uint32_t* counts) {
    if (is_group(h)) return group_copy_exact_table(h, keys, counts);
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (!h->exact) return fail(NK_ERR_UNSUPPORTED, "exact k-mer table is off");
    return exact_copy_table_of(h, h->xt, keys, counts);
}

int nk_copy_uniques(nk_counter* h, uint32_t* out) {
    if (is_group(h)) return group_copy_uniques(h, out);
    if (!h || !out) return fail(NK_ERR_BAD_ARG, "null argument");
    if (!h->exact) return fail(NK_ERR_UNSUPPORTED, "exact k-mer table is off");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    NK_TRY(resolve(h));
    if (!h->xt.uniques) { std::memset(out, 0, h->cfg.pool_size * 4); return NK_OK; }
    NK_CUDA(cudaMemcpyAsync(out, h->xt.uniques, h->cfg.pool_size * 4, cudaMemcpyDeviceToHost, h->stream));
    NK_CUDA(cudaStreamSynchronize(h->stream));
    return NK_OK;
}

// seq != null: ASCII input; else the pre-packed form (codes, other)
static int debug_kmers_any(nk_counter* h, const uint8_t* seq, const uint32_t* codes, const uint32_t* other, uint64_t len,
                           uint64_t* fwd, uint64_t* rc, uint64_t* words, uint64_t* idx, uint64_t* n_out) {
    if (!h || !n_out) return fail(NK_ERR_BAD_ARG, "null argument");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    *n_out = 0;
    if (len < h->cfg.k) return NK_OK;
    if (!seq && !codes) return fail(NK_ERR_BAD_ARG, "null sequence pointer");
    const bool packed = seq == nullptr;
    NK_TRY(resolve(h));
    const uint64_t n = len - h->cfg.k + 1;
    NK_CUDA(cudaStreamSynchronize(h->stream));
    h->fp_valid = false;  // the staged buffers are about to be reused
    if (packed) NK_TRY(ensure_devbuf_packed(h->staged, len, true));
    else NK_TRY(ensure_devbuf(h->staged, len));
    NK_TRY(ensure_offsets(&h->staged_offsets, &h->staged_offsets_cap, 2));
    const uint64_t offs[2] = {0, len};
    unsigned long long* d_out = nullptr;
    NK_CUDA(cudaMalloc(&d_out, 4 * n * sizeof(unsigned long long)));
    int rc_ = NK_OK;
    do {
#define NK_D(expr) { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { rc_ = fail(NK_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); break; } }
        if (packed) {
            NK_D(cudaMemcpyAsync(h->staged.codes, codes, (len + 15) / 16 * 4, cudaMemcpyHostToDevice, h->stream));
            if (other) NK_D(cudaMemcpyAsync(h->staged.other, other, (len + 31) / 32 * 4, cudaMemcpyHostToDevice, h->stream));
        } else {
            NK_D(cudaMemcpyAsync(h->staged.bases, seq, len, cudaMemcpyHostToDevice, h->stream));
        }
        NK_D(cudaMemcpyAsync(h->staged_offsets, offs, sizeof offs, cudaMemcpyHostToDevice, h->stream));
        NK_D(cudaStreamSynchronize(h->stream));  // offs is on this stack frame
        NK_D(nk::launch_mark_invalid(h->staged.invalid, h->staged_offsets, 0, 1, 0, len, h->cfg.k, h->scalars + 3,
                                     h->stream, nullptr));
        nk::CountParams p{};
        p.bases = packed ? h->staged.codes : h->staged.bases;
        p.other = packed && other ? h->staged.other : nullptr;
        p.packed = packed ? 1 : 0;
        p.invalid = h->staged.invalid; p.acc = h->acc; p.tile_counter = h->tile_counter;
        p.ntiles = nk::count_ntiles(len); p.fm = h->fm; p.rm = nk::make_rotmul(); p.k = h->cfg.k;
        p.out_fwd = fwd ? d_out : nullptr;
        p.out_rc = rc ? d_out + n : nullptr;
        p.out_word = words ? d_out + 2 * n : nullptr;
        p.out_idx = idx ? d_out + 3 * n : nullptr;
        NK_D(nk::launch_count(p, h->cfg.use_canonical != 0, 1, h->stream));
        if (fwd) NK_D(cudaMemcpyAsync(fwd, d_out, n * 8, cudaMemcpyDeviceToHost, h->stream));
        if (rc) NK_D(cudaMemcpyAsync(rc, d_out + n, n * 8, cudaMemcpyDeviceToHost, h->stream));
        if (words) NK_D(cudaMemcpyAsync(words, d_out + 2 * n, n * 8, cudaMemcpyDeviceToHost, h->stream));
        if (idx) NK_D(cudaMemcpyAsync(idx, d_out + 3 * n, n * 8, cudaMemcpyDeviceToHost, h->stream));
        NK_D(cudaStreamSynchronize(h->stream));
#undef NK_D
    } while (0);
    cudaFree(d_out);
    if (rc_ == NK_OK) *n_out = n;
    return rc_;
}

int nk_debug_kmers(nk_counter* h, const uint8_t* seq, uint64_t len, uint64_t* fwd, uint64_t* rc, uint64_t* words,
                   uint64_t* idx, uint64_t* n_out) {
    if (is_group(h)) return nk_debug_kmers(h->group[0], seq, len, fwd, rc, words, idx, n_out);
    return debug_kmers_any(h, seq, nullptr, nullptr, len, fwd, rc, words, idx, n_out);
}

int nk_debug_kmers_packed(nk_counter* h, const uint32_t* codes, const uint32_t* other, uint64_t len, uint64_t* fwd,
                          uint64_t* rc, uint64_t* words, uint64_t* idx, uint64_t* n_out) {
    if (is_group(h)) return nk_debug_kmers_packed(h->group[0], codes, other, len, fwd, rc, words, idx, n_out);
    return debug_kmers_any(h, nullptr, codes, other, len, fwd, rc, words, idx, n_out);
}

int nk_debug_hash(nk_counter* h, const uint64_t* words, uint64_t n, uint64_t* hashes, uint64_t* idx) {
    if (is_group(h)) return nk_debug_hash(h->group[0], words, n, hashes, idx);
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (n == 0) return NK_OK;
    if (!words) return fail(NK_ERR_BAD_ARG, "null words");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    unsigned long long* d = nullptr;
    NK_CUDA(cudaMalloc(&d, 3 * n * sizeof(unsigned long long)));
    int rc_ = NK_OK;
    do {
#define NK_D(expr) { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { rc_ = fail(NK_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); break; } }
        NK_D(cudaMemcpyAsync(d, words, n * 8, cudaMemcpyHostToDevice, h->stream));
        NK_D(nk::launch_hash_words(d, n, h->fm, d + n, d + 2 * n, h->stream));
        if (hashes) NK_D(cudaMemcpyAsync(hashes, d + n, n * 8, cudaMemcpyDeviceToHost, h->stream));
        if (idx) NK_D(cudaMemcpyAsync(idx, d + 2 * n, n * 8, cudaMemcpyDeviceToHost, h->stream));
        NK_D(cudaStreamSynchronize(h->stream));
#undef NK_D
    } while (0);
    cudaFree(d);
    return rc_;
}

int nk_debug_mod(const uint64_t* values, uint64_t n, uint64_t pool_size, int which, uint64_t* out) {
    if (n == 0) return NK_OK;
    if (!values || !out) return fail(NK_ERR_BAD_ARG, "null argument");
    if (pool_size == 0 || pool_size >= (1ull << 32)) return fail(NK_ERR_BAD_ARG, "pool_size must be in [1, 2^32)");
    if (which < 0 || which > 2)
        return fail(NK_ERR_BAD_ARG, "which must be 0 (two-stage FP64 form), 1 (integer form) or 2 (the count kernel's pick)");
    unsigned long long* d = nullptr;
    NK_CUDA(cudaMalloc(&d, 2 * n * sizeof(unsigned long long)));
    cudaError_t e = cudaMemcpy(d, values, n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = nk::launch_mod_words(d, n, nk::make_fastmod(pool_size), d + n, which, nullptr);
    if (e == cudaSuccess) e = cudaMemcpy(out, d + n, n * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    NK_CUDA(e);
    return NK_OK;
}

}  // extern "C"
namespace nkd {
int copy_out(nk_counter* h, void* dst, const void* src, size_t bytes) {
    if (!h || !dst) return fail(NK_ERR_BAD_ARG, "null argument");
    NK_TRY(resolve(h));
    NK_TRY(materialize_zero(h));
    NK_CUDA(cudaSetDevice(h->cfg.device));
    NK_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
    NK_CUDA(cudaStreamSynchronize(h->stream));
    return NK_OK;
}
}  // namespace nkd
extern "C" {
int nk_copy_currents(nk_counter* h, uint64_t* out) { if (is_group(h)) return group_copy(h, 0, out); return copy_out(h, out, h ? h->currents : nullptr, h ? h->cfg.pool_size * 8 : 0); }
int nk_copy_spike_counts(nk_counter* h, uint64_t* out) { if (is_group(h)) return group_copy(h, 1, out); return copy_out(h, out, h ? h->spikes : nullptr, h ? h->cfg.pool_size * 8 : 0); }
int nk_copy_voltages(nk_counter* h, float* out) { if (is_group(h)) return group_copy(h, 2, out); return copy_out(h, out, h ? h->v : nullptr, h ? h->cfg.pool_size * 4 : 0); }
int nk_copy_refractory(nk_counter* h, uint32_t* out) { if (is_group(h)) return group_copy(h, 3, out); return copy_out(h, out, h ? h->r : nullptr, h ? h->cfg.pool_size * 4 : 0); }

int nk_last_timings(const nk_counter* h, nk_timings* out) {
    if (is_group(h)) return out ? group_timings(const_cast<nk_counter*>(h), out) : fail(NK_ERR_BAD_ARG, "null argument");
    if (!h || !out) return fail(NK_ERR_BAD_ARG, "null argument");
    NK_TRY(resolve(const_cast<nk_counter*>(h)));
    *out = h->last;
    return NK_OK;
}

int nk_debug_set_lif_path(nk_counter* h, int mode) {
    if (is_group(h)) { for (nk_counter* c : h->group) NK_TRY(nk_debug_set_lif_path(c, mode)); return NK_OK; }
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (mode < 0 || mode > 2) return fail(NK_ERR_BAD_ARG, "mode must be 0, 1 or 2");
    h->force_direct = mode;
    return NK_OK;
}

int nk_debug_set_fold_limit(nk_counter* h, uint64_t limit) {
    if (is_group(h)) { for (nk_counter* c : h->group) NK_TRY(nk_debug_set_fold_limit(c, limit)); return NK_OK; }
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (limit == 0 || limit > 0xFFFFFFFFull) return fail(NK_ERR_BAD_ARG, "limit must be in 1..2^32-1");
    h->fold_limit = limit;
    return NK_OK;
}

int nk_calibrate(nk_counter* h, int which, double* out) {
    if (is_group(h)) return nk_calibrate(h->group[0], which, out);
    if (!h || !out) return fail(NK_ERR_BAD_ARG, "null argument");
    if (which < 0 || which > 2) return fail(NK_ERR_BAD_ARG, "which must be 0, 1 or 2");
    if (h->streaming || h->acc_dirty) return fail(NK_ERR_STATE, "nk_calibrate needs an idle counter");
    NK_TRY(resolve(h));
    NK_CUDA(cudaSetDevice(h->cfg.device));
    int sms = 0;
    NK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->cfg.device));
    const int blocks = sms * 8;
    cudaEvent_t a, b;
    h->ev_used = 0;
    NK_TRY(get_event(h, &a));
    NK_TRY(get_event(h, &b));
    double best = 0.0;
    // which == 2: the addresses of real k-mer traffic (SipHash-1-3 of consecutive words % pool), precomputed
    const unsigned long long n_idx = 64ull << 20;
    unsigned int* d_idx = nullptr;
    if (which == 2) {
        NK_CUDA(cudaMalloc(&d_idx, n_idx * sizeof(unsigned int)));
        cudaError_t e = nk::launch_hashed_idx(d_idx, n_idx, 0x5EEDull, h->fm, h->stream);
        if (e != cudaSuccess) { cudaFree(d_idx); NK_CUDA(e); }
    }
    for (int rep = 0; rep < 4; ++rep) {
        double work = 0.0;
        NK_CUDA(cudaEventRecord(a, h->stream));
        if (which <= 1) {
            const unsigned iters = 4096;
            NK_CUDA(nk::launch_int_peak(which, h->tile_counter + 4, blocks, iters, h->stream));
            work = (double)blocks * 256.0 * iters * (double)nk::int_peak_ops_per_iter(which);
        } else {
            cudaError_t e = nk::launch_red_peak(h->acc, d_idx, n_idx, blocks, h->stream);
            if (e != cudaSuccess) { cudaFree(d_idx); NK_CUDA(e); }
            work = (double)n_idx;
        }
        NK_CUDA(cudaEventRecord(b, h->stream));
        NK_CUDA(cudaStreamSynchronize(h->stream));
        const float ms = ev_ms(a, b);
        if (rep > 0 && ms > 0.f) best = std::max(best, work / (ms * 1e-3));
    }
    if (which == 2) NK_CUDA(cudaMemsetAsync(h->acc, 0, h->cfg.pool_size * sizeof(unsigned int), h->stream));
    NK_CUDA(cudaStreamSynchronize(h->stream));
    if (d_idx) cudaFree(d_idx);
    *out = best;
    return NK_OK;
}

// ---------------------------------------------------------------------------
// Multi-GPU, sharded pool: "reduce-scatter fused into the LIF kernel" over NVLink peer memory.
// Every rank counts its shard into its own full u32 accumulator array; rank r then owns the
// neuron slice [lo_r, hi_r) and its post kernel reads that slice of EVERY rank's accumulators
// straight through peer mappings (CUDA IPC), sums, runs LIF + a slice-local top-N, and emits a
// small result pack; the packs are all-gathered (a few hundred bytes) and merged on the device.
// Against all-reduce(16 MB) + full-pool LIF + full-pool top-N on every rank this moves 1/world of
// the bytes and does 1/world of the per-neuron work.  Fresh-state (table) path only: otherwise
// nk_dist_post returns NK_ERR_UNSUPPORTED and the caller uses nk_stream_accumulated + all-reduce.
// ---------------------------------------------------------------------------
int nk_dist_export(nk_counter* h, void* handle64, void** raw_acc) {
    if (is_group(h)) return group_unsupported("nk_dist_export");
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    if (handle64) {
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        cudaIpcMemHandle_t mh;
        NK_CUDA(cudaIpcGetMemHandle(&mh, h->acc));
        std::memcpy(handle64, &mh, 64);
    }
    if (raw_acc) *raw_acc = h->acc;
    return NK_OK;
}

int nk_dist_setup(nk_counter* h, int rank, int world, const void* handles, void* const* raw_ptrs) {
    if (is_group(h)) return group_unsupported("nk_dist_setup");
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (world < 1 || world > 16 || rank < 0 || rank >= world) return fail(NK_ERR_BAD_ARG, "bad rank/world (world <= 16)");
    if (!handles && !raw_ptrs) return fail(NK_ERR_BAD_ARG, "need IPC handles or raw pointers");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    NK_CUDA(cudaStreamSynchronize(h->stream));
    for (int r = 0; r < 16; ++r) {  // a repeated setup replaces the previous mappings
        if (h->dist_ipc_opened[r]) cudaIpcCloseMemHandle((void*)h->dist_peer[r]);
        h->dist_ipc_opened[r] = false;
        h->dist_peer[r] = nullptr;
    }
    for (int r = 0; r < world; ++r) {
        if (r == rank) { h->dist_peer[r] = h->acc; continue; }
        if (raw_ptrs) { h->dist_peer[r] = (const unsigned int*)raw_ptrs[r]; continue; }
        cudaIpcMemHandle_t mh;
        std::memcpy(&mh, (const char*)handles + 64 * r, 64);
        void* p = nullptr;
        NK_CUDA(cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
        h->dist_peer[r] = (const unsigned int*)p;
        h->dist_ipc_opened[r] = true;
    }
    h->dist_rank = rank;
    h->dist_world = world;
    const unsigned long long mail_off = nk::dist_mail_offset(h->cfg.pool_size);
    for (int r = 0; r < 16; ++r)
        h->dist_mail[r] = r < world ? reinterpret_cast<unsigned char*>(const_cast<unsigned int*>(h->dist_peer[r])) + mail_off : nullptr;
    // epochs restart with every setup: the setup is a collective call, and a cross-rank barrier must
    // separate it from the first nk_dist_run (peers write into this mailbox from then on)
    NK_CUDA(cudaMemset(h->dist_mail[rank], 0, nk::DIST_MAIL_BYTES));
    h->dist_epoch = 0;
    h->dist_failed = false;
    const unsigned long long P = h->cfg.pool_size, per = (P + world - 1) / world;
    h->dist_lo = std::min(P, per * rank);
    h->dist_len = std::min(P, h->dist_lo + per) - h->dist_lo;
    if (!h->d_merged) NK_CUDA(cudaMalloc(&h->d_merged, nk::PACK_MAX_U64 * sizeof(unsigned long long)));
    return NK_OK;
}

}  // extern "C"

namespace nkd {

// parameters of this rank's slice post kernel (shared by nk_dist_post and nk_dist_run)
int dist_build_post(nk_counter* h, nk::PostParams& q, unsigned long long* n_top_out) {
    const unsigned long long per = (h->cfg.pool_size + h->dist_world - 1) / h->dist_world;
    const unsigned long long n_top = std::min<unsigned long long>(h->topn_hint, per);
    const bool table_ok = h->fresh && !h->force_direct && !h->exact && h->cfg.steps > 0 && std::isfinite(h->cfg.threshold) &&
                          std::isfinite(h->cfg.leak) && saturation_count(h->cfg) < (1ull << 20);
    if (!table_ok || n_top < 1 || n_top * h->dist_world > 2048 || h->dist_len == 0)
        return fail(NK_ERR_UNSUPPORTED, "sharded post needs the fresh-state table path, world*top_n <= 2048 and a non-empty slice");
    nk::LifParams p{};
    p.currents = h->currents + h->dist_lo;
    p.acc = h->acc;
    p.fold_mode = 2;
    p.zero_state = 1;  // the slice's v, r, spikes are written, never read (fresh state)
    p.v = h->v + h->dist_lo;
    p.r = h->r + h->dist_lo;
    p.spikes = h->spikes + h->dist_lo;
    p.total_new = h->scalars + 0;
    p.max_spikes = h->scalars + 1;
    p.pool = h->dist_len;
    p.steps = h->cfg.steps;
    p.thr = h->cfg.threshold;
    p.leak = h->cfg.leak;
    p.period = h->cfg.refractory;
    p.skip_zero = 0;
    // the table (depends on the LIF parameters only)
    const unsigned long long table_n = saturation_count(h->cfg) + 1;
    NK_TRY(ensure_table(h, table_n, h->stream));  // usually prelaunched by nk_stream_begin on the side stream
    if (h->table_inflight) {
        NK_CUDA(cudaStreamWaitEvent(h->stream, h->table_ready, 0));
        h->table_inflight = false;
    }
    const unsigned long long per_call = (h->cfg.steps + h->cfg.refractory) / ((unsigned long long)h->cfg.refractory + 1ull);
    q = nk::PostParams{};
    q.lif = p;
    q.table = h->table;
    q.table_n = table_n;
    q.n = std::min<unsigned long long>(n_top, h->dist_len);
    int bits = 0;
    while (bits < 64 && (per_call >> bits)) ++bits;
    q.passes = std::max(1, (bits + 7) / 8);
    q.single_pass = per_call < (unsigned long long)nk::POST_EXACT_BINS ? 1 : 0;
    q.ctrl = h->post_zero;
    q.hist = reinterpret_cast<unsigned int*>(h->post_zero + 8);
    q.block_ties = q.hist + 8 * 256;
    q.seg_counts = h->topn.block_counts;
    q.out_idx = h->topn.out_idx;
    q.out_spikes = h->topn.out_spikes;
    q.pack = h->d_pack;
    q.kmers = h->scalars + 2;
    q.npeers = h->dist_world;
    const unsigned long long spill_off = nk::dist_spill_offset(h->cfg.pool_size);
    for (int r = 0; r < h->dist_world; ++r) {
        q.peer_acc[r] = h->dist_peer[r];
        q.peer_spill[r] = reinterpret_cast<const unsigned long long*>(reinterpret_cast<const unsigned char*>(h->dist_peer[r]) + spill_off);
        q.peer_mail[r] = h->dist_mail[r];
    }
    q.slice_lo = h->dist_lo;
    q.rank = h->dist_rank;
    // rows are padded to n_top per rank so that every rank's pack has the same size
    NK_CUDA(cudaMemsetAsync(h->d_pack, 0, (nk::PACK_HDR + 2 * n_top) * sizeof(unsigned long long), h->stream));
    if (!h->fired_clean) NK_CUDA(cudaMemsetAsync(h->scalars, 0, sizeof(unsigned long long), h->stream));
    h->fired_clean = h->kmers_clean = true;  // the slice kernel every caller launches next leaves them zeroed
    *n_top_out = n_top;
    h->dist_n_each = n_top;
    return NK_OK;
}

// where the merge kernel of a sharded job writes the merged pack: the mapped pinned host buffer when there is one
// (no D2H copy operation behind the kernel), else device memory that dist_finish copies back
unsigned long long* merged_pack_out(nk_counter* h) {
    h->pack_direct = h->h_pack_dev != nullptr;
    return h->pack_direct ? h->h_pack_dev : h->d_merged;
}

// host-side state after a sharded job's merged pack has been queued for read-back
int dist_finish(nk_counter* h, unsigned long long n_out) {
    // accumulators are all-zero between jobs (every peer has read them by now)
    NK_CUDA(cudaMemsetAsync(h->acc, 0, h->cfg.pool_size * sizeof(unsigned int), h->stream));
    if (h->spill_dirty) {  // the spill array itself is overwritten by its next use
        NK_CUDA(cudaMemsetAsync(nk::dist_mail_flags(own_mail(h), 2), 0, 8, h->stream));
        h->spill_dirty = false;
    }
    h->dist_job = true;
    h->slice_only = true;
    const size_t bytes = (nk::PACK_HDR + 2 * n_out) * sizeof(unsigned long long);
    if (!h->pack_direct) NK_CUDA(cudaMemcpyAsync(h->h_pack, h->d_merged, bytes, cudaMemcpyDeviceToHost, h->stream));
    h->pack_direct = false;
    h->last.d2h_bytes += bytes;   // (written by the merge kernel itself when the pack lives in mapped host memory)
    h->acc_dirty = false;
    h->acc_kmers = 0;
    h->currents_valid_overwrite = false;
    h->lazy_zero = false;  // NOTE: only this rank's slice of currents/v/r/spikes is defined in this mode
    h->fresh = false;
    h->spike_bound += (h->cfg.steps + h->cfg.refractory) / ((unsigned long long)h->cfg.refractory + 1ull);
    h->top_cached_n = n_out;
    h->top_cache_valid = true;
    h->pending_pack = true;
    h->pending = true;
    h->pending_lif = true;
    h->pending_timings = true;
    h->pend_pe = h->stream_pe;
    h->streaming = false;
    return NK_OK;
}

}  // namespace nkd

extern "C" {

// Enqueue this rank's slice post kernel.  PRECONDITION (caller): every rank's counting is complete
// and ordered before this call on the handle's stream (e.g. a 1-element NCCL all-reduce on it).
int nk_dist_post(nk_counter* h, void** dev_pack, uint64_t* pack_u64s, uint64_t* n_each) {
    if (is_group(h)) return group_unsupported("nk_dist_post");
    if (!h || !dev_pack || !pack_u64s || !n_each) return fail(NK_ERR_BAD_ARG, "null argument");
    if (!h->dist_world) return fail(NK_ERR_STATE, "nk_dist_setup was not called");
    if (!h->streaming) return fail(NK_ERR_STATE, "nk_dist_post without nk_stream_begin");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    nk::PostParams q{};
    unsigned long long n_top = 0;
    NK_TRY(dist_build_post(h, q, &n_top));
    NK_CUDA(nk::launch_post(q, h->post_grid, h->stream));
    ++h->last.launches;
    h->last.lif_path = 4;
    *dev_pack = h->d_pack;
    *pack_u64s = nk::PACK_HDR + 2 * n_top;
    *n_each = n_top;
    return NK_OK;
}

// The whole exchange with NO host or NCCL synchronisation per job: the ranks signal each other through
// flags in peer memory (NVLink stores), the post kernel waits for "everyone finished counting" before it
// reads the peers' counts and delivers its result pack into every rank's mailbox, and a one-block kernel
// waits for all packs and merges them.  One handle per GPU (the waiting kernels of two handles on one GPU
// could starve each other); a lost peer surfaces as NK_ERR_STATE after NK_DIST_TIMEOUT_MS (default 30 s).
int nk_dist_run(nk_counter* h) {
    if (is_group(h)) return group_unsupported("nk_dist_run");
    NvtxRange nvtx("nk:exchange (signal + slice reduce/LIF/top-N + pack merge)");
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (!h->dist_world) return fail(NK_ERR_STATE, "nk_dist_setup was not called");
    if (!h->streaming) return fail(NK_ERR_STATE, "nk_dist_run without nk_stream_begin");
    if (h->dist_failed) return fail(NK_ERR_STATE, "an earlier multi-GPU job timed out: call nk_dist_setup again on every rank");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    unsigned long long timeout_ms = 30000;
    if (const char* e = getenv("NK_DIST_TIMEOUT_MS")) {
        const unsigned long long t = strtoull(e, nullptr, 10);
        if (t >= 1) timeout_ms = t;
    }
    nk::PostParams q{};
    unsigned long long n_top = 0;
    NK_TRY(dist_build_post(h, q, &n_top));  // validates before any signal is sent
    const unsigned long long epoch = ++h->dist_epoch;
    // "this rank finished counting": ordered after its count kernels on the stream
    NK_CUDA(nk::launch_dist_signal(h->dist_mail, h->dist_world, h->dist_rank, 0, epoch, h->stream));
    ++h->last.launches;
    q.wait_flags = nk::dist_mail_flags(h->dist_mail[h->dist_rank], 0);
    q.epoch = epoch;
    q.timeout_ns = timeout_ms * 1000000ull;
    NK_CUDA(nk::launch_post(q, h->post_grid, h->stream));
    ++h->last.launches;
    h->last.lif_path = 4;
    const unsigned long long n_out = std::min<unsigned long long>(h->topn_hint, h->cfg.pool_size);
    NK_CUDA(nk::launch_merge_mailbox(h->dist_mail[h->dist_rank], h->dist_world, h->dist_rank, n_top, n_out, epoch, q.timeout_ns,
                                     merged_pack_out(h), h->stream));
    ++h->last.launches;
    return dist_finish(h, n_out);
}

// `dev_gathered`: the world packs, rank-major, all-gathered by the caller on the handle's stream
// (which also proves that every rank has finished reading this rank's accumulators).
int nk_dist_complete(nk_counter* h, const void* dev_gathered, uint64_t n_each) {
    if (is_group(h)) return group_unsupported("nk_dist_complete");
    if (!h || !dev_gathered) return fail(NK_ERR_BAD_ARG, "null argument");
    if (!h->dist_world || !h->streaming) return fail(NK_ERR_STATE, "nk_dist_complete without nk_dist_post");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    const unsigned long long n_out = std::min<unsigned long long>(h->topn_hint, h->cfg.pool_size);
    NK_CUDA(nk::launch_merge_packs((const unsigned long long*)dev_gathered, h->dist_world, h->dist_rank, n_each, n_out, merged_pack_out(h), h->stream));
    ++h->last.launches;
    return dist_finish(h, n_out);
}

int nk_dist_slice(const nk_counter* h, uint64_t* lo, uint64_t* len) {
    if (is_group(h)) return group_unsupported("nk_dist_slice");
    if (!h || !lo || !len) return fail(NK_ERR_BAD_ARG, "null argument");
    *lo = h->dist_lo;
    *len = h->dist_len;
    return NK_OK;
}

int nk_stage_reserve(nk_counter* h, uint64_t nbytes, uint64_t nseq, void** dev_bases, void** dev_offsets) {
    if (is_group(h)) return group_unsupported("nk_stage_reserve");
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    NK_TRY(resolve(h));
    NK_CUDA(cudaStreamSynchronize(h->stream));
    h->fp_valid = false;
    NK_TRY(ensure_devbuf(h->staged, nbytes));
    NK_TRY(ensure_offsets(&h->staged_offsets, &h->staged_offsets_cap, nseq + 1));
    if (dev_bases) *dev_bases = h->staged.bases;
    if (dev_offsets) *dev_offsets = h->staged_offsets;
    return NK_OK;
}

int nk_process_staged(nk_counter* h, uint64_t nbytes, uint64_t nseq, int mode) {
    if (is_group(h)) return group_unsupported("nk_process_staged");
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (mode != 0 && mode != 1) return fail(NK_ERR_BAD_ARG, "mode must be 0 or 1");
    if (mode == 1 && !h->streaming) return fail(NK_ERR_STATE, "mode 1 needs nk_stream_begin");
    if (mode == 0 && h->streaming) return fail(NK_ERR_STATE, "mode 0 inside nk_stream_begin/end");
    if (nk::count_padded_bases(nbytes) > h->staged.bases_cap || nseq + 1 > h->staged_offsets_cap)
        return fail(NK_ERR_STATE, "staged batch larger than the last nk_stage_reserve");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    PhaseEvents pe;
    if (mode == 0) {
        begin_call(h);
        NK_TRY(get_event(h, &pe.begin));
        NK_CUDA(cudaEventRecord(pe.begin, h->stream));
        NK_TRY(zero_kmers(h));
        h->currents_valid_overwrite = true;
    }
    // the device-resident batch is counted in slices of < 2^32 window starts (one launch each; the
    // u32 accumulators are folded between slices when they could overflow).  Every slice passes the
    // whole sequence range: the marking kernel clips each sequence to the slice.
    const unsigned long long slice = (0xFFFFFFFFull / nk::COUNT_TILE - 1) * nk::COUNT_TILE;
    for (unsigned long long c0 = 0; c0 < nbytes && nseq > 0; c0 += slice) {
        const unsigned long long n = std::min(slice, nbytes - c0);
        DevBuf view = h->staged;
        view.bases = h->staged.bases + c0;
        NK_TRY(count_chunk(h, view, h->staged_offsets, 0, nseq, c0, n, n, &pe, false));
    }
    if (mode == 0) {
        NK_TRY(fold_and_simulate(h, /*skip_zero=*/true, pe, true));
        NK_TRY(get_event(h, &pe.end));
        NK_CUDA(cudaEventRecord(pe.end, h->stream));
        NK_TRY(finish_call(h, true, &pe));
    } else {
        // the mark/count event pairs of this push are collected with the stream's final read-back
        h->stream_pe.mark0.insert(h->stream_pe.mark0.end(), pe.mark0.begin(), pe.mark0.end());
        h->stream_pe.count0.insert(h->stream_pe.count0.end(), pe.count0.begin(), pe.count0.end());
        h->stream_pe.count1.insert(h->stream_pe.count1.end(), pe.count1.begin(), pe.count1.end());
    }
    return NK_OK;
}

int nk_stage_reserve_packed(nk_counter* h, uint64_t nbases, uint64_t nseq, void** dev_codes, void** dev_other,
                            void** dev_offsets) {
    if (is_group(h)) return group_unsupported("nk_stage_reserve_packed");
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    NK_TRY(resolve(h));
    NK_CUDA(cudaStreamSynchronize(h->stream));
    h->fp_valid = false;
    NK_TRY(ensure_devbuf_packed(h->staged, nbases, true));
    NK_TRY(ensure_offsets(&h->staged_offsets, &h->staged_offsets_cap, nseq + 1));
    if (dev_codes) *dev_codes = h->staged.codes;
    if (dev_other) *dev_other = h->staged.other;
    if (dev_offsets) *dev_offsets = h->staged_offsets;
    return NK_OK;
}

int nk_process_staged_packed(nk_counter* h, uint64_t nbases, uint64_t nseq, int mode, int has_other) {
    if (is_group(h)) return group_unsupported("nk_process_staged_packed");
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (mode != 0 && mode != 1) return fail(NK_ERR_BAD_ARG, "mode must be 0 or 1");
    if (mode == 1 && !h->streaming) return fail(NK_ERR_STATE, "mode 1 needs nk_stream_begin");
    if (mode == 0 && h->streaming) return fail(NK_ERR_STATE, "mode 0 inside nk_stream_begin/end");
    if (nk::count_padded_codes(nbases) > h->staged.codes_cap || nk::count_padded_other(nbases) > h->staged.other_cap ||
        nseq + 1 > h->staged_offsets_cap)
        return fail(NK_ERR_STATE, "staged batch larger than the last nk_stage_reserve_packed");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    PhaseEvents pe;
    if (mode == 0) {
        begin_call(h);
        NK_TRY(get_event(h, &pe.begin));
        NK_CUDA(cudaEventRecord(pe.begin, h->stream));
        NK_TRY(zero_kmers(h));
        h->currents_valid_overwrite = true;
    }
    const unsigned long long slice = (0xFFFFFFFFull / nk::COUNT_TILE - 1) * nk::COUNT_TILE;
    for (unsigned long long c0 = 0; c0 < nbases && nseq > 0; c0 += slice) {
        const unsigned long long n = std::min(slice, nbases - c0);
        DevBuf view = h->staged;
        view.codes = h->staged.codes + c0 / 4;
        view.other = h->staged.other + c0 / 8;
        view.has_other = has_other != 0;
        NK_TRY(count_chunk(h, view, h->staged_offsets, 0, nseq, c0, n, n, &pe, true));
    }
    if (mode == 0) {
        NK_TRY(fold_and_simulate(h, /*skip_zero=*/true, pe, true));
        NK_TRY(get_event(h, &pe.end));
        NK_CUDA(cudaEventRecord(pe.end, h->stream));
        NK_TRY(finish_call(h, true, &pe));
    } else {
        h->stream_pe.mark0.insert(h->stream_pe.mark0.end(), pe.mark0.begin(), pe.mark0.end());
        h->stream_pe.count0.insert(h->stream_pe.count0.end(), pe.count0.begin(), pe.count0.end());
        h->stream_pe.count1.insert(h->stream_pe.count1.end(), pe.count1.begin(), pe.count1.end());
    }
    return NK_OK;
}

int nk_cuda_stream(nk_counter* h, void** stream) {
    if (is_group(h)) return nk_cuda_stream(h->group[0], stream);
    if (!h || !stream) return fail(NK_ERR_BAD_ARG, "null argument");
    *stream = (void*)h->stream;
    return NK_OK;
}
int nk_synchronize(nk_counter* h) {
    if (is_group(h)) return group_synchronize(h);
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    NK_CUDA(cudaStreamSynchronize(h->copy_stream));
    NK_CUDA(cudaStreamSynchronize(h->stream));
    return resolve(h);
}

int nk_synth_fill(nk_counter* h, void* dev_out, uint64_t seed, uint64_t start, uint64_t n, uint32_t flags) {
    if (is_group(h)) return group_unsupported("nk_synth_fill");
    if (!h || (!dev_out && n)) return fail(NK_ERR_BAD_ARG, "null argument");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    NK_CUDA(nk::launch_synth((unsigned char*)dev_out, seed, start, n, flags, h->stream));
    return NK_OK;
}

int nk_host_alloc(void** ptr, uint64_t nbytes) {
    if (!ptr) return fail(NK_ERR_BAD_ARG, "null ptr");
    *ptr = nullptr;
    cudaError_t e = cudaMallocHost(ptr, nbytes ? nbytes : 1);
    if (e != cudaSuccess) return fail(NK_ERR_OOM, "cudaMallocHost(%llu): %s", (unsigned long long)nbytes, cudaGetErrorString(e));
    return NK_OK;
}
int nk_host_free(void* ptr) {
    if (ptr) NK_CUDA(cudaFreeHost(ptr));
    return NK_OK;
}

uint64_t nk_pack_kmer(const uint8_t* kmer, uint64_t len) { return nk::host_pack_kmer(kmer, len); }

// ---- uniques of the top rows by a second pass over the input (no O(windows) table) -------------------
static int uniques_begin(nk_counter* h, uint64_t top_n) {
    if (h->streaming) return fail(NK_ERR_STATE, "nk_uniques_begin inside nk_stream_begin/end");
    if (h->uniques_open) return fail(NK_ERR_STATE, "nk_uniques_begin called twice");
    if (h->slice_only && !h->uniques_whole_input)
        return fail(NK_ERR_UNSUPPORTED, "uniques pass after a sharded-pool job: every rank holds only its shard of the input "
                    "(the per-rank word sets are not merged)");
    const uint64_t n = std::min<uint64_t>(top_n, h->cfg.pool_size);
    if (n == 0 || n > 2048) return fail(NK_ERR_BAD_ARG, "nk_uniques_begin: top_n must be in 1..2048");
    std::vector<nk_top_entry> rows(n);
    uint64_t got = 0;
    NK_TRY(nk_top_n(h, n, rows.data(), &got));
    NK_CUDA(cudaStreamSynchronize(h->stream));
    h->rows_valid = false;
    h->row_idx.resize(got);
    for (uint64_t i = 0; i < got; ++i) h->row_idx[i] = rows[i].idx;
    h->row_uniques.assign(got, 0u);
    const unsigned long long fwords = (h->cfg.pool_size + 31) / 32;
    if (!h->d_filter) NK_CUDA(cudaMalloc(&h->d_filter, fwords * sizeof(unsigned int)));
    if (got > h->d_rows_cap) {
        cudaFree(h->d_rows);
        h->d_rows = nullptr;
        h->d_rows_cap = 0;
        NK_CUDA(cudaMalloc(&h->d_rows, got * sizeof(unsigned long long)));
        h->d_rows_cap = got;
    }
    NK_CUDA(cudaMemsetAsync(h->d_filter, 0, fwords * sizeof(unsigned int), h->stream));
    NK_CUDA(cudaMemcpyAsync(h->d_rows, h->row_idx.data(), got * sizeof(unsigned long long), cudaMemcpyHostToDevice, h->stream));
    NK_CUDA(nk::launch_filter_set(h->d_filter, h->d_rows, got, h->stream));
    NK_CUDA(nk::exact_grow_words(h->ut, 1ull << 20, 0, h->stream));
    NK_CUDA(nk::exact_clear(h->ut, h->cfg.pool_size, true, h->stream));
    NK_CUDA(cudaStreamSynchronize(h->stream));
    h->ut_count = 0;
    h->uniques_open = true;
    return NK_OK;
}

static int uniques_end(nk_counter* h) {
    h->uniques_open = false;
    cudaError_t e = nk::exact_finalize(h->ut, h->fm, h->cfg.pool_size, nullptr, false, h->stream);
    if (e == cudaErrorInvalidValue) return fail(NK_ERR_UNSUPPORTED, "uniques pass: the collected words could not be partitioned");
    NK_CUDA(e);
    const uint64_t n = h->row_idx.size();
    if (n) {
        cudaFree(h->d_top_uniques);
        h->d_top_uniques = nullptr;
        NK_CUDA(cudaMalloc(&h->d_top_uniques, n * sizeof(unsigned int)));
        NK_CUDA(nk::exact_gather_uniques(h->ut, h->d_rows, n, h->d_top_uniques, h->stream));
        NK_CUDA(cudaMemcpyAsync(h->row_uniques.data(), h->d_top_uniques, n * sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
        NK_CUDA(cudaStreamSynchronize(h->stream));
    }
    h->ut.valid = false;  // not an exact table: nk_get_count must not answer from it
    h->rows_valid = true;
    return NK_OK;
}

int nk_uniques_begin(nk_counter* h, uint64_t top_n) {
    if (is_group(h)) {
        // the second pass runs on group[0], which is handed the WHOLE input again (the rows are the merged ones)
        if (top_n == 0 || top_n > 2048) return fail(NK_ERR_BAD_ARG, "nk_uniques_begin: top_n must be in 1..2048");
        std::vector<nk_top_entry> rows(top_n);
        uint64_t got = 0;
        NK_TRY(group_top_n(h, top_n, rows.data(), &got));  // brings the state to group[0] if the slices computed fewer rows
        nk_counter* c0 = h->group[0];
        NK_CUDA(cudaSetDevice(c0->cfg.device));
        c0->uniques_whole_input = true;
        return uniques_begin(c0, top_n);
    }
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    return uniques_begin(h, top_n);
}

int nk_uniques_push(nk_counter* h, const uint8_t* bases, const uint64_t* offsets, uint64_t nseq) {
    if (is_group(h)) return nk_uniques_push(h->group[0], bases, offsets, nseq);
    NK_TRY(validate_batch(h, bases, offsets, nseq));
    if (!h->uniques_open) return fail(NK_ERR_STATE, "nk_uniques_push without nk_uniques_begin");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    if (nseq > 0 && offsets[nseq] > 0 && !bases) return fail(NK_ERR_BAD_ARG, "null bases");
    return uniques_host_batch(h, bases, nullptr, nullptr, offsets, nseq);
}

int nk_uniques_push_packed(nk_counter* h, const uint32_t* codes, const uint32_t* other, const uint64_t* offsets,
                           uint64_t nseq) {
    if (is_group(h)) return nk_uniques_push_packed(h->group[0], codes, other, offsets, nseq);
    NK_TRY(validate_packed(h, codes, offsets, nseq));
    if (!h->uniques_open) return fail(NK_ERR_STATE, "nk_uniques_push_packed without nk_uniques_begin");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    return uniques_host_batch(h, nullptr, codes, other, offsets, nseq);
}

int nk_uniques_end(nk_counter* h) {
    if (is_group(h)) return nk_uniques_end(h->group[0]);
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (!h->uniques_open) return fail(NK_ERR_STATE, "nk_uniques_end without nk_uniques_begin");
    NK_CUDA(cudaSetDevice(h->cfg.device));
    return uniques_end(h);
}

int nk_set_file_uniques(nk_counter* h, uint64_t top_n) {
    if (!h) return fail(NK_ERR_BAD_ARG, "null handle");
    if (top_n > 2048) return fail(NK_ERR_BAD_ARG, "nk_set_file_uniques: top_n must be at most 2048");
    h->file_uniques = top_n;
    return NK_OK;
}

// Host-only: run the C++ record reader over a file and digest what it yields, so that the reader
// (formats, compression sniffing, stop-at-first-malformed-record) can be checked without a device.
int nk_debug_fastx_digest(const char* path, uint64_t* nrecords, uint64_t* nbases, uint64_t* fnv1a) {
    if (!path) return fail(NK_ERR_BAD_ARG, "null path");
    nk::FastxReader rd;
    std::string err;
    if (rd.open(path, &err) != 0) return fail(NK_ERR_IO, "%s", err.c_str());
    uint64_t nrec = 0, nb = 0, hsh = 0xcbf29ce484222325ull;
    std::vector<uint8_t> buf(1u << 20), rec;
    while (rd.next_record()) {
        rec.clear();
        bool done = false;
        while (!done) {
            const size_t n = rd.read_seq(buf.data(), buf.size(), &done);
            rec.insert(rec.end(), buf.begin(), buf.begin() + n);
        }
        if (!rd.finish_record(rec.size())) break;  // malformed FASTQ record: iteration ends, record dropped
        if (fnv1a) {  // NULL: parse only (reader throughput measurements)
            for (uint8_t b : rec) { hsh ^= b; hsh *= 0x100000001b3ull; }
            hsh ^= 0xFFu; hsh *= 0x100000001b3ull;  // record separator
        }
        ++nrec;
        nb += rec.size();
    }
    if (nrecords) *nrecords = nrec;
    if (nbases) *nbases = nb;
    if (fnv1a) *fnv1a = hsh;
    return NK_OK;
}

// Host-only: the same digest through the PARALLEL ingest's window planner + parser (run serially here):
// the file is cut into windows of `window` bytes exactly as count_fasta_parallel does, every window is
// stripped by fasta_parse_window, and the records are re-assembled from the per-window record starts.
int nk_debug_fasta_windows_digest(const char* path, uint64_t window, uint64_t* nrecords, uint64_t* nbases, uint64_t* fnv1a) {
    if (!path || window < 64) return fail(NK_ERR_BAD_ARG, "null path or window < 64");
    const int fd = ::open(path, O_RDONLY);
    if (fd < 0) return fail(NK_ERR_IO, "cannot open %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size == 0) { ::close(fd); return fail(NK_ERR_IO, "%s: empty file", path); }
    const size_t size = (size_t)st.st_size;
    void* map = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (map == MAP_FAILED) return fail(NK_ERR_IO, "mmap failed");
    const uint8_t* file = (const uint8_t*)map;
    uint64_t nrec = 0, nb = 0, hsh = 0xcbf29ce484222325ull;
    bool open_rec = false;
    std::vector<uint8_t> data((size_t)window + nk::kFastaSlack + 64);
    std::vector<uint64_t> starts;
    size_t ws = 0;
    int state = nk::FA_LINE_START;
    while (ws < size) {
        int next_state = nk::FA_LINE_START;
        const nk::FastaWindowPlan w = nk::fasta_plan_window(file, size, ws, (size_t)window, state, &next_state);
        size_t fill = 0;
        nk::fasta_parse_window(file, w, data.data(), &fill, &starts);
        size_t at = 0;
        for (size_t r = 0; r <= starts.size(); ++r) {
            const size_t end = r < starts.size() ? (size_t)starts[r] : fill;
            if (fnv1a)
                for (size_t i = at; i < end; ++i) { hsh ^= data[i]; hsh *= 0x100000001b3ull; }
            nb += end - at;
            at = end;
            if (r < starts.size()) {  // a header: the open record (if any) ends, a new one begins
                if (open_rec) { hsh ^= 0xFFu; hsh *= 0x100000001b3ull; ++nrec; }
                open_rec = true;
            }
        }
        ws = w.we;
        state = next_state;
    }
    if (open_rec) { hsh ^= 0xFFu; hsh *= 0x100000001b3ull; ++nrec; }
    munmap(map, size);
    if (nrecords) *nrecords = nrec;
    if (nbases) *nbases = nb;
    if (fnv1a) *fnv1a = hsh;
    return NK_OK;
}

int nk_process_file(nk_counter* h, const char* path, int streaming) {
    if (!h || !path) return fail(NK_ERR_BAD_ARG, "null argument");
    if (h->streaming || h->group_streaming) return fail(NK_ERR_STATE, "nk_process_file inside nk_stream_begin/end");
    std::string err;
    int rc = nk::process_file(h, path, streaming != 0, &err, false);
    if (rc != NK_OK) return fail(rc, "%s", err.c_str());
    if (h->file_uniques > 0 && !h->exact) {
        // `uniques` of the top rows: read the file a second time, keeping only the windows of those neurons
        NK_TRY(nk_uniques_begin(h, h->file_uniques));
        rc = nk::process_file(h, path, streaming != 0, &err, true);
        if (rc != NK_OK) {
            (is_group(h) ? h->group[0] : h)->uniques_open = false;
            return fail(rc, "%s", err.c_str());
        }
        NK_TRY(nk_uniques_end(h));
    }
    return NK_OK;
}

}  // extern "C"
