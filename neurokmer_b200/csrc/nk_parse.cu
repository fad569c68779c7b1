// nk_parse.cu — FASTA / FASTQ record parsing ON THE DEVICE (sm_100a).
//
// The reference's entry point is a path: needletail splits the file into records on one producer thread
// (src/utils.rs:9-24, src/spiking_hash.rs:285-303).  Here the raw file bytes are copied to the GPU and
// parsed there at HBM speed — no host pass over the bytes at all:
//   FASTA  the bytes of sequence lines, with every '\n' and '\r' removed, are compacted into one base
//          array; a '>' at a line start opens a record (its line is dropped); offsets[r] = position of
//          record r's first base.  "Am I inside a header line" is a scan over two kinds of events
//          ('\n', line-start '>'): the last event wins.
//   FASTQ  a record is exactly 4 lines, so the line number (a prefix sum over '\n') modulo 4 tells what a
//          byte is; line 4r+1 is record r's sequence.  Every record is then validated exactly as the host
//          reader does ('@' / '+' first bytes, |quality| == |sequence|) and the iteration ends at the
//          first malformed one (src/utils.rs:17-20).
// Record rules = nk_fastx.cpp's FastxReader (the host twin, which the oracle-checked file tests pin).
// All prefix sums are hand-written three-phase scans (segment summaries -> one-block scan -> apply).
#include "nk_kernels.cuh"

namespace nk {

namespace {

constexpr int PT = 256;                  // threads per block
constexpr int STEP = PT * 16;            // bytes one block consumes per iteration (16 per thread)
constexpr int SEG_ITERS = 4;
constexpr int SEG = STEP * SEG_ITERS;    // bytes per segment = unit of the inter-block scans (16 KiB)

// 16 file bytes of this thread -> bit masks (bit j = byte j)
struct Bits16 {
    unsigned nl, cr, gt, valid;
};

__device__ __forceinline__ unsigned eq4(unsigned w, unsigned pattern) {
    return ((__vcmpeq4(w, pattern) & 0x08040201u) * 0x01010101u) >> 24;
}

__device__ __forceinline__ Bits16 load_bits(const unsigned char* __restrict__ file, unsigned long long pos,
                                            unsigned long long size) {
    Bits16 b{0u, 0u, 0u, 0u};
    if (pos >= size) return b;
    uint4 v;
    if (pos + 16 <= size) {
        v = *reinterpret_cast<const uint4*>(file + pos);  // the buffer is 16-byte aligned and padded
        b.valid = 0xFFFFu;
    } else {
        unsigned char tmp[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) tmp[j] = pos + j < size ? file[pos + j] : (unsigned char)'x';
        v = *reinterpret_cast<uint4*>(tmp);
        b.valid = (1u << (unsigned)(size - pos)) - 1u;
    }
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        b.nl |= eq4(w[q], 0x0A0A0A0Au) << (4 * q);
        b.cr |= eq4(w[q], 0x0D0D0D0Du) << (4 * q);
        b.gt |= eq4(w[q], 0x3E3E3E3Eu) << (4 * q);
    }
    b.nl &= b.valid; b.cr &= b.valid; b.gt &= b.valid;
    return b;
}

// block-wide exclusive scan of one unsigned per thread; *total = block sum
__device__ __forceinline__ unsigned block_exscan(unsigned v, unsigned* s_warp, unsigned* total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (unsigned)o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < PT / 32; ++w) {
        const unsigned x = s_warp[w];
        if ((unsigned)w < warp) before += x;
        all += x;
    }
    __syncthreads();
    *total = all;
    return before + incl - v;
}

// ---- FASTA ------------------------------------------------------------------------------------------
// Events of a thread's 16 bytes: N = '\n', S = '>' whose previous byte is '\n' (or the file start).
// state 1 = inside a header line.  A chunk with any event fixes the state behind it (1 iff its last event is
// an S); a chunk without events passes its incoming state on.
struct FaThread {
    unsigned nl, cr, S, valid;
    bool has_event, last_is_S;
};

__device__ __forceinline__ FaThread fa_classify(const unsigned char* __restrict__ file, unsigned long long pos,
                                                unsigned long long size) {
    const Bits16 b = load_bits(file, pos, size);
    FaThread t;
    t.nl = b.nl; t.cr = b.cr; t.valid = b.valid;
    const bool prev_nl = pos == 0 || (pos < size && file[pos - 1] == '\n');
    t.S = b.gt & ((b.nl << 1) | (prev_nl ? 1u : 0u)) & 0xFFFFu;
    const unsigned ev = t.nl | t.S;
    t.has_event = ev != 0u;
    t.last_is_S = ev != 0u && ((t.S >> (31 - __clz(ev))) & 1u);
    return t;
}

// bytes of the chunk that are inside a header line, given the state at its first byte.  Adding the S bits to
// "no newline here" ripples a carry from every header start up to (and including) the newline that ends it.
__device__ __forceinline__ unsigned fa_header_mask(const FaThread& t, unsigned state_in) {
    const unsigned A = ~t.nl & 0xFFFFu;
    unsigned S = t.S;
    if (state_in) S |= 1u;
    return (A ^ (A + S)) & 0xFFFFu;
}

// state at this thread's first byte relative to the state `carry` at the start of the iteration.  s_ev/s_hs:
// per-warp summaries in shared memory (written here).  Returns also, through *iter_out, the state behind the
// whole iteration.
__device__ __forceinline__ unsigned fa_state_in(const FaThread& t, unsigned carry, unsigned* s_ev, unsigned* s_hs,
                                                unsigned* iter_out) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const unsigned E = __ballot_sync(0xFFFFFFFFu, t.has_event);
    const unsigned H = __ballot_sync(0xFFFFFFFFu, t.last_is_S);
    if (lane == 0) { s_ev[warp] = E; s_hs[warp] = H; }
    __syncthreads();
    unsigned st = carry;
    bool found = false;
    const unsigned below = E & ((1u << lane) - 1u);
    if (below) { st = (H >> (31 - __clz(below))) & 1u; found = true; }
    if (!found) {
        for (int w = (int)warp - 1; w >= 0; --w) {
            const unsigned e = s_ev[w];
            if (e) { st = (s_hs[w] >> (31 - __clz(e))) & 1u; break; }
        }
    }
    unsigned out = carry;
    for (int w = PT / 32 - 1; w >= 0; --w) {
        const unsigned e = s_ev[w];
        if (e) { out = (s_hs[w] >> (31 - __clz(e))) & 1u; break; }
    }
    *iter_out = out;
    __syncthreads();
    return st;
}

struct FaSummary {           // per segment
    unsigned kept[2];        // kept bytes if the segment starts outside / inside a header line
    unsigned nrec;           // header starts
    unsigned event;          // bit 0: any event, bit 1: the last event is a header start
};

__global__ void __launch_bounds__(PT) fa_summary_kernel(const unsigned char* __restrict__ file, unsigned long long size,
                                                         unsigned long long seg_first, FaSummary* __restrict__ sum) {
    __shared__ unsigned s_ev[PT / 32], s_hs[PT / 32], s_red[3][PT / 32];
    const unsigned long long seg0 = (seg_first + blockIdx.x) * SEG;
    unsigned carry0 = 0, carry1 = 1;   // the two assumptions about the state at the segment's first byte
    unsigned k0 = 0, k1 = 0, nrec = 0;
    for (int it = 0; it < SEG_ITERS; ++it) {
        const unsigned long long pos = seg0 + (unsigned long long)it * STEP + threadIdx.x * 16ull;
        const FaThread t = fa_classify(file, pos, size);
        unsigned out0, out1;
        const unsigned st0 = fa_state_in(t, carry0, s_ev, s_hs, &out0);
        // (the second assumption differs only while no event has been seen: recompute from the same ballots)
        unsigned st1 = st0, o1 = out0;
        if (carry1 != carry0) {
            st1 = fa_state_in(t, carry1, s_ev, s_hs, &o1);
        }
        out1 = o1;
        const unsigned body = ~t.nl & ~t.cr & t.valid;
        k0 += __popc(body & ~fa_header_mask(t, st0));
        k1 += __popc(body & ~fa_header_mask(t, st1));
        nrec += __popc(t.S);
        carry0 = out0; carry1 = out1;
    }
    k0 = __reduce_add_sync(0xFFFFFFFFu, k0);
    k1 = __reduce_add_sync(0xFFFFFFFFu, k1);
    nrec = __reduce_add_sync(0xFFFFFFFFu, nrec);
    if ((threadIdx.x & 31) == 0) { s_red[0][threadIdx.x >> 5] = k0; s_red[1][threadIdx.x >> 5] = k1; s_red[2][threadIdx.x >> 5] = nrec; }
    __syncthreads();
    if (threadIdx.x == 0) {
        FaSummary s{{0u, 0u}, 0u, 0u};
        for (int w = 0; w < PT / 32; ++w) { s.kept[0] += s_red[0][w]; s.kept[1] += s_red[1][w]; s.nrec += s_red[2][w]; }
        // any event in the segment <=> both assumptions ended in the same state ... or the segment is event-free and
        // they still differ: then it passes its incoming state on
        s.event = carry0 == carry1 ? (1u | (carry0 << 1)) : 0u;
        sum[seg_first + blockIdx.x] = s;
    }
}

struct SegPrefix {               // per segment, after the one-block scan
    unsigned long long out;      // kept bytes before the segment
    unsigned long long rec;      // records opened before the segment (FASTQ: '\n' before the segment)
    unsigned state;              // FASTA: inside a header line at the segment's first byte
    unsigned _pad;
};

// one block: sequential meaning, parallel execution — thread t owns a contiguous run of segments
// Scans the segments [seg_first, seg_first + nseg).  carry (device, 3 u64: state, bases, records) is what the segments
// before them left behind — {0, 0, 0} for a whole file — and is updated for the next range (files parsed chunk by chunk).
__global__ void __launch_bounds__(1024) fa_scan_kernel(const FaSummary* __restrict__ sum, unsigned long long seg_first,
                                                        unsigned long long nseg, SegPrefix* __restrict__ pre,
                                                        unsigned long long* __restrict__ totals, unsigned long long* __restrict__ carry) {
    __shared__ unsigned s_event[1024];
    __shared__ unsigned long long s_out[1024], s_rec[1024];
    const unsigned t = threadIdx.x;
    sum += seg_first;
    pre += seg_first;
    const unsigned long long per = (nseg + 1023) / 1024;
    const unsigned long long a = (unsigned long long)t * per < nseg ? (unsigned long long)t * per : nseg;
    const unsigned long long b = a + per < nseg ? a + per : nseg;
    const unsigned carry_state = (unsigned)carry[0];
    const unsigned long long carry_out = carry[1], carry_rec = carry[2];
    unsigned ev = 0;
    for (unsigned long long s = a; s < b; ++s) {
        const unsigned e = sum[s].event;
        if (e & 1u) ev = e;
    }
    s_event[t] = ev;
    __syncthreads();
    unsigned state = carry_state;  // a file starts at a line start, outside a header (its first '>' is an event)
    for (int q = (int)t - 1; q >= 0; --q)
        if (s_event[q] & 1u) { state = s_event[q] >> 1; break; }
    unsigned long long out = 0, rec = 0;
    unsigned st = state;
    for (unsigned long long s = a; s < b; ++s) {
        const FaSummary x = sum[s];
        out += x.kept[st];
        rec += x.nrec;
        if (x.event & 1u) st = x.event >> 1;
    }
    s_out[t] = out; s_rec[t] = rec;
    __syncthreads();
    __shared__ unsigned s_last_state;
    if (t == 1023) s_last_state = st;  // state behind the last segment of the range (empty thread ranges pass it on)
    __syncthreads();
    if (t == 0) {
        unsigned long long ro = carry_out, rr = carry_rec;
        for (int q = 0; q < 1024; ++q) {
            const unsigned long long o = s_out[q], r = s_rec[q];
            s_out[q] = ro; s_rec[q] = rr;
            ro += o; rr += r;
        }
        totals[0] = ro;  // bases so far
        totals[1] = rr;  // records so far
        carry[0] = s_last_state; carry[1] = ro; carry[2] = rr;
    }
    __syncthreads();
    out = s_out[t]; rec = s_rec[t]; st = state;
    for (unsigned long long s = a; s < b; ++s) {
        const FaSummary x = sum[s];
        pre[s] = SegPrefix{out, rec, st, 0u};
        out += x.kept[st];
        rec += x.nrec;
        if (x.event & 1u) st = x.event >> 1;
    }
}

// compacted bytes of one iteration: staged in shared memory at the output's own 16-byte phase, then copied out
// with aligned 16-byte stores (head / tail bytes singly: neighbouring blocks own the other bytes of those chunks)
__device__ __forceinline__ void flush_bytes(unsigned char* __restrict__ dst, unsigned long long out_pos, unsigned n,
                                            const unsigned char* s_buf /* index (out_pos & 15) + j */) {
    const unsigned shift = (unsigned)(out_pos & 15ull);
    unsigned char* base = dst + (out_pos - shift);  // 16-byte aligned; smem index i <-> base[i]
    const unsigned lo = shift, hi = shift + n;
    const unsigned alo = (lo + 15u) & ~15u, ahi = hi & ~15u;
    if (alo < ahi) {
        for (unsigned i = lo + threadIdx.x; i < alo; i += PT) base[i] = s_buf[i];
        for (unsigned i = alo / 16 + threadIdx.x; i < ahi / 16; i += PT)
            reinterpret_cast<uint4*>(base)[i] = reinterpret_cast<const uint4*>(s_buf)[i];
        for (unsigned i = ahi + threadIdx.x; i < hi; i += PT) base[i] = s_buf[i];
    } else {
        for (unsigned i = lo + threadIdx.x; i < hi; i += PT) base[i] = s_buf[i];
    }
}

__global__ void __launch_bounds__(PT) fa_write_kernel(const unsigned char* __restrict__ file, unsigned long long size,
                                                       unsigned long long seg_first, const SegPrefix* __restrict__ pre,
                                                       unsigned char* __restrict__ bases, unsigned long long* __restrict__ offsets,
                                                       unsigned long long offsets_cap) {
    __shared__ unsigned s_ev[PT / 32], s_hs[PT / 32], s_warp[PT / 32];
    __shared__ __align__(16) unsigned char s_buf[STEP + 32];
    const unsigned long long seg0 = (seg_first + blockIdx.x) * SEG;
    const SegPrefix p = pre[seg_first + blockIdx.x];
    unsigned long long out_pos = p.out, rec = p.rec;
    unsigned carry = p.state;
    for (int it = 0; it < SEG_ITERS; ++it) {
        const unsigned long long pos = seg0 + (unsigned long long)it * STEP + threadIdx.x * 16ull;
        if (seg0 + (unsigned long long)it * STEP >= size) break;  // block-uniform
        const FaThread t = fa_classify(file, pos, size);
        unsigned out_state;
        const unsigned st = fa_state_in(t, carry, s_ev, s_hs, &out_state);
        const unsigned keep = ~t.nl & ~t.cr & t.valid & ~fa_header_mask(t, st);
        const unsigned cnt = __popc(keep);
        unsigned total = 0, total_s = 0;
        const unsigned mine = block_exscan(cnt, s_warp, &total);
        const unsigned mine_s = block_exscan(__popc(t.S), s_warp, &total_s);
        const unsigned shift = (unsigned)(out_pos & 15ull);
        // this thread's kept bytes, in order
        if (cnt) {
            const uint4 v = pos + 16 <= size ? *reinterpret_cast<const uint4*>(file + pos) : make_uint4(0, 0, 0, 0);
            const unsigned w[4] = {v.x, v.y, v.z, v.w};
            unsigned o = shift + mine;
            if (pos + 16 <= size) {
                for (unsigned m = keep; m; m &= m - 1) {
                    const unsigned j = __ffs(m) - 1;
                    s_buf[o++] = (unsigned char)(w[j >> 2] >> (8 * (j & 3)));
                }
            } else {
                for (unsigned m = keep; m; m &= m - 1) s_buf[o++] = file[pos + __ffs(m) - 1];
            }
        }
        // record starts: the first base of record r is the next kept byte
        unsigned rs = 0;
        for (unsigned m = t.S; m; m &= m - 1, ++rs) {
            const unsigned j = __ffs(m) - 1;
            if (rec + mine_s + rs < offsets_cap) offsets[rec + mine_s + rs] = out_pos + mine + __popc(keep & ((1u << j) - 1u));
        }
        __syncthreads();
        flush_bytes(bases, out_pos, total, s_buf);
        __syncthreads();
        out_pos += total;
        rec += total_s;
        carry = out_state;
    }
}

// ---- FASTQ ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PT) fq_newlines_kernel(const unsigned char* __restrict__ file, unsigned long long size,
                                                          unsigned* __restrict__ nl_count) {
    __shared__ unsigned s_red[PT / 32];
    const unsigned long long seg0 = (unsigned long long)blockIdx.x * SEG;
    unsigned n = 0;
    for (int it = 0; it < SEG_ITERS; ++it) {
        const unsigned long long pos = seg0 + (unsigned long long)it * STEP + threadIdx.x * 16ull;
        n += __popc(load_bits(file, pos, size).nl);
    }
    n = __reduce_add_sync(0xFFFFFFFFu, n);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = n;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < PT / 32; ++w) t += s_red[w];
        nl_count[blockIdx.x] = t;
    }
}

// one block: exclusive scan of a u32 array into the `rec` (which == 0) or `out` (which == 1) field of pre[]
__global__ void __launch_bounds__(1024) scan_u32_kernel(const unsigned* __restrict__ in, unsigned long long n,
                                                         SegPrefix* __restrict__ pre, int which,
                                                         unsigned long long* __restrict__ total) {
    __shared__ unsigned long long s_sum[1024];
    const unsigned t = threadIdx.x;
    const unsigned long long per = (n + 1023) / 1024;
    const unsigned long long a = (unsigned long long)t * per, b = a + per < n ? a + per : n;
    unsigned long long s = 0;
    for (unsigned long long i = a; i < b; ++i) s += in[i];
    s_sum[t] = s;
    __syncthreads();
    if (t == 0) {
        unsigned long long run = 0;
        for (int q = 0; q < 1024; ++q) { const unsigned long long x = s_sum[q]; s_sum[q] = run; run += x; }
        *total = run;
    }
    __syncthreads();
    s = s_sum[t];
    for (unsigned long long i = a; i < b; ++i) {
        if (which == 0) pre[i].rec = s; else pre[i].out = s;
        s += in[i];
    }
}

// positions of the chunk that lie on a sequence line (line number = 1 mod 4), given the number of the line the
// chunk's first byte is on
__device__ __forceinline__ unsigned fq_seq_line_mask(unsigned nl, unsigned long long line_in) {
    unsigned m = 0, rest = 0xFFFFu;
    unsigned long long line = line_in;
    for (;;) {
        const unsigned upto = nl ? (nl & (0u - nl)) : 0x10000u;  // the next newline (it belongs to its own line)
        if ((line & 3ull) == 1ull) m |= (upto - 1u) & rest;
        if (!nl) break;
        rest &= ~((upto << 1) - 1u);
        nl &= nl - 1u;
        ++line;
    }
    return m;
}

// MODE 0: count the kept bytes per segment.  MODE 1: write them, the position of every newline, and
// offsets[r+1] = bases before the end of record r's sequence line.
template <int MODE>
__global__ void __launch_bounds__(PT) fq_pass_kernel(const unsigned char* __restrict__ file, unsigned long long size,
                                                      const SegPrefix* __restrict__ pre, unsigned* __restrict__ kept_count,
                                                      unsigned char* __restrict__ bases, unsigned long long* __restrict__ offsets,
                                                      unsigned long long* __restrict__ line_end) {
    __shared__ unsigned s_warp[PT / 32];
    __shared__ __align__(16) unsigned char s_buf[MODE == 1 ? STEP + 32 : 16];
    const unsigned long long seg0 = (unsigned long long)blockIdx.x * SEG;
    const SegPrefix p = pre[blockIdx.x];
    unsigned long long line0 = p.rec, out_pos = MODE == 1 ? p.out : 0ull;
    unsigned seg_kept = 0;
    for (int it = 0; it < SEG_ITERS; ++it) {
        const unsigned long long pos = seg0 + (unsigned long long)it * STEP + threadIdx.x * 16ull;
        if (seg0 + (unsigned long long)it * STEP >= size) break;  // block-uniform
        const Bits16 b = load_bits(file, pos, size);
        unsigned total_nl = 0;
        const unsigned nl_before = block_exscan(__popc(b.nl), s_warp, &total_nl);
        const unsigned long long line_in = line0 + nl_before;
        const unsigned keep = fq_seq_line_mask(b.nl, line_in) & ~b.nl & ~b.cr & b.valid;
        const unsigned cnt = __popc(keep);
        if (MODE == 0) {
            seg_kept += cnt;
        } else {
            unsigned total = 0;
            const unsigned mine = block_exscan(cnt, s_warp, &total);
            const unsigned shift = (unsigned)(out_pos & 15ull);
            if (cnt) {
                unsigned o = shift + mine;
                for (unsigned m = keep; m; m &= m - 1) s_buf[o++] = file[pos + __ffs(m) - 1];
            }
            unsigned long long line = line_in;
            for (unsigned m = b.nl; m; m &= m - 1, ++line) {
                const unsigned j = __ffs(m) - 1;
                line_end[line] = pos + j;
                if ((line & 3ull) == 1ull) offsets[(line >> 2) + 1] = out_pos + mine + __popc(keep & ((1u << j) - 1u));
            }
            __syncthreads();
            flush_bytes(bases, out_pos, total, s_buf);
            __syncthreads();
            out_pos += total;
        }
        line0 += total_nl;
    }
    if (MODE == 0) {
        seg_kept = __reduce_add_sync(0xFFFFFFFFu, seg_kept);
        if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = seg_kept;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned t = 0;
            for (int w = 0; w < PT / 32; ++w) t += s_warp[w];
            kept_count[blockIdx.x] = t;
        }
    }
}

// one thread per complete record: the host reader's checks (nk_fastx.cpp: next_record / finish_record)
__global__ void fq_validate_kernel(const unsigned char* __restrict__ file, unsigned long long size_real,
                                   const unsigned long long* __restrict__ line_end, const unsigned long long* __restrict__ offsets,
                                   unsigned long long nrec, unsigned long long* __restrict__ first_bad) {
    const unsigned long long r = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (r >= nrec) return;
    const unsigned long long hdr = r ? line_end[4 * r - 1] + 1 : 0ull;
    const unsigned long long sep = line_end[4 * r + 1] + 1;
    const unsigned long long q0 = line_end[4 * r + 2] + 1, q1 = line_end[4 * r + 3];
    bool ok = file[hdr] == '@' && file[sep] == '+';
    unsigned long long qlen = q1 - q0;
    // a '\r' before the terminating '\n' is not part of the quality string (a line ended by EOF keeps it)
    if (qlen > 0 && q1 < size_real && file[q1 - 1] == '\r') --qlen;
    ok = ok && qlen == offsets[r + 1] - offsets[r];
    if (!ok) atomicMin(first_bad, r);
}

__global__ void fq_finish_kernel(const unsigned long long* __restrict__ offsets, unsigned long long nrec_complete,
                                 unsigned long long* __restrict__ totals /* [0] bases [1] records */,
                                 const unsigned long long* __restrict__ first_bad) {
    const unsigned long long n = *first_bad < nrec_complete ? *first_bad : nrec_complete;
    totals[1] = n;
    totals[0] = offsets[n];
}

unsigned long long nseg_of(unsigned long long size) { return (size + SEG - 1) / SEG; }

}  // namespace

size_t parse_scratch_bytes(unsigned long long size) {
    const unsigned long long nseg = nseg_of(size);
    return (size_t)(nseg * (sizeof(FaSummary) + sizeof(SegPrefix) + 2 * sizeof(unsigned)) + 256);
}

// segments [seg_first, seg_first + nseg) of the raw bytes (a whole file: 0, all); carry: see fa_scan_kernel.
// The range must end at the end of the file or at a multiple of parse_segment_bytes().
cudaError_t launch_fasta_plan(const unsigned char* file, unsigned long long size, unsigned long long seg_first,
                              unsigned long long nseg, void* scratch, unsigned long long* totals, unsigned long long* carry,
                              cudaStream_t s) {
    const unsigned long long nseg_all = nseg_of(size);
    if (nseg == 0 || nseg_all > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    FaSummary* sum = reinterpret_cast<FaSummary*>(scratch);
    SegPrefix* pre = reinterpret_cast<SegPrefix*>(sum + nseg_all);
    fa_summary_kernel<<<(unsigned)nseg, PT, 0, s>>>(file, size, seg_first, sum);
    fa_scan_kernel<<<1, 1024, 0, s>>>(sum, seg_first, nseg, pre, totals, carry);
    return cudaGetLastError();
}

cudaError_t launch_fasta_write(const unsigned char* file, unsigned long long size, unsigned long long seg_first,
                               unsigned long long nseg, void* scratch, unsigned char* bases, unsigned long long* offsets,
                               unsigned long long offsets_cap, cudaStream_t s) {
    const unsigned long long nseg_all = nseg_of(size);
    FaSummary* sum = reinterpret_cast<FaSummary*>(scratch);
    SegPrefix* pre = reinterpret_cast<SegPrefix*>(sum + nseg_all);
    fa_write_kernel<<<(unsigned)nseg, PT, 0, s>>>(file, size, seg_first, pre, bases, offsets, offsets_cap);
    return cudaGetLastError();
}

unsigned long long parse_segment_bytes() { return SEG; }

cudaError_t launch_fastq_lines(const unsigned char* file, unsigned long long size, void* scratch,
                               unsigned long long* totals /* [2] <- number of '\n' */, cudaStream_t s) {
    const unsigned long long nseg = nseg_of(size);
    if (nseg == 0 || nseg > 0x7FFFFFFFull) return cudaErrorInvalidValue;
    FaSummary* sum = reinterpret_cast<FaSummary*>(scratch);
    SegPrefix* pre = reinterpret_cast<SegPrefix*>(sum + nseg);
    unsigned* cnt = reinterpret_cast<unsigned*>(pre + nseg);
    fq_newlines_kernel<<<(unsigned)nseg, PT, 0, s>>>(file, size, cnt);
    scan_u32_kernel<<<1, 1024, 0, s>>>(cnt, nseg, pre, 0, totals + 2);
    return cudaGetLastError();
}

cudaError_t launch_fastq_write(const unsigned char* file, unsigned long long size, unsigned long long size_real, void* scratch,
                               unsigned char* bases, unsigned long long* offsets, unsigned long long* line_end,
                               unsigned long long nlines, unsigned long long* totals, cudaStream_t s) {
    const unsigned long long nseg = nseg_of(size);
    FaSummary* sum = reinterpret_cast<FaSummary*>(scratch);
    SegPrefix* pre = reinterpret_cast<SegPrefix*>(sum + nseg);
    unsigned* cnt = reinterpret_cast<unsigned*>(pre + nseg);
    unsigned* kept = cnt + nseg;
    fq_pass_kernel<0><<<(unsigned)nseg, PT, 0, s>>>(file, size, pre, kept, nullptr, nullptr, nullptr);
    scan_u32_kernel<<<1, 1024, 0, s>>>(kept, nseg, pre, 1, totals + 3);
    cudaError_t e = cudaMemsetAsync(offsets, 0, sizeof(unsigned long long), s);  // offsets[0] = 0
    if (e != cudaSuccess) return e;
    fq_pass_kernel<1><<<(unsigned)nseg, PT, 0, s>>>(file, size, pre, nullptr, bases, offsets, line_end);
    const unsigned long long nrec = nlines / 4;
    e = cudaMemsetAsync(totals + 4, 0xFF, sizeof(unsigned long long), s);         // first_bad = u64 max
    if (e != cudaSuccess) return e;
    if (nrec) fq_validate_kernel<<<(unsigned)((nrec + 255) / 256), 256, 0, s>>>(file, size_real, line_end, offsets, nrec, totals + 4);
    fq_finish_kernel<<<1, 1, 0, s>>>(offsets, nrec, totals, totals + 4);
    return cudaGetLastError();
}

}  // namespace nk
