// nk_count.cu — fused windowing → SipHash-1-3 → % pool → pool update (sm_100a).
//
// Replaces the reference's per-base hot loop:
//   RollingKmerHash::init/slide + canonical   src/models.rs:206-286
//   pack_kmer (non-canonical)                 src/utils.rs:26-39
//   map_kmer_to_neuron                        src/spiking_hash.rs:78-82
//   currents[idx] += 1                        src/spiking_hash.rs:112,126,136,332,342,349
//
// Data flow per CTA (persistent, dynamic tile scheduler):
//   HBM --cp.async.bulk (TMA 1-D, mbarrier)--> smem stage {tile bytes + halo, invalid-start bits}
//   warp: 512-position chunks; lane L converts its 16 ASCII bytes into a forward
//   word F (MSB-first 2-bit codes) and a complement word R (LSB-first), pulls the
//   words of lanes L+1, L+2 by shuffle (k-1 overlap; the last two lanes read the
//   next chunk's first words), and then every one of its 16 window starts is two
//   funnel-shift extractions: fwd(j) = bits of F-stream, rc(j) = bits of R-stream.
//   A k-mer is a pure function of its own k bytes (the reference has no N-break
//   state: non-ACGT -> code 0 on BOTH strands), so there is no serial dependency.
//   min(fwd, rc) -> SipHash-1-3 -> exact mod -> RED.ADD.U32 into the L2-resident pool.
#include <atomic>
#include <mutex>

#include "nk_kernels.cuh"

// windows hashed side by side in the inner loop.  Round 1 (other kernel body): 2 / 4 / 8 / 16 -> 0.967 / 0.971 / 0.972 /
// 1.027 ms; the round-2 kernel: 1 / 2 / 3 / 4 / 8 -> 0.7624 / 0.7606 / 0.7606 / 0.7690 / 0.7813 ms (profiles/r02_variants.md)
#ifndef NK_COUNT_UNROLL
#define NK_COUNT_UNROLL 2
#endif

namespace nk {

namespace {

constexpr int kBytesPerStage = COUNT_TILE + COUNT_HALO;    // multiple of 16
constexpr int kBitsPerStage = COUNT_TILE / 8;              // multiple of 16
constexpr int kStageStride = ((kBytesPerStage + kBitsPerStage + 127) / 128) * 128;
constexpr int kSmemTotal = COUNT_STAGES * kStageStride;
constexpr int kCompactBytes = COUNT_WARPS * COUNT_CHUNK * 8;  // MODE 3: one 512-word buffer per warp

// pre-packed input: 2-bit codes (+ 64 bases of halo) and `other` bits (+ 128 bases of halo) per stage
constexpr int kPkCodes = COUNT_TILE / 4 + 16;
constexpr int kPkOther = COUNT_TILE / 8 + 16;
static_assert(kPkCodes + kPkOther + kBitsPerStage <= kStageStride, "a packed stage fits an ASCII stage");

static_assert(kBytesPerStage % 16 == 0 && kBitsPerStage % 16 == 0, "TMA bulk copies are 16-byte granular");
static_assert(COUNT_HALO == 32, "the last warp converts exactly two 16-byte halo words");

struct WindowConsts {
    unsigned wide;     // k <= 16: the forward window lives in (F0:F1) only
    unsigned sh;       // (64 - 2k) & 31
    unsigned mask_lo;  // low / high 32 bits of 2^(2k)-1
    unsigned mask_hi;
    unsigned k;
};

__device__ __forceinline__ WindowConsts make_window_consts(unsigned k) {
    WindowConsts c;
    const unsigned s0 = 64u - 2u * k;
    c.wide = s0 >= 32u;
    c.sh = s0 & 31u;
    c.mask_lo = k >= 16u ? 0xFFFFFFFFu : ((1u << (2u * k)) - 1u);
    c.mask_hi = k <= 16u ? 0u : (k == 32u ? 0xFFFFFFFFu : ((1u << (2u * k - 32u)) - 1u));
    c.k = k;
    return c;
}

template <bool PACKED, bool BITMAP>
__device__ __forceinline__ void issue_tile(unsigned char* stage, unsigned long long* bar,
                                           const CountParams& p, unsigned long long tile) {
    // bytes of a tile copy that starts at `off`, clamped to the caller's array when it has a known end
    auto clamp = [](unsigned full, unsigned long long off, unsigned long long total) -> unsigned {
        return (total != 0 && off + full > total) ? (unsigned)(total - off) : full;
    };
    if constexpr (PACKED) {
        const bool has_other = p.other != nullptr;
        const unsigned long long coff = tile * (COUNT_TILE / 4), ooff = tile * (COUNT_TILE / 8);
        const unsigned cb = clamp(kPkCodes, coff, p.bases_bytes);
        const unsigned ob = has_other ? clamp(kPkOther, ooff, p.other_bytes) : 0u;
        mbar_expect_tx(bar, cb + (BITMAP ? kBitsPerStage : 0) + ob);
        tma_load_1d(stage, p.bases + coff, cb, bar);
        if (has_other) tma_load_1d(stage + kPkCodes, p.other + ooff, ob, bar);
        if (BITMAP)
            tma_load_1d(stage + kPkCodes + kPkOther, (const unsigned char*)p.invalid + tile * kBitsPerStage,
                        kBitsPerStage, bar);
    } else {
        const unsigned long long boff = tile * COUNT_TILE;
        const unsigned bb = clamp(kBytesPerStage, boff, p.bases_bytes);
        mbar_expect_tx(bar, bb + (BITMAP ? kBitsPerStage : 0));
        tma_load_1d(stage, p.bases + boff, bb, bar);
        if (BITMAP)
            tma_load_1d(stage + kBytesPerStage, (const unsigned char*)p.invalid + tile * kBitsPerStage,
                        kBitsPerStage, bar);
    }
}

// code words of the 16 positions [16*i, 16*i + 16) of a staged tile
template <bool CANON, bool PACKED>
__device__ __forceinline__ Codes16 load_codes(const unsigned char* sb, unsigned i, bool has_other) {
    if constexpr (PACKED) {
        const unsigned f = reinterpret_cast<const unsigned*>(sb)[i];
        const unsigned x = has_other ? reinterpret_cast<const unsigned short*>(sb + kPkCodes)[i] : 0u;
        return convert16_packed<!CANON>(f, x);
    } else {
        return convert16<!CANON>(reinterpret_cast<const uint4*>(sb)[i]);
    }
}

// ---- MODE 5: invalid window starts without a bitmap ---------------------------------------------------
// Start x is invalid iff the first sequence end e > x satisfies x > e - k (the window would cross e), or x lies
// beyond the chunk.  Per tile, one thread looks up the (at most kTileEnds) sequence ends that can reach into the
// tile; lanes turn them into their 16 bits.  A tile with more ends than that (a run of tiny sequences inside a
// long-sequence batch) makes every lane walk the offsets itself.
constexpr int kTileEnds = 4;
struct TileEnds {
    int n;                          // ends stored; kTileEnds + 1: dense tile, lanes walk `offsets` from `first`
    int limit;                      // tile-relative end of the chunk's window starts (COUNT_TILE, less in the last tile)
    int rel[kTileEnds];             // tile-relative positions of the ends (> 0, < COUNT_TILE + 32)
    unsigned long long first;       // index (into offsets[1..]) of the first end > tile start
};

__device__ __forceinline__ void find_tile_ends(const CountParams& p, unsigned long long tile, TileEnds* te,
                                               unsigned long long* cursor = nullptr) {
    const unsigned long long t0 = p.origin + tile * COUNT_TILE;
    // first end > t0 among offsets[1 .. nseq]; index i means offsets[i + 1], the answer lies in [lo, hi].
    // A CTA's tiles only move forward, so the search gallops from the previous tile's answer (`cursor`): in a
    // long-sequence batch that is ONE load per tile instead of log2(nseq) dependent ones.
    unsigned long long lo = cursor ? *cursor : 0ull, hi = p.nseq;
    if (cursor) {
        unsigned long long step = 1, probe = lo;
        while (probe < p.nseq) {
            if (__ldg(p.offsets + probe + 1) > t0) { hi = probe; break; }
            lo = probe + 1;
            probe += step;
            step <<= 1;
        }
    }
    while (lo < hi) {
        const unsigned long long mid = (lo + hi) >> 1;
        if (__ldg(p.offsets + mid + 1) > t0) hi = mid; else lo = mid + 1;
    }
    if (cursor) *cursor = lo;
    te->first = lo;
    const unsigned long long reach = t0 + COUNT_TILE + (p.k - 1);  // an end e invalidates starts e-k+1 .. e-1
    int n = 0;
    for (unsigned long long i = lo; i < p.nseq; ++i) {
        const unsigned long long e = __ldg(p.offsets + i + 1);
        if (e >= reach) break;
        if (n == kTileEnds) { n = kTileEnds + 1; break; }
        te->rel[n++] = (int)(e - t0);
    }
    te->n = n;
    const unsigned long long left = p.nstarts - tile * COUNT_TILE;
    te->limit = left < (unsigned long long)COUNT_TILE ? (int)left : COUNT_TILE;
}

// bits j (0..15): start q + j (tile-relative) is invalid
__device__ __forceinline__ unsigned range_bits(int lo, int hi) {  // positions [lo, hi) clipped to [0, 16)
    lo = lo < 0 ? 0 : lo;
    hi = hi > 16 ? 16 : hi;
    return hi > lo ? ((1u << hi) - 1u) & ~((1u << lo) - 1u) : 0u;
}
__device__ __forceinline__ unsigned invalid_bits(const CountParams& p, const TileEnds& te, unsigned long long tile, int q) {
    unsigned inv = range_bits(te.limit - q, 16);
    const int km1 = (int)p.k - 1;
    if (te.n <= kTileEnds) {
        for (int i = 0; i < te.n; ++i) inv |= range_bits(te.rel[i] - km1 - q, te.rel[i] - q);
    } else {
        const unsigned long long t0 = p.origin + tile * COUNT_TILE;
        for (unsigned long long i = te.first; i < p.nseq; ++i) {
            const long long e = (long long)(__ldg(p.offsets + i + 1) - t0);  // tile-relative, > 0
            if (e - km1 >= (long long)q + 16) break;
            inv |= range_bits((int)(e - km1) - q, e - q > 16 ? 16 : (int)e - q);
        }
    }
    return inv;
}

// word of lane (lane + d) of the 64-word sequence {cur[0..31], nxt[0..31]}
__device__ __forceinline__ unsigned neighbour(unsigned cur, unsigned nxt, unsigned lane, unsigned d) {
    return __shfl_sync(0xFFFFFFFFu, lane >= d ? cur : nxt, (lane + d) & 31u);
}

// pack_kmer over a window that contains non-ACGT bytes: they are skipped (src/utils.rs:35)
__device__ __noinline__ unsigned long long pack_skip(unsigned F0, unsigned F1, unsigned F2, unsigned V0,
                                                     unsigned V1, unsigned V2, unsigned j, unsigned k) {
    unsigned long long word = 0;
    for (unsigned i = 0; i < k; ++i) {
        const unsigned b = j + i, w = b >> 4, sh = 30u - 2u * (b & 15u);
        const unsigned f = w == 0 ? F0 : (w == 1 ? F1 : F2);
        const unsigned v = w == 0 ? V0 : (w == 1 ? V1 : V2);
        if ((v >> sh) & 1u) word = (word << 2) | ((f >> sh) & 3u);
    }
    return word;
}

template <bool CANON, int MODE, int MODK, bool KHI>
__device__ __forceinline__ void process_chunk(const CountParams& p, const WindowConsts& wc, const Codes16& cur,
                                              const Codes16& nxt, unsigned inv16, unsigned lane,
                                              unsigned long long pos0, unsigned long long* wbuf) {
    // (inv16 is a by-value copy: the N-run shortcut below edits it)
    const unsigned F0 = cur.F;
    const unsigned F1 = neighbour(cur.F, nxt.F, lane, 1);
    const unsigned F2 = neighbour(cur.F, nxt.F, lane, 2);
    // forward stream B = F0:F1:F2 (base b in bits [94-2b, 96-2b)); G = B >> (64-2k), so that
    // window j is (G >> (32-2j)) & mask.
    const unsigned t0 = wc.wide ? 0u : F0, t1 = wc.wide ? F0 : F1, t2 = wc.wide ? F1 : F2;
    const unsigned G0 = t0 >> wc.sh;
    const unsigned G1 = __funnelshift_r(t1, t0, wc.sh);
    const unsigned G2 = __funnelshift_r(t2, t1, wc.sh);

    unsigned R0 = 0, R1 = 0, R2 = 0, H0 = 0, H1 = 0, H2 = 0, V0 = 0, V1 = 0, V2 = 0;
    if (CANON) {
        // complement stream L = R2:R1:R0 (base b in bits [2b, 2b+2)); window j is (L >> 2j) & mask
        R0 = cur.R;
        R1 = neighbour(cur.R, nxt.R, lane, 1);
        R2 = neighbour(cur.R, nxt.R, lane, 2);
    } else {
        V0 = cur.V;
        V1 = neighbour(cur.V, nxt.V, lane, 1);
        V2 = neighbour(cur.V, nxt.V, lane, 2);
        const unsigned u0 = wc.wide ? 0u : V0, u1 = wc.wide ? V0 : V1, u2 = wc.wide ? V1 : V2;
        H0 = u0 >> wc.sh;
        H1 = __funnelshift_r(u1, u0, wc.sh);
        H2 = __funnelshift_r(u2, u1, wc.sh);
    }

    constexpr bool EMIT = MODE == 1;
#ifndef NK_NO_NRUN
    if constexpr (CANON && (MODE == 0 || MODE == 2 || MODE == 4)) {
        // Runs of N (assembly gaps, centromeres: megabases of them in real genomes), poly-A, poly-T.  A base that is not
        // ACGT is code 0 on BOTH strands (src/models.rs:237,249), A is 0 on the forward strand, T on the complement:
        // a lane whose 48 bases are all code 0 on ONE strand has 16 windows whose word on that strand is 0, and 0 is the
        // canonical minimum whatever the other strand holds.  All of them land on ONE neuron: half a percent of N in the
        // input sent 565,000 of the bench job's reductions to a single address, and the L2 slice that owns it made the
        // whole kernel 11 % slower (tools/count_ablate.py).  The warp adds such lanes' windows up and sends one reduction.
        // Exact tables (mode 2): those windows are not appended one by one either — their number goes to
        // words_cursor[3] and exact_finalize adds ONE record {word 0, that count} (565,000 copies of one word in one
        // bucket serialised its scatter cursor and its shared-memory counter).  Uniques pass (mode 4): one copy.
        const bool all_other = (F0 | F1 | F2) == 0u || (R0 | R1 | R2) == 0u;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, all_other);
        if (m) {
            const unsigned cnt = __reduce_add_sync(0xFFFFFFFFu, all_other ? 16u - __popc(inv16 & 0xFFFFu) : 0u);
            if (lane == 0 && cnt) {
                const unsigned idx0 = fastmod_kind_dev<MODK>(siphash13_dev(0u, 0u, p.rm), p.fm);
                if (MODE != 4) atomicAdd(p.acc + idx0, cnt);
                if (MODE == 2) atomicAdd(p.words_cursor + 3, (unsigned long long)cnt);
                if (MODE == 4 && ((__ldg(p.filter + (idx0 >> 5)) >> (idx0 & 31u)) & 1u)) {
                    const unsigned long long slot = atomicAdd(p.words_cursor, 1ull);
                    if (slot < p.words_cap) { p.words[slot] = 0ull; p.widx[slot] = idx0; }
                }
            }
            if (m == 0xFFFFFFFFu) return;
            if (all_other) inv16 = 0xFFFFu;
        }
    }
#endif
    // MODE 2: reserve this warp's slots in the word array with ONE atomic per 512-position chunk
    unsigned long long* wslot = nullptr;
    unsigned int* islot = nullptr;
    unsigned wstride = 1u;
    if (MODE == 2) {
        const unsigned mine = 16u - __popc(inv16 & 0xFFFFu);
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        unsigned long long base = 0;
        if (lane == 31) base = atomicAdd(p.words_cursor, (unsigned long long)incl);
        base = __shfl_sync(0xFFFFFFFFu, base, 31);
        wslot = p.words + base + (incl - mine);
        islot = p.widx + base + (incl - mine);
        if (__all_sync(0xFFFFFFFFu, mine == 16u)) {
            // the common case: window j of lane l goes to slot 32 j + l, so that every store instruction of the warp
            // writes 32 consecutive words (the order inside the array means nothing: it is partitioned afterwards)
            wslot = p.words + base + lane;
            islot = p.widx + base + lane;
            wstride = 32u;
        }
    }
    constexpr int kUnroll = NK_COUNT_UNROLL;
    // word of window j of this lane (canonical min, or pack_kmer) — pure function of the code words
    auto window = [&](unsigned j, unsigned long long& fwd, unsigned long long& rc) -> unsigned long long {
        const unsigned sr = 32u - 2u * j;
        // KHI (k > 16): the low word is all window, only the high word needs the mask;
        // else the window fits the low word and the high word is zero
        const unsigned flo = KHI ? __funnelshift_rc(G2, G1, sr) : (__funnelshift_rc(G2, G1, sr) & wc.mask_lo);
        const unsigned fhi = KHI ? (__funnelshift_rc(G1, G0, sr) & wc.mask_hi) : 0u;
        fwd = ((unsigned long long)fhi << 32) | flo;
        rc = 0;
        if (CANON) {
            const unsigned rlo = KHI ? __funnelshift_r(R0, R1, 2u * j) : (__funnelshift_r(R0, R1, 2u * j) & wc.mask_lo);
            const unsigned rhi = KHI ? (__funnelshift_r(R1, R2, 2u * j) & wc.mask_hi) : 0u;
            rc = ((unsigned long long)rhi << 32) | rlo;
            return fwd < rc ? fwd : rc;  // src/models.rs:284-286
        }
        const unsigned vlo = KHI ? __funnelshift_rc(H2, H1, sr) : (__funnelshift_rc(H2, H1, sr) & wc.mask_lo);
        const unsigned vhi = KHI ? (__funnelshift_rc(H1, H0, sr) & wc.mask_hi) : 0u;
        // all k bytes ACGT: pack_kmer == forward word; otherwise the skip rule (src/utils.rs:35)
        if (vlo != wc.mask_lo || (KHI && vhi != wc.mask_hi)) return pack_skip(F0, F1, F2, V0, V1, V2, j, wc.k);
        return fwd;
    };

    if constexpr (MODE == 3) {
        // Short-read batches: ~(k-1)/L of the window starts are read ends (20 % at 150 bp, k=31).
        // Hashing them and predicating the RED away wastes ALU-pipe slots, so the warp first
        // compacts the words of its VALID starts into shared memory, then hashes them densely.
        const unsigned mine = 16u - __popc(inv16 & 0xFFFFu);
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (unsigned)o) incl += t;
        }
        const unsigned total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        unsigned slot = incl - mine;
#pragma unroll 4
        for (unsigned j = 0; j < 16; ++j) {
            unsigned long long fwd, rc;
            const unsigned long long word = window(j, fwd, rc);
            if (!((inv16 >> j) & 1u)) wbuf[slot++] = word;
        }
        __syncwarp();
        const unsigned full = total & ~127u;  // groups of 4 x 32 words in which every lane has work
        for (unsigned t = lane; t < full; t += 128u) {
#pragma unroll
            for (unsigned u = 0; u < 4; ++u) {
                const unsigned long long word = wbuf[t + 32u * u];
                const U64 h = siphash13_dev((unsigned)word, (unsigned)(word >> 32), p.rm);
                // (red, not atomicAdd: the compiler emitted ATOMG with a discarded return value here, which the L2 answers)
                asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p.acc + fastmod_kind_dev<MODK>(h, p.fm)), "r"(1u) : "memory");
            }
        }
        for (unsigned t = full + lane; t < total; t += 32u) {
            const unsigned long long word = wbuf[t];
            const U64 h = siphash13_dev((unsigned)word, (unsigned)(word >> 32), p.rm);
            asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p.acc + fastmod_kind_dev<MODK>(h, p.fm)), "r"(1u) : "memory");
        }
        __syncwarp();
    } else {
#ifndef NK_NO_FASTVALID
    // a warp whose 512 starts are all valid (the common case in long sequences) skips the per-window validity bit and
    // the predicate on the reduction (-0.8 % kernel time, profiles/r02_variants.md)
    if (MODE == 0 && __all_sync(0xFFFFFFFFu, inv16 == 0u)) {
#pragma unroll(kUnroll)
        for (unsigned j = 0; j < 16; ++j) {
            unsigned long long fwd, rc;
            const unsigned long long word = window(j, fwd, rc);
            const U64 h = siphash13_dev((unsigned)word, (unsigned)(word >> 32), p.rm);
            const unsigned idx = fastmod_kind_dev<MODK>(h, p.fm);
#ifdef NK_EXP_NORED
            if (idx == 0xFFFFFFFFu) p.acc[0] = 1u;  // diagnostic build only (tools/variants.sh): no pool update
#else
            asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p.acc + idx), "r"(1u) : "memory");
#endif
        }
        return;
    }
#endif
#pragma unroll(kUnroll)
    for (unsigned j = 0; j < 16; ++j) {
        unsigned long long fwd, rc;
        const unsigned long long word = window(j, fwd, rc);
        const unsigned bad = (inv16 >> j) & 1u;
        const U64 h = siphash13_dev((unsigned)word, (unsigned)(word >> 32), p.rm);
        const unsigned idx = fastmod_kind_dev<MODK>(h, p.fm);
        if (EMIT) {
            if (!bad) {
                const unsigned long long pos = pos0 + j;
                if (p.out_fwd) p.out_fwd[pos] = fwd;
                if (p.out_rc) p.out_rc[pos] = rc;
                if (p.out_word) p.out_word[pos] = word;
                if (p.out_idx) p.out_idx[pos] = idx;
            }
        } else if (MODE == 4) {
            // uniques pass: keep the words that map to a neuron of the filter (rare: a few rows of millions)
            if (!bad && ((__ldg(p.filter + (idx >> 5)) >> (idx & 31u)) & 1u)) {
                const unsigned long long slot = atomicAdd(p.words_cursor, 1ull);
                if (slot < p.words_cap) { p.words[slot] = word; p.widx[slot] = idx; }
            }
        } else {
            if (MODE == 2 && !bad) { *wslot = word; *islot = idx; wslot += wstride; islot += wstride; }
#ifdef NK_EXP_NORED
            // diagnostic build only (tools/variants.sh): no pool update, keep the value alive
            if (idx == 0xFFFFFFFFu) p.acc[0] = bad;
#elif defined(NK_EXP_ADDR32)
            // experiment: 64-bit address by an explicit 32-bit carry chain instead of IMAD.WIDE
            {
                const unsigned long long base = reinterpret_cast<unsigned long long>(p.acc);
                unsigned alo, ahi;
                asm("add.cc.u32 %0, %2, %4;\n\taddc.u32 %1, %3, 0;" : "=r"(alo), "=r"(ahi)
                    : "r"((unsigned)base), "r"((unsigned)(base >> 32)), "r"(idx << 2));
                const unsigned long long addr = ((unsigned long long)ahi << 32) | alo;
                asm volatile(
                    "{\n\t.reg .pred q;\n\t"
                    "setp.eq.u32 q, %2, 0;\n\t"
                    "@q red.global.add.u32 [%0], %1;\n\t}" ::"l"(addr), "r"(1u), "r"(bad)
                    : "memory");
            }
#elif defined(NK_EXP_REDVAL)
            // experiment: unconditional RED of (1 - bad): no predicate, no branch region
            asm volatile("red.global.add.u32 [%0], %1;" ::"l"(p.acc + idx), "r"(1u - bad) : "memory");
#else
            // predicated RED.E.ADD (no divergence region around a single instruction)
            asm volatile(
                "{\n\t.reg .pred q;\n\t"
                "setp.eq.u32 q, %2, 0;\n\t"
                "@q red.global.add.u32 [%0], %1;\n\t}" ::"l"(p.acc + idx),
                "r"(1u), "r"(bad)
                : "memory");
#endif
        }
    }
    }  // MODE != 3
}

#ifndef NK_COUNT_MINBLOCKS
#define NK_COUNT_MINBLOCKS 2
#endif
// (Round 2 also tried a barrier-free tile pipeline here — `full` / `empty` mbarriers per stage, 2-4 stages, every warp
// converting its own halo so that no __syncthreads is left in the loop.  It measured 1.0-1.9 % SLOWER than a loop with
// barriers (profiles/r02_variants.md): the ALU pipe was busy with the other CTAs' warps anyway.  Kept out.)
template <bool CANON, int MODE, int MODK, bool KHI, bool PACKED>
__global__ void __launch_bounds__(COUNT_THREADS, NK_COUNT_MINBLOCKS) count_kernel(const __grid_constant__ CountParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[COUNT_STAGES];
    constexpr bool BITMAP = MODE != 5;
    unsigned valid_starts = 0;  // MODE 5: window starts this thread counted
    const WindowConsts wc = make_window_consts(p.k);
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const bool has_other = PACKED && p.other != nullptr;

    // ONE __syncthreads per tile.  Everything a warp reads from a stage (its 16 bytes per lane, the halo, the bitmap
    // word) is in registers before the barrier, so right after it thread 0 refills the stage with the tile after
    // next, and the warps hash without meeting again until the next tile's barrier.  The tile ticket (a global atomic
    // that queues behind this kernel's own reductions in L2) is fetched one tile ahead of its use, and the sequence
    // ends of a tile are looked up from the previous tile's position: the serial work of thread 0 that the other
    // seven warps wait for at the barrier shrinks to a few hundred cycles per tile.  Worth 0.4 % against the loop with a
    // barrier on each side of the hashing (0.783 against 0.786 ms, profiles/r02_variants.md).
    static_assert(COUNT_CHUNKS_PER_SPAN == 1 && COUNT_STAGES == 2, "the one-barrier loop reads a stage once, before the barrier");
    constexpr unsigned kSlots = 4;                        // per-tile records live for three tiles
    __shared__ unsigned long long tile_of[kSlots];
    __shared__ TileEnds tends[kSlots];
    __shared__ unsigned int halo[COUNT_STAGES][COUNT_WARPS][2][3];
    __shared__ unsigned long long cursor;                 // thread 0 only: where the previous tile's end search stopped
    unsigned ticket = 0;                                  // thread 0 only: the ticket of tile it+2, in flight in a register

    if (threadIdx.x == 0) {
        for (int s = 0; s < COUNT_STAGES; ++s) mbar_init(&bars[s], 1);
        mbar_fence_init();
        cursor = 0;
        for (unsigned s = 0; s < 2; ++s) {
            const unsigned long long t = atomicAdd(p.tile_counter, 1u);
            tile_of[s] = t;
            if (t < p.ntiles) {
                issue_tile<PACKED, BITMAP>(smem + s * kStageStride, &bars[s], p, t);
                if (!BITMAP) find_tile_ends(p, t, &tends[s], &cursor);
            }
        }
        ticket = atomicAdd(p.tile_counter, 1u);
    }
    __syncthreads();

    for (unsigned it = 0;; ++it) {
        const unsigned s = it & 1u, slot = it & (kSlots - 1u);
        const unsigned long long tile = tile_of[slot];
        if (tile >= p.ntiles) break;
        mbar_wait(&bars[s], (it >> 1) & 1u);

        const unsigned char* sb = smem + s * kStageStride;
        const unsigned short* bits =
            reinterpret_cast<const unsigned short*>(sb + (PACKED ? kPkCodes + kPkOther : kBytesPerStage));
        const unsigned span0 = warp * COUNT_SPAN;
        const Codes16 cur = load_codes<CANON, PACKED>(sb, (span0 >> 4) + lane, has_other);
        // The k-1 overlap past the END of a warp's span is the first two code words of the next warp's span: they
        // are exchanged through shared memory instead of being converted twice (the last warp converts the tile's
        // 32-byte halo).  This is what makes small tiles cheap, and small tiles keep the persistent grid's tail short.
        if (lane < 2) {
            halo[s][warp][lane][0] = cur.F;
            halo[s][warp][lane][1] = cur.R;
            halo[s][warp][lane][2] = cur.V;
        }
        Codes16 tail{0u, 0u, 0u};
        if (warp == COUNT_WARPS - 1)
            tail = load_codes<CANON, PACKED>(sb, COUNT_TILE / 16 + (lane & 1u), has_other);
        unsigned inv16 = 0;
        if (BITMAP) inv16 = bits[(span0 >> 4) + lane];
        __syncthreads();
        if (threadIdx.x == 0) {  // stage s is free: refill it with tile it+2; ask for the ticket of tile it+3
            const unsigned long long tn = ticket;
            ticket = atomicAdd(p.tile_counter, 1u);
            tile_of[(it + 2u) & (kSlots - 1u)] = tn;
            if (tn < p.ntiles) {
                issue_tile<PACKED, BITMAP>(smem + s * kStageStride, &bars[s], p, tn);
                if (!BITMAP) find_tile_ends(p, tn, &tends[(it + 2u) & (kSlots - 1u)], &cursor);
            }
        }
        if (warp < COUNT_WARPS - 1 && lane < 2) {
            tail.F = halo[s][warp + 1][lane][0];
            tail.R = halo[s][warp + 1][lane][1];
            tail.V = halo[s][warp + 1][lane][2];
        }
        if (!BITMAP) {
            inv16 = invalid_bits(p, tends[slot], tile, (int)(span0 + 16u * lane));
            valid_starts += 16u - __popc(inv16);
        }
        process_chunk<CANON, (MODE == 5 ? 0 : MODE), MODK, KHI>(p, wc, cur, tail, inv16, lane, tile * COUNT_TILE + span0 + 16u * lane,
                                              reinterpret_cast<unsigned long long*>(smem + kSmemTotal) + warp * COUNT_CHUNK);
    }
    if (!BITMAP) {  // the metric's unit: windows counted (the bitmap path's marking kernel does this otherwise)
        valid_starts = __reduce_add_sync(0xFFFFFFFFu, valid_starts);
        if (lane == 0 && valid_starts) atomicAdd(p.kmers_out, (unsigned long long)valid_starts);
    }
    // the last CTA to leave re-arms the tile scheduler for the next launch
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(p.tile_counter + 1, 1u) == gridDim.x - 1u) {
            p.tile_counter[0] = 0u;
            p.tile_counter[1] = 0u;
        }
    }
}

// One thread per sequence in [seq_lo, seq_hi): starts max(start, end-k+1) .. end-1 are invalid.
// Also counts the windows that start inside this chunk (the metric's unit) into *kmers.
__global__ void mark_seq_ends_kernel(unsigned int* invalid, const unsigned long long* __restrict__ offsets,
                                     unsigned long long seq_lo, unsigned long long seq_hi,
                                     unsigned long long origin, unsigned long long nbytes, unsigned k,
                                     unsigned long long* kmers) {
    unsigned long long s = seq_lo + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long mine = 0;
    if (s < seq_hi) {
        unsigned long long a = offsets[s], e = offsets[s + 1];
        const unsigned long long cend = origin + nbytes;
        if (e - a >= k) {  // windows start at a .. e-k
            const unsigned long long w0 = a > origin ? a : origin;
            const unsigned long long w1 = (e - k + 1) < cend ? (e - k + 1) : cend;
            if (w1 > w0) mine = w1 - w0;
        }
        unsigned long long lo = (e - a >= (unsigned long long)(k - 1)) ? e - (k - 1) : a;
        // clip to this chunk [origin, origin + nbytes)
        if (lo < origin) lo = origin;
        if (e > cend) e = cend;
        if (lo < e) {
            lo -= origin; e -= origin;
            // at most 31 bits: spans at most two words
            unsigned long long wlo = lo >> 5, whi = (e - 1) >> 5;
            unsigned blo = (unsigned)(lo & 31), bhi = (unsigned)((e - 1) & 31);
            if (wlo == whi) {
                unsigned m = (bhi == 31 ? 0xFFFFFFFFu : ((1u << (bhi + 1)) - 1u)) & ~((1u << blo) - 1u);
                atomicOr(invalid + wlo, m);
            } else {
                atomicOr(invalid + wlo, ~((1u << blo) - 1u));
                atomicOr(invalid + whi, bhi == 31 ? 0xFFFFFFFFu : ((1u << (bhi + 1)) - 1u));
            }
        }
    }
    // one atomic per block on the k-mer total (a per-warp atomic to this single address cost
    // ~0.5 ms per 10 M reads)
    __shared__ unsigned long long s_k[8];
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_down_sync(0xFFFFFFFFu, mine, o);
    if ((threadIdx.x & 31) == 0) s_k[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (unsigned w = 0; w < (blockDim.x + 31) / 32; ++w) t += s_k[w];
        if (t) atomicAdd(kmers, t);
    }
}

// Everything in [nbytes, nbits_total) is invalid (tile padding).
__global__ void mark_tail_kernel(unsigned int* invalid, unsigned long long nbytes, unsigned long long nbits_total) {
    unsigned long long w = (nbytes >> 5) + blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    unsigned long long wend = (nbits_total + 31) >> 5;
    if (w >= wend) return;
    unsigned m = 0xFFFFFFFFu;
    if (w == (nbytes >> 5)) m = ~((1u << (nbytes & 31)) - 1u);
    atomicOr(invalid + w, m);
}

}  // namespace

size_t count_smem_bytes() { return (size_t)kSmemTotal; }

template <bool CANON, int MODE, int MODK, bool KHI, bool PACKED>
static cudaError_t launch_count_t(const CountParams& p, cudaStream_t s) {
    constexpr int kSmem = kSmemTotal + (MODE == 3 ? kCompactBytes : 0);
    // persistent grid of this instantiation: SMs x resident CTAs (a multiple of 148 on B200); function
    // attributes are per device, so the cache is too
    // (distinct handles may launch from different threads: the cache is atomic, its fill serialised)
    static std::atomic<int> max_grid_of[64];
    static std::mutex fill_mu;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (max_grid_of[dev].load(std::memory_order_acquire) == 0) {
        std::lock_guard<std::mutex> lk(fill_mu);
        int sms = 0, per_sm = 0;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(count_kernel<CANON, MODE, MODK, KHI, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e != cudaSuccess) return e;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, count_kernel<CANON, MODE, MODK, KHI, PACKED>, COUNT_THREADS, kSmem);
        if (e != cudaSuccess) return e;
        max_grid_of[dev].store(sms * (per_sm < 1 ? 1 : per_sm), std::memory_order_release);
    }
    const int max_grid = max_grid_of[dev].load(std::memory_order_acquire);
    const unsigned long long grid = p.ntiles < (unsigned long long)max_grid ? p.ntiles : (unsigned long long)max_grid;
    if (grid == 0) return cudaSuccess;
    count_kernel<CANON, MODE, MODK, KHI, PACKED><<<(unsigned)grid, COUNT_THREADS, kSmem, s>>>(p);
    return cudaGetLastError();
}

template <bool CANON, int MODE, int MODK, bool PACKED>
static cudaError_t launch_count_k(const CountParams& p, cudaStream_t s) {
    return p.k > 16 ? launch_count_t<CANON, MODE, MODK, true, PACKED>(p, s) : launch_count_t<CANON, MODE, MODK, false, PACKED>(p, s);
}

template <bool CANON, int MODE, bool PACKED>
static cudaError_t launch_count_cm(const CountParams& p, cudaStream_t s) {
    if (p.fm.is_pow2) return launch_count_k<CANON, MODE, 1, PACKED>(p, s);
    // the one-stage remainder (nk_device.cuh) for the modes a job's time is spent in; the tap and table modes keep
    // the two-stage form
    if constexpr (MODE == 0 || MODE == 3 || MODE == 5) {
        if (p.fm.kind == 3u) return launch_count_k<CANON, MODE, 3, PACKED>(p, s);
        if (p.fm.kind == 2u) return launch_count_k<CANON, MODE, 2, PACKED>(p, s);
    }
    return launch_count_k<CANON, MODE, 0, PACKED>(p, s);
}

template <bool CANON, bool PACKED>
static cudaError_t launch_count_c(const CountParams& p, int mode, cudaStream_t s) {
    switch (mode) {
        case 0: return launch_count_cm<CANON, 0, PACKED>(p, s);
        case 1: return launch_count_cm<CANON, 1, PACKED>(p, s);
        case 2: return launch_count_cm<CANON, 2, PACKED>(p, s);
        case 4: return launch_count_cm<CANON, 4, PACKED>(p, s);
        case 5: return launch_count_cm<CANON, 5, PACKED>(p, s);
        default: return launch_count_cm<CANON, 3, PACKED>(p, s);
    }
}

cudaError_t launch_count(const CountParams& p, bool canonical, int mode, cudaStream_t s) {
    if (p.packed)
        return canonical ? launch_count_c<true, true>(p, mode, s) : launch_count_c<false, true>(p, mode, s);
    return canonical ? launch_count_c<true, false>(p, mode, s) : launch_count_c<false, false>(p, mode, s);
}

cudaError_t launch_mark_invalid(unsigned int* invalid, const unsigned long long* offsets,
                                unsigned long long seq_lo, unsigned long long seq_hi,
                                unsigned long long origin, unsigned long long nbytes, unsigned k,
                                unsigned long long* kmers, cudaStream_t s, uint64_t* launches) {
    const unsigned long long words = count_bitmap_words(nbytes);
    cudaError_t e = cudaMemsetAsync(invalid, 0, words * sizeof(unsigned int), s);
    if (e != cudaSuccess) return e;
    const unsigned long long nseq = seq_hi - seq_lo;
    if (nseq > 0) {
        const unsigned long long blocks = (nseq + 255) / 256;
        mark_seq_ends_kernel<<<(unsigned)blocks, 256, 0, s>>>(invalid, offsets, seq_lo, seq_hi, origin, nbytes, k,
                                                              kmers);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        if (launches) ++*launches;
    }
    const unsigned long long nbits_total = count_ntiles(nbytes) * COUNT_TILE;
    if (nbits_total > nbytes) {
        const unsigned long long nwords = ((nbits_total + 31) >> 5) - (nbytes >> 5);
        mark_tail_kernel<<<(unsigned)((nwords + 255) / 256), 256, 0, s>>>(invalid, nbytes, nbits_total);
        e = cudaGetLastError();
        if (launches) ++*launches;
    }
    return e;
}

}  // namespace nk
