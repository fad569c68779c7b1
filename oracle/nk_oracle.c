/*
 * nk_oracle.c — CPU restatement of NeuroKmer's counting hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this file.  The product path
 * (neurokmer_b200/csrc) never links, imports or calls it.
 *
 * PARITY UNPINNED: the reference (a Rust crate) cannot be built in this
 * environment (no cargo/rustc), ships no golden vectors and its only test
 * asserts nothing (reference tests/test_counting.rs:73-184).  The pins this
 * oracle has instead are listed in DESIGN.md §3: SipHash-2-4 published test
 * vectors for the generic round function, CPython's own SipHash-1-3
 * (PYTHONHASHSEED=0) for the 1-3 variant with zero keys, a pure-Python twin
 * (oracle/oracle_py.py) written independently from the same reference lines,
 * and algebraic invariants.
 *
 * Every function cites the reference file:line it restates (paths relative
 * to the reference checkout).  siphasher 1.0.2 (Cargo.lock:1520-1523) is a
 * third-party crate that is not vendored; its algorithm is the published
 * SipHash (Aumasson & Bernstein 2012) with c=1, d=3.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NKO_EXPORT __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* 2-bit codes — src/models.rs:231-239 and :243-251                    */
/* ------------------------------------------------------------------ */
static inline uint64_t base_to_bits(uint8_t b) {
    switch (b) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 0; /* N and everything else read as A on the forward strand */
    }
}

static inline uint64_t base_to_complement_bits(uint8_t b) {
    switch (b) {
    case 'A': case 'a': return 3;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 1;
    case 'T': case 't': return 0;
    default: return 0; /* …but as 0 (not 3) on the reverse strand: models.rs:249 */
    }
}

/* RollingKmerHash — src/models.rs:177-299 */
typedef struct {
    unsigned k;
    uint64_t forward, reverse, mask, power;
} rolling_t;

/* src/models.rs:186-203 */
static void rolling_new(rolling_t *r, unsigned k) {
    r->k = k;
    r->mask = (k < 32) ? ((1ULL << (2 * k)) - 1) : ~0ULL;
    uint64_t power = 1;
    for (unsigned i = 0; i + 1 < k; i++) power = (power << 2) & r->mask;
    r->power = power;
    r->forward = r->reverse = 0;
}

/* src/models.rs:206-227 */
static void rolling_init(rolling_t *r, const uint8_t *first_k) {
    r->forward = 0;
    for (unsigned j = 0; j < r->k; j++)
        r->forward = ((r->forward << 2) & r->mask) | base_to_bits(first_k[j]);
    r->reverse = 0;
    for (unsigned j = r->k; j-- > 0;)
        r->reverse = ((r->reverse << 2) & r->mask) | base_to_complement_bits(first_k[j]);
}

/* src/models.rs:254-269 */
static void rolling_slide(rolling_t *r, uint8_t next_base, uint8_t prev_base) {
    uint64_t prev_bits = base_to_bits(prev_base);
    uint64_t next_bits = base_to_bits(next_base);
    r->forward = r->forward - prev_bits * r->power; /* wrapping_sub */
    r->forward = ((r->forward << 2) | next_bits) & r->mask;
    uint64_t comp_next = base_to_complement_bits(next_base);
    r->reverse = (r->reverse >> 2) | (comp_next << (2 * (r->k - 1)));
    r->reverse &= r->mask;
}

/* src/utils.rs:26-39 — non-canonical pack: non-ACGT bytes are skipped, no mask */
NKO_EXPORT uint64_t nko_pack_kmer(const uint8_t *kmer, uint64_t len) {
    uint64_t packed = 0;
    for (uint64_t i = 0; i < len; i++) {
        uint64_t bits;
        switch (kmer[i]) {
        case 'A': case 'a': bits = 0; break;
        case 'C': case 'c': bits = 1; break;
        case 'G': case 'g': bits = 2; break;
        case 'T': case 't': bits = 3; break;
        default: continue;
        }
        packed = (packed << 2) | bits;
    }
    return packed;
}

/* ------------------------------------------------------------------ */
/* SipHash — siphasher 1.0.2 `SipHasher13::new_with_keys`, called at   */
/* src/spiking_hash.rs:78-82 and :313-317.                             */
/* ------------------------------------------------------------------ */
#define ROTL64(x, b) (((x) << (b)) | ((x) >> (64 - (b))))
#define SIPROUND                                                        \
    do {                                                                \
        v0 += v1; v1 = ROTL64(v1, 13); v1 ^= v0; v0 = ROTL64(v0, 32);   \
        v2 += v3; v3 = ROTL64(v3, 16); v3 ^= v2;                        \
        v0 += v3; v3 = ROTL64(v3, 21); v3 ^= v0;                        \
        v2 += v1; v1 = ROTL64(v1, 17); v1 ^= v2; v2 = ROTL64(v2, 32);   \
    } while (0)

/* Generic SipHash-c-d over a byte string (so the 2-4 published vectors can
 * pin the round function and padding rule). */
NKO_EXPORT uint64_t nko_siphash(unsigned c_rounds, unsigned d_rounds, uint64_t k0, uint64_t k1,
                                const uint8_t *in, uint64_t inlen) {
    uint64_t v0 = 0x736f6d6570736575ULL ^ k0;
    uint64_t v1 = 0x646f72616e646f6dULL ^ k1;
    uint64_t v2 = 0x6c7967656e657261ULL ^ k0;
    uint64_t v3 = 0x7465646279746573ULL ^ k1;
    const uint8_t *end = in + (inlen - (inlen % 8));
    uint64_t b = inlen << 56;
    for (; in != end; in += 8) {
        uint64_t m = 0;
        for (int i = 7; i >= 0; i--) m = (m << 8) | in[i]; /* little endian */
        v3 ^= m;
        for (unsigned i = 0; i < c_rounds; i++) SIPROUND;
        v0 ^= m;
    }
    for (int i = (int)(inlen % 8) - 1; i >= 0; i--) b |= (uint64_t)in[i] << (8 * i);
    v3 ^= b;
    for (unsigned i = 0; i < c_rounds; i++) SIPROUND;
    v0 ^= b;
    v2 ^= 0xff;
    for (unsigned i = 0; i < d_rounds; i++) SIPROUND;
    return v0 ^ v1 ^ v2 ^ v3;
}

/* `packed.hash(&mut hasher)` on a u64 feeds its 8 native-endian (LE on
 * x86-64, the only arch the reference builds on) bytes, no length prefix. */
NKO_EXPORT uint64_t nko_siphash13_u64(uint64_t x) {
    uint8_t le[8];
    for (int i = 0; i < 8; i++) le[i] = (uint8_t)(x >> (8 * i));
    return nko_siphash(1, 3, 0, 0, le, 8);
}

/* src/spiking_hash.rs:78-82 */
NKO_EXPORT uint64_t nko_neuron_index(uint64_t packed, uint64_t pool_size) {
    return nko_siphash13_u64(packed) % pool_size;
}

/* ------------------------------------------------------------------ */
/* Window emission — loops at src/spiking_hash.rs:102-139 / :322-352   */
/* ------------------------------------------------------------------ */
/* Writes one word per window of `seq`; returns the number of windows
 * (max(0, len-k+1)).  canonical != 0 → rolling min(fwd, rc); else pack_kmer. */
NKO_EXPORT uint64_t nko_kmer_words(const uint8_t *seq, uint64_t len, unsigned k, int canonical,
                                   uint64_t *out) {
    if (len < k) return 0; /* :102 / :206-208 / :322-324 (windows() of a short slice is empty) */
    uint64_t n = len - k + 1;
    if (canonical) {
        rolling_t r;
        rolling_new(&r, k);
        rolling_init(&r, seq);
        out[0] = r.forward < r.reverse ? r.forward : r.reverse;
        for (uint64_t i = 1; i <= len - k; i++) {
            rolling_slide(&r, seq[i + k - 1], seq[i - 1]);
            out[i] = r.forward < r.reverse ? r.forward : r.reverse;
        }
    } else {
        for (uint64_t i = 0; i < n; i++) out[i] = nko_pack_kmer(seq + i, k);
    }
    return n;
}

/* Same as above but also returns forward and reverse words (debug tap). */
NKO_EXPORT uint64_t nko_kmer_fwd_rc(const uint8_t *seq, uint64_t len, unsigned k, uint64_t *fwd,
                                    uint64_t *rc) {
    if (len < k) return 0;
    rolling_t r;
    rolling_new(&r, k);
    rolling_init(&r, seq);
    fwd[0] = r.forward; rc[0] = r.reverse;
    for (uint64_t i = 1; i <= len - k; i++) {
        rolling_slide(&r, seq[i + k - 1], seq[i - 1]);
        fwd[i] = r.forward; rc[i] = r.reverse;
    }
    return len - k + 1;
}

/* Exact-count map of the reference: `local_counts: HashMap<u64, u32>` per rayon task
 * (src/spiking_hash.rs:96,110,124,134) merged into `counts: DashMap<u64, AtomicU32>` (:157-165).
 * Open addressing, linear probing, grown at 60 % load; a slot is empty while its count is 0. */
typedef struct {
    uint64_t *keys;
    uint32_t *vals;
    uint64_t cap, n; /* cap is a power of two */
} kmap_t;

static void kmap_init(kmap_t *m, uint64_t cap) {
    m->cap = cap;
    m->n = 0;
    m->keys = (uint64_t *)malloc(cap * sizeof(uint64_t));
    m->vals = (uint32_t *)calloc(cap, sizeof(uint32_t));
}
static void kmap_free(kmap_t *m) { free(m->keys); free(m->vals); m->keys = NULL; m->vals = NULL; }
static inline uint64_t kmap_hash(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static void kmap_add(kmap_t *m, uint64_t key, uint32_t by);
static void kmap_grow(kmap_t *m) {
    kmap_t big;
    kmap_init(&big, m->cap * 2);
    for (uint64_t i = 0; i < m->cap; i++)
        if (m->vals[i]) kmap_add(&big, m->keys[i], m->vals[i]);
    kmap_free(m);
    *m = big;
}
static void kmap_add(kmap_t *m, uint64_t key, uint32_t by) {
    if ((m->n + 1) * 5 > m->cap * 3) kmap_grow(m);
    uint64_t i = kmap_hash(key) & (m->cap - 1);
    for (;;) {
        if (!m->vals[i]) { m->keys[i] = key; m->vals[i] = by; m->n++; return; }
        if (m->keys[i] == key) { m->vals[i] += by; return; } /* wraps like AtomicU32::fetch_add */
        i = (i + 1) & (m->cap - 1);
    }
}

/* currents[idx] += 1 for every window of every sequence in [s_lo, s_hi) —
 * the body of the rayon fold (src/spiking_hash.rs:98-143) and of the
 * streaming worker (:319-352), minus the exact-count HashMap. */
static uint64_t accumulate_range(const uint8_t *bases, const uint64_t *offsets, uint64_t s_lo,
                                 uint64_t s_hi, unsigned k, uint64_t pool_size, int canonical,
                                 uint64_t *currents) {
    uint64_t total = 0;
    for (uint64_t s = s_lo; s < s_hi; s++) {
        const uint8_t *seq = bases + offsets[s];
        uint64_t len = offsets[s + 1] - offsets[s];
        if (len < k) continue;
        if (canonical) {
            rolling_t r;
            rolling_new(&r, k);
            rolling_init(&r, seq);
            uint64_t w = r.forward < r.reverse ? r.forward : r.reverse;
            currents[nko_neuron_index(w, pool_size)] += 1;
            for (uint64_t i = 1; i <= len - k; i++) {
                rolling_slide(&r, seq[i + k - 1], seq[i - 1]);
                w = r.forward < r.reverse ? r.forward : r.reverse;
                currents[nko_neuron_index(w, pool_size)] += 1;
            }
        } else {
            for (uint64_t i = 0; i + k <= len; i++)
                currents[nko_neuron_index(nko_pack_kmer(seq + i, k), pool_size)] += 1;
        }
        total += len - k + 1;
    }
    return total;
}

/* The same WITH the exact-count HashMap — a separate copy, so that the hot-path-only baseline above
 * is compiled exactly as it was before this variant existed. */
static uint64_t accumulate_range_exact(const uint8_t *bases, const uint64_t *offsets, uint64_t s_lo,
                                       uint64_t s_hi, unsigned k, uint64_t pool_size, int canonical,
                                       uint64_t *currents, kmap_t *map) {
    uint64_t total = 0;
    for (uint64_t s = s_lo; s < s_hi; s++) {
        const uint8_t *seq = bases + offsets[s];
        uint64_t len = offsets[s + 1] - offsets[s];
        if (len < k) continue;
        if (canonical) {
            rolling_t r;
            rolling_new(&r, k);
            rolling_init(&r, seq);
            uint64_t w = r.forward < r.reverse ? r.forward : r.reverse;
            currents[nko_neuron_index(w, pool_size)] += 1;
            kmap_add(map, w, 1);
            for (uint64_t i = 1; i <= len - k; i++) {
                rolling_slide(&r, seq[i + k - 1], seq[i - 1]);
                w = r.forward < r.reverse ? r.forward : r.reverse;
                currents[nko_neuron_index(w, pool_size)] += 1;
                kmap_add(map, w, 1);
            }
        } else {
            for (uint64_t i = 0; i + k <= len; i++) {
                const uint64_t w = nko_pack_kmer(seq + i, k);
                currents[nko_neuron_index(w, pool_size)] += 1;
                kmap_add(map, w, 1);
            }
        }
        total += len - k + 1;
    }
    return total;
}

/* Single-threaded accumulate over a batch: `currents` is ADDED to. */
NKO_EXPORT uint64_t nko_accumulate(const uint8_t *bases, const uint64_t *offsets, uint64_t nseq,
                                   unsigned k, uint64_t pool_size, int canonical,
                                   uint64_t *currents) {
    return accumulate_range(bases, offsets, 0, nseq, k, pool_size, canonical, currents);
}

/* ------------------------------------------------------------------ */
/* Multi-threaded accumulate (CPU baseline).  Structure of the rayon   */
/* fold/reduce at src/spiking_hash.rs:94-154: private u64 pools per    */
/* worker, element-wise sum at the end.  It is MORE parallel than the  */
/* reference: long sequences are also cut into chunks with a k-1 halo  */
/* (the reference schedules one task per whole sequence).              */
/* ------------------------------------------------------------------ */
typedef struct {
    const uint8_t *bases;
    uint64_t lo, hi; /* window-start range [lo, hi) inside one sequence piece */
    uint64_t seq_end;
} piece_t;

typedef struct {
    const uint8_t *bases;
    const piece_t *pieces;
    uint64_t npieces;
    uint64_t *next; /* shared atomic cursor */
    unsigned k;
    uint64_t pool_size;
    int canonical;
    uint64_t *currents; /* private */
    uint64_t total;
    kmap_t *map;        /* private exact-count map, or NULL (hot path only) */
    /* merge phase of the exact variant: thread t owns the keys with kmap_hash(key) >> 40 % nthreads == t */
    kmap_t *all_maps;
    unsigned nthreads, tid;
    kmap_t merged;
} mt_arg_t;

static void *mt_worker(void *p) {
    mt_arg_t *a = (mt_arg_t *)p;
    for (;;) {
        uint64_t i = __atomic_fetch_add(a->next, 1, __ATOMIC_RELAXED);
        if (i >= a->npieces) break;
        const piece_t *pc = &a->pieces[i];
        /* a piece is the sub-sequence [lo, min(hi+k-1, seq_end)) processed as
         * if it were a sequence of its own: identical windows, no double count */
        uint64_t end = pc->hi + a->k - 1;
        if (end > pc->seq_end) end = pc->seq_end;
        uint64_t off[2] = {pc->lo, end};
        if (a->map)
            a->total += accumulate_range_exact(a->bases, off, 0, 1, a->k, a->pool_size, a->canonical, a->currents, a->map);
        else
            a->total += accumulate_range(a->bases, off, 0, 1, a->k, a->pool_size, a->canonical, a->currents);
    }
    return NULL;
}

NKO_EXPORT uint64_t nko_accumulate_mt(const uint8_t *bases, const uint64_t *offsets,
                                      uint64_t nseq, unsigned k, uint64_t pool_size,
                                      int canonical, uint64_t *currents, unsigned nthreads) {
    const uint64_t CHUNK = 1u << 20;
    if (nthreads == 0) nthreads = 1;
    /* build pieces */
    uint64_t npieces = 0;
    for (uint64_t s = 0; s < nseq; s++) {
        uint64_t len = offsets[s + 1] - offsets[s];
        if (len < k) continue;
        npieces += (len - k + 1 + CHUNK - 1) / CHUNK;
    }
    piece_t *pieces = (piece_t *)malloc((npieces ? npieces : 1) * sizeof(piece_t));
    uint64_t ip = 0;
    for (uint64_t s = 0; s < nseq; s++) {
        uint64_t len = offsets[s + 1] - offsets[s];
        if (len < k) continue;
        uint64_t nwin = len - k + 1;
        for (uint64_t w = 0; w < nwin; w += CHUNK) {
            pieces[ip].bases = bases;
            pieces[ip].lo = offsets[s] + w;
            pieces[ip].hi = offsets[s] + (w + CHUNK < nwin ? w + CHUNK : nwin);
            pieces[ip].seq_end = offsets[s + 1];
            ip++;
        }
    }
    uint64_t next = 0, total = 0;
    pthread_t *th = (pthread_t *)malloc(nthreads * sizeof(pthread_t));
    mt_arg_t *args = (mt_arg_t *)calloc(nthreads, sizeof(mt_arg_t));
    for (unsigned t = 0; t < nthreads; t++) {
        args[t].bases = bases; args[t].pieces = pieces; args[t].npieces = npieces;
        args[t].next = &next; args[t].k = k; args[t].pool_size = pool_size;
        args[t].canonical = canonical;
        args[t].currents = (t == 0) ? currents : (uint64_t *)calloc(pool_size, sizeof(uint64_t));
        pthread_create(&th[t], NULL, mt_worker, &args[t]);
    }
    for (unsigned t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        total += args[t].total;
        if (t > 0) { /* reduce: src/spiking_hash.rs:145-154 */
            for (uint64_t i = 0; i < pool_size; i++) currents[i] += args[t].currents[i];
            free(args[t].currents);
        }
    }
    free(args); free(th); free(pieces);
    return total;
}

/* The same with the reference's exact side tables (the work its CPU path ALWAYS does):
 * per-task HashMap<u64,u32> (:96), merged into the global `counts` map (:157-165), then
 * kmer_per_neuron = distinct k-mers per neuron (:167-172).  The merge is key-partitioned
 * over the threads (the reference's DashMap shards play that role).  *n_distinct = |counts|;
 * uniques (pool_size u32, may be NULL) receives kmer_per_neuron. */
static void *merge_worker(void *p) {
    mt_arg_t *a = (mt_arg_t *)p;
    kmap_init(&a->merged, 1u << 16);
    for (unsigned t = 0; t < a->nthreads; t++) {
        const kmap_t *m = &a->all_maps[t];
        for (uint64_t i = 0; i < m->cap; i++)
            if (m->vals[i] && (kmap_hash(m->keys[i]) >> 40) % a->nthreads == a->tid)
                kmap_add(&a->merged, m->keys[i], m->vals[i]);
    }
    return NULL;
}

NKO_EXPORT uint64_t nko_accumulate_exact_mt(const uint8_t *bases, const uint64_t *offsets,
                                            uint64_t nseq, unsigned k, uint64_t pool_size,
                                            int canonical, uint64_t *currents, unsigned nthreads,
                                            uint64_t *n_distinct, uint32_t *uniques) {
    const uint64_t CHUNK = 1u << 20;
    if (nthreads == 0) nthreads = 1;
    uint64_t npieces = 0;
    for (uint64_t s = 0; s < nseq; s++) {
        uint64_t len = offsets[s + 1] - offsets[s];
        if (len < k) continue;
        npieces += (len - k + 1 + CHUNK - 1) / CHUNK;
    }
    piece_t *pieces = (piece_t *)malloc((npieces ? npieces : 1) * sizeof(piece_t));
    uint64_t ip = 0;
    for (uint64_t s = 0; s < nseq; s++) {
        uint64_t len = offsets[s + 1] - offsets[s];
        if (len < k) continue;
        uint64_t nwin = len - k + 1;
        for (uint64_t w = 0; w < nwin; w += CHUNK) {
            pieces[ip].bases = bases;
            pieces[ip].lo = offsets[s] + w;
            pieces[ip].hi = offsets[s] + (w + CHUNK < nwin ? w + CHUNK : nwin);
            pieces[ip].seq_end = offsets[s + 1];
            ip++;
        }
    }
    uint64_t next = 0, total = 0;
    pthread_t *th = (pthread_t *)malloc(nthreads * sizeof(pthread_t));
    mt_arg_t *args = (mt_arg_t *)calloc(nthreads, sizeof(mt_arg_t));
    kmap_t *maps = (kmap_t *)calloc(nthreads, sizeof(kmap_t));
    for (unsigned t = 0; t < nthreads; t++) {
        kmap_init(&maps[t], 1u << 16);
        args[t].bases = bases; args[t].pieces = pieces; args[t].npieces = npieces;
        args[t].next = &next; args[t].k = k; args[t].pool_size = pool_size;
        args[t].canonical = canonical;
        args[t].currents = (t == 0) ? currents : (uint64_t *)calloc(pool_size, sizeof(uint64_t));
        args[t].map = &maps[t];
        pthread_create(&th[t], NULL, mt_worker, &args[t]);
    }
    for (unsigned t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        total += args[t].total;
        if (t > 0) {
            for (uint64_t i = 0; i < pool_size; i++) currents[i] += args[t].currents[i];
            free(args[t].currents);
        }
    }
    /* merge the per-task maps (counts), then the per-neuron distinct counts (kmer_per_neuron) */
    for (unsigned t = 0; t < nthreads; t++) {
        args[t].all_maps = maps; args[t].nthreads = nthreads; args[t].tid = t;
        pthread_create(&th[t], NULL, merge_worker, &args[t]);
    }
    uint64_t distinct = 0;
    for (unsigned t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        const kmap_t *m = &args[t].merged;
        distinct += m->n;
        if (uniques)
            for (uint64_t i = 0; i < m->cap; i++)
                if (m->vals[i]) uniques[nko_neuron_index(m->keys[i], pool_size)] += 1;
        kmap_free(&args[t].merged);
    }
    for (unsigned t = 0; t < nthreads; t++) kmap_free(&maps[t]);
    if (n_distinct) *n_distinct = distinct;
    free(maps); free(args); free(th); free(pieces);
    return total;
}

/* ------------------------------------------------------------------ */
/* LIF                                                                 */
/* ------------------------------------------------------------------ */
/* LifNeuron::update — src/models.rs:34-51.  The product and the sum are two
 * separately rounded f32 operations, as rustc emits them: this file MUST be
 * built with -ffp-contract=off (oracle/Makefile does; tests/test_oracle.py
 * checks an end voltage that an FMA would change). */
#if defined(__FP_FAST_FMA) && !defined(NKO_CONTRACT_OFF)
#error "build nk_oracle.c with -ffp-contract=off -DNKO_CONTRACT_OFF (see oracle/Makefile)"
#endif
static inline int lif_update(float *v, uint32_t *r, uint64_t *n, float thr, float leak,
                             uint32_t period, float input) {
    if (*r > 0) { *r -= 1; return 0; }
    float prod = *v * leak;
    *v = prod + input;
    if (*v >= thr) { *v = 0.0f; *r = period; *n += 1; return 1; }
    return 0;
}

/* In-memory driver — src/spiking_hash.rs:187-200 (≡ simulate_spikes :488-505):
 * neurons whose current is 0 are SKIPPED (state untouched). Returns new spikes. */
NKO_EXPORT uint64_t nko_lif_scalar(const uint64_t *currents, uint64_t lo, uint64_t hi,
                                   uint64_t steps, float thr, float leak, uint32_t period,
                                   float *v, uint32_t *r, uint64_t *spikes) {
    uint64_t fired = 0;
    for (uint64_t i = lo; i < hi; i++) {
        double total_current = (double)currents[i];
        if (total_current == 0.0) continue;
        double per_step = total_current / (double)steps;
        float input = (float)per_step;
        for (uint64_t t = 0; t < steps; t++)
            fired += lif_update(&v[i], &r[i], &spikes[i], thr, leak, period, input);
    }
    return fired;
}

/* Streaming driver — simulate_spikes_simd, src/spiking_hash.rs:544-659,
 * restated lane by lane: EVERY neuron is stepped (zero current included);
 * an inactive (refractory) lane compares its old voltage against f32::MAX
 * (:611-613); steps == 0 returns early (:548-551). */
NKO_EXPORT uint64_t nko_lif_simd_semantics(const uint64_t *currents, uint64_t lo, uint64_t hi,
                                           uint64_t steps, float thr, float leak,
                                           uint32_t period, float *v, uint32_t *r,
                                           uint64_t *spikes) {
    if (steps == 0) return 0;
    const float FMAX = 3.40282346638528859811704183484516925e+38f;
    uint64_t fired = 0;
    for (uint64_t i = lo; i < hi; i++) {
        float vv = v[i];
        uint32_t rr = r[i];
        uint64_t cnt = 0;
        float c = (float)((double)currents[i] / (double)steps);
        for (uint64_t t = 0; t < steps; t++) {
            int active = (rr == 0);
            float leaked = vv * leak;            /* _mm256_mul_ps :606 */
            float vnew = leaked + c;             /* _mm256_add_ps :607 */
            if (active) vv = vnew;               /* blendv :608 */
            float t_eff = active ? thr : FMAX;   /* :611-612 */
            int spike = (vv >= t_eff);           /* _CMP_GE_OQ :613 */
            if (spike) vv = 0.0f;                /* :617-618 */
            if (rr > 0) rr -= 1;                 /* :626-627 */
            else if (spike) rr = period;         /* :628-629 */
            else rr = 0;
            cnt += (uint64_t)spike;              /* :634 */
        }
        v[i] = vv; r[i] = rr; spikes[i] += cnt; fired += cnt;
    }
    return fired;
}

typedef struct {
    const uint64_t *currents; uint64_t lo, hi, steps; float thr, leak; uint32_t period;
    float *v; uint32_t *r; uint64_t *spikes; int simd; uint64_t fired;
} lif_arg_t;

static void *lif_worker(void *p) {
    lif_arg_t *a = (lif_arg_t *)p;
    a->fired = a->simd
        ? nko_lif_simd_semantics(a->currents, a->lo, a->hi, a->steps, a->thr, a->leak, a->period,
                                 a->v, a->r, a->spikes)
        : nko_lif_scalar(a->currents, a->lo, a->hi, a->steps, a->thr, a->leak, a->period, a->v,
                         a->r, a->spikes);
    return NULL;
}

/* Threaded over neuron ranges (the reference runs it on one thread; the
 * baseline is allowed to be faster than the reference, not slower). */
NKO_EXPORT uint64_t nko_lif_mt(const uint64_t *currents, uint64_t pool_size, uint64_t steps,
                               float thr, float leak, uint32_t period, float *v, uint32_t *r,
                               uint64_t *spikes, int simd_semantics, unsigned nthreads) {
    if (nthreads == 0) nthreads = 1;
    pthread_t *th = (pthread_t *)malloc(nthreads * sizeof(pthread_t));
    lif_arg_t *args = (lif_arg_t *)calloc(nthreads, sizeof(lif_arg_t));
    uint64_t per = (pool_size + nthreads - 1) / nthreads, fired = 0;
    for (unsigned t = 0; t < nthreads; t++) {
        uint64_t lo = per * t, hi = lo + per;
        if (lo > pool_size) lo = pool_size;
        if (hi > pool_size) hi = pool_size;
        lif_arg_t a = {currents, lo, hi, steps, thr, leak, period, v, r, spikes, simd_semantics, 0};
        args[t] = a;
        pthread_create(&th[t], NULL, lif_worker, &args[t]);
    }
    for (unsigned t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); fired += args[t].fired; }
    free(args); free(th);
    return fired;
}

/* EnergyTracker — src/models.rs:159-172, src/spiking_hash.rs:649-655 */
/* `(cost * 1000.0) as u64`: Rust's cast saturates (NaN -> 0, negative -> 0, >= 2^64 -> u64::MAX); the product
 * wraps (release-mode u64 multiply / repeated fetch_add). */
static uint64_t cost_fixed(double spike_cost) {
    const double x = spike_cost * 1000.0;
    if (!(x > 0.0)) return 0;
    if (x >= 18446744073709551616.0) return UINT64_MAX;
    return (uint64_t)x;
}
NKO_EXPORT uint64_t nko_energy_fixed(uint64_t new_spikes, double spike_cost) {
    return new_spikes * cost_fixed(spike_cost);
}
NKO_EXPORT double nko_energy_total(uint64_t fixed) { return (double)fixed / 1000.0; }

/* ------------------------------------------------------------------ */
/* Top-N — src/spiking_hash.rs:661-673: stable sort descending by      */
/* spike_count ⇒ ties keep ascending neuron index.                     */
/* ------------------------------------------------------------------ */
typedef struct { uint64_t idx, spikes; } top_t;
static int top_cmp(const void *a, const void *b) {
    const top_t *x = (const top_t *)a, *y = (const top_t *)b;
    if (x->spikes != y->spikes) return x->spikes > y->spikes ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);
}
NKO_EXPORT uint64_t nko_top_n(const uint64_t *spikes, uint64_t pool_size, uint64_t top_n,
                              uint64_t *out_idx, uint64_t *out_spikes) {
    top_t *all = (top_t *)malloc((pool_size ? pool_size : 1) * sizeof(top_t));
    for (uint64_t i = 0; i < pool_size; i++) { all[i].idx = i; all[i].spikes = spikes[i]; }
    qsort(all, pool_size, sizeof(top_t), top_cmp);
    uint64_t n = top_n < pool_size ? top_n : pool_size;
    for (uint64_t i = 0; i < n; i++) { out_idx[i] = all[i].idx; out_spikes[i] = all[i].spikes; }
    free(all);
    return n;
}

/* ------------------------------------------------------------------ */
/* process_sequence — src/spiking_hash.rs:203-273 (per-sequence API):  */
/* accumulate, then ONE tick per neuron with the raw count as current, */
/* currents zeroed afterwards.                                         */
/* ------------------------------------------------------------------ */
NKO_EXPORT uint64_t nko_process_sequence(const uint8_t *seq, uint64_t len, unsigned k,
                                         uint64_t pool_size, int canonical, float thr,
                                         float leak, uint32_t period, uint64_t *scratch_currents,
                                         float *v, uint32_t *r, uint64_t *spikes) {
    if (len < k) return 0;
    uint64_t off[2] = {0, len};
    accumulate_range(seq, off, 0, 1, k, pool_size, canonical, scratch_currents);
    uint64_t fired = 0;
    for (uint64_t i = 0; i < pool_size; i++) {
        double current = (double)scratch_currents[i];
        if (current > 0.0)
            fired += lif_update(&v[i], &r[i], &spikes[i], thr, leak, period, (float)current);
        scratch_currents[i] = 0;
    }
    return fired;
}
