// nk_device.cuh — device-side building blocks shared by the sm_100a kernels.
//
// Semantics follow the reference (paths relative to the reference checkout):
//   2-bit codes            src/models.rs:231-251
//   fwd / revcomp words    src/models.rs:206-269 (closed form of init + slide)
//   SipHash-1-3, keys 0,0  src/spiking_hash.rs:78-82 (siphasher 1.0.2)
//   % pool_size            src/spiking_hash.rs:81
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nk {

// ---------------------------------------------------------------------------
// ASCII -> 2-bit code streams, four bases per 32-bit word (SIMD within a register).
//
// For every byte b of `w`:
//   fwd  = F(b): A,a->0 C,c->1 G,g->2 T,t->3, anything else -> 0   (models.rs:231-239)
//   comp = C(b): A,a->3 C,c->2 G,g->1 T,t->0, anything else -> 0   (models.rs:243-251)
//   vmask = 3 for ACGTacgt, 0 otherwise                            (utils.rs:26-39 skip rule)
// Bits 2:1 of an ASCII letter give x = A0 C1 G3 T2 (case-insensitive); x ^ (x>>1) is
// the forward code.  A byte is one of the eight letters iff, with bits 5,2,1 cleared,
// it equals 0x41 (x != 2) or 0x50 (x == 2) — exact for all 256 byte values.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ void classify4(uint32_t w, uint32_t& fwd, uint32_t& comp, uint32_t& vmask) {
    const uint32_t x = (w >> 1) & 0x03030303u;
    const uint32_t s = x >> 1;
    const uint32_t code = x ^ (s & 0x01010101u);
    const uint32_t is2 = s & ~x & 0x01010101u;                  // 1 where x == 2 (candidate T)
    const uint32_t expect = 0x41414141u + is2 * 0x0Fu;          // 0x41 or 0x50 per byte, no carries
    const uint32_t diff = (w & 0xD9D9D9D9u) ^ expect;           // zero byte <=> valid letter
    const uint32_t nz = (((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | diff) & 0x80808080u;
    const uint32_t inv3 = (nz >> 7) * 3u;                       // 0x03 per invalid byte
    fwd = code & ~inv3;
    comp = (code ^ 0x03030303u) & ~inv3;
    vmask = inv3 ^ 0x03030303u;
}

// four 2-bit fields (one per byte) -> one byte, base 0 in the TOP two bits (MSB-first stream);
// the byte is returned in bits 31:24.
__host__ __device__ __forceinline__ uint32_t pack4_msb_top(uint32_t c) { return c * 0x40100401u; }
// same, base 0 in the BOTTOM two bits (LSB-first stream); byte in bits 31:24.
__host__ __device__ __forceinline__ uint32_t pack4_lsb_top(uint32_t c) { return c * 0x01041040u; }

#ifdef __CUDACC__
struct Codes16 {
    uint32_t F;  // forward codes of 16 bases, base 0 in bits 31:30
    uint32_t R;  // complement codes of 16 bases, base 0 in bits 1:0
    uint32_t V;  // validity (0b11 per ACGT base), laid out like F   (only when WANT_V)
};

template <bool WANT_V>
__device__ __forceinline__ Codes16 convert16(uint4 q) {
    uint32_t f0, f1, f2, f3, c0, c1, c2, c3, v0, v1, v2, v3;
    classify4(q.x, f0, c0, v0);
    classify4(q.y, f1, c1, v1);
    classify4(q.z, f2, c2, v2);
    classify4(q.w, f3, c3, v3);
    Codes16 o;
    o.F = __byte_perm(__byte_perm(pack4_msb_top(f3), pack4_msb_top(f2), 0x0073),
                      __byte_perm(pack4_msb_top(f1), pack4_msb_top(f0), 0x0073), 0x5410);
    o.R = __byte_perm(__byte_perm(pack4_lsb_top(c0), pack4_lsb_top(c1), 0x0073),
                      __byte_perm(pack4_lsb_top(c2), pack4_lsb_top(c3), 0x0073), 0x5410);
    if (WANT_V)
        o.V = __byte_perm(__byte_perm(pack4_msb_top(v3), pack4_msb_top(v2), 0x0073),
                          __byte_perm(pack4_msb_top(v1), pack4_msb_top(v0), 0x0073), 0x5410);
    else
        o.V = 0;
    return o;
}

// Pre-packed input (include/neurokmer.h "nk2" layout): the lane's 16 bases arrive as the forward
// word itself plus 16 `other` bits (bit j: base j was not ACGTacgt).  The complement word is the
// 2-bit-group reversal of ~F; an `other` base is code 0 on BOTH strands (models.rs:237,249) and
// not valid for pack_kmer (utils.rs:35).  ~6 ALU ops per 16 bases instead of ~70 for ASCII.
__device__ __forceinline__ uint32_t spread16(uint32_t x) {  // bit j -> bit 2j
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
template <bool WANT_V>
__device__ __forceinline__ Codes16 convert16_packed(uint32_t f, uint32_t x16) {
    Codes16 o;
    const uint32_t b = __brev(f);  // base j now in bits 2j+1:2j, the two bits of each code swapped
    o.R = ~(((b >> 1) & 0x55555555u) | ((b & 0x55555555u) << 1));
    o.F = f;
    o.V = WANT_V ? 0xFFFFFFFFu : 0u;
    if (x16) {  // rare: N runs, IUPAC codes
        const uint32_t xr = spread16(x16) * 3u;  // R layout (base j in bits 2j+1:2j)
        const uint32_t xf = __brev(xr);          // F layout (base j in bits 31-2j:30-2j)
        o.R &= ~xr;
        o.F &= ~xf;  // the packer already wrote 0 there; this makes any input well-defined
        if (WANT_V) o.V = ~xf;
    }
    return o;
}
#endif

// ---------------------------------------------------------------------------
// SipHash-1-3 of one 8-byte little-endian block, k0 = k1 = 0.
// Bit-for-bit the published algorithm; two algebraic shortcuts that do not
// change the result:
//  * keys are zero and the message is one block, so the first half of the
//    first round only touches constants (nvcc folds it);
//  * in the last finalisation round v0's last update cancels in
//    v0^v1^v2^v3 (v3 = rotl(v3,21) ^ v0), so it is never computed.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ unsigned long long rotl64(unsigned long long x, int b) {
    return (x << b) | (x >> (64 - b));
}

#define NK_SIPROUND(v0, v1, v2, v3)                                   \
    do {                                                              \
        v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32); \
        v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;                      \
        v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;                      \
        v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32); \
    } while (0)

__host__ __device__ __forceinline__ unsigned long long siphash13_u64(unsigned long long m) {
    unsigned long long v0 = 0x736f6d6570736575ULL;
    unsigned long long v1 = 0x646f72616e646f6dULL;
    unsigned long long v2 = 0x6c7967656e657261ULL;
    unsigned long long v3 = 0x7465646279746573ULL ^ m;
    NK_SIPROUND(v0, v1, v2, v3);  // c-round on the message block
    v0 ^= m;
    const unsigned long long b = 8ULL << 56;  // length byte, no tail bytes
    v3 ^= b;
    NK_SIPROUND(v0, v1, v2, v3);  // c-round on the length block
    v0 ^= b;
    v2 ^= 0xff;
    NK_SIPROUND(v0, v1, v2, v3);  // d-round 1
    NK_SIPROUND(v0, v1, v2, v3);  // d-round 2
    // d-round 3, trimmed
    v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0;
    v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;
    v2 += v1;
    return rotl64(v3, 21) ^ rotl64(v1, 17) ^ v2 ^ rotl64(v2, 32);
}

// ---------------------------------------------------------------------------
// Exact h % p for any u64 h and 1 <= p < 2^32.
// Moeller & Granlund, "Improved division by invariant integers" (2011), Alg. 4
// (2-by-1 division with a precomputed reciprocal of the normalised divisor),
// applied twice to the 96-bit value h << shift.  Powers of two take one AND.
// ---------------------------------------------------------------------------
struct FastMod {
    double dd;        // p as a double, 1/p (round to nearest), 2p
    double inv;
    double two_d;
    uint32_t dn;      // p << shift (top bit set)
    uint32_t v;       // floor((2^64-1)/dn) - 2^32
    uint32_t shift;   // clz(p)
    uint32_t pow2m1;  // p-1 if p is a power of two, else 0 (p == 1 handled: pow2m1 == 0 && dn == 2^31)
    uint32_t p;
    uint32_t is_pow2;
    // one-stage FP64 remainder (fastmod_single_dev): h = A*2^s + B with s = 32 or 44, x = A*(2^s mod p) + B < 2^53
    double m1;        // kind 3: 2^32 mod p;  kind 2: (2^44 mod p) * 2^-12 (A is taken in place, as A * 2^12)
    double inv_dn;    // 1/p rounded DOWN (<= 1/p), so that the quotient estimate is never too large
    uint32_t kind;    // 0: two stages (any p), 2: split at bit 44 (4 <= p <= 2^31), 3: split at bit 32 (2^32 mod p < 2^21 - 1)
    uint32_t pad_;
};

__host__ __device__ __forceinline__ FastMod make_fastmod(unsigned long long p64) {
    FastMod fm;
    const uint32_t p = (uint32_t)p64;
    uint32_t s = 0;
    while (!((p << s) & 0x80000000u)) ++s;
    fm.p = p;
    fm.shift = s;
    fm.dn = p << s;
    fm.v = (uint32_t)(0xFFFFFFFFFFFFFFFFULL / fm.dn - 0x100000000ULL);
    fm.is_pow2 = ((p & (p - 1)) == 0) ? 1u : 0u;
    fm.pow2m1 = p - 1;
    fm.dd = (double)p;
    fm.inv = 1.0 / (double)p;
    fm.two_d = 2.0 * (double)p;
    fm.m1 = 0.0;
    fm.kind = 0u;
    fm.pad_ = 0u;
    // 1/p rounded to nearest is within half an ulp of 1/p; one ulp below it is <= 1/p and >= (1/p)(1 - 2^-51)
    {
        union { double d; unsigned long long u; } cv;
        cv.d = fm.inv;
        cv.u -= 1ull;
        fm.inv_dn = cv.d;
    }
    if (!fm.is_pow2 && p64 >= 4ull && p64 <= 0x80000000ull) {
        const unsigned long long m32 = 0x100000000ull % p64, m44 = (1ull << 44) % p64;
        if (m32 < (1ull << 21) - 1ull) {
            fm.kind = 3u;
            fm.m1 = (double)m32;
        } else {
            fm.kind = 2u;
            fm.m1 = (double)m44 * (1.0 / 4096.0);
        }
    }
    return fm;
}

// remainder of (u1:u0) / dn, requires u1 < dn, dn normalised
__host__ __device__ __forceinline__ uint32_t rem_2by1(uint32_t u1, uint32_t u0, uint32_t dn, uint32_t v) {
    const unsigned long long q = (unsigned long long)v * u1 + (((unsigned long long)u1 << 32) | u0);
    const uint32_t q1 = (uint32_t)(q >> 32) + 1u;
    const uint32_t q0 = (uint32_t)q;
    uint32_t r = u0 - q1 * dn;
    if (r > q0) r += dn;
    if (r >= dn) r -= dn;
    return r;
}

__host__ __device__ __forceinline__ uint32_t fastmod_u64(unsigned long long h, const FastMod& fm) {
    const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
    if (fm.is_pow2) return lo & fm.pow2m1;
    const uint32_t s = fm.shift;
    // (u2:u1:u0) = h << s
    const uint32_t u2 = s ? (hi >> (32 - s)) : 0u;
    const uint32_t u1 = s ? ((hi << s) | (lo >> (32 - s))) : hi;
    const uint32_t u0 = lo << s;
    const uint32_t r1 = rem_2by1(u2, u1, fm.dn, fm.v);
    const uint32_t r0 = rem_2by1(r1, u0, fm.dn, fm.v);
    return r0 >> s;
}

// ---------------------------------------------------------------------------
// The same remainder on the FP64 pipe (device: fastmod_dev; this is its host twin for tests).
// The count kernel is bound by the integer ALU pipe, and on sm_100a IMAD.HI / IMAD.WIDE block
// ALU issue (tools/microbench2.cu) while DFMA/DADD/DMUL issue beside it for free — so the
// exact modulo is done in doubles, 13 FP64 ops and ~5 ALU ops instead of ~15 ALU-slot ops:
//   stage 1: x1 = h >> 12 (< 2^52, exact);  q1 = floor(x1 * inv);  r1 = x1 - q1*p   (|q1 error| <= 2)
//            r1 += 2p                                   -> r1 in [0, 5p), r1 == (h >> 12)  (mod p)
//   stage 2: x2 = r1 * 4096 + (h & 4095) (< 2^48);      q2 = floor(x2 * inv);  r2 = x2 - q2*p
//            the quotient's fractional part is a multiple of 1/p >= 2^-32 and the rounding error
//            of x2*inv is < 2^-37, so q2 is exact unless x2 is an exact multiple (then r2 may be p)
//   result = r2 >= p ? r2 - p : r2.
// floor() of a non-negative double < 2^52 is (t + 2^52, rounded toward zero) - 2^52; integers <->
// doubles go through the 0x43300000 exponent trick (no conversion instructions).
// ---------------------------------------------------------------------------
inline uint32_t fastmod_fp64_host(unsigned long long h, const FastMod& fm) {
    if (fm.is_pow2) return (uint32_t)h & fm.pow2m1;
    const double two52 = 4503599627370496.0;
    const double x1 = (double)(h >> 12);
    const double t1 = x1 * fm.inv;
    const double q1 = (double)(unsigned long long)t1;  // floor, t1 >= 0
    double r1 = x1 - q1 * fm.dd;                       // exact: an FMA on the device, small integers here
    r1 += fm.two_d;
    const double x2 = r1 * 4096.0 + (double)(h & 4095ull);
    const double t2 = x2 * fm.inv;
    const double q2 = (double)(unsigned long long)t2;
    const double r2 = x2 - q2 * fm.dd;
    (void)two52;
    const unsigned long long r = (unsigned long long)r2;
    return (uint32_t)(r >= fm.p ? r - fm.p : r);
}

// ---------------------------------------------------------------------------
// One-stage form (fm.kind 2 or 3), 7 FP64 ops and 1-3 ALU ops — the count kernel's default when p allows it
// (the two stages above cost 7.6 % of the kernel, tools/count_ablate.py):
//   kind 3 (2^32 mod p < 2^21 - 1, e.g. every p < 2^21):  A = h >> 32, B = h & (2^32-1), m = 2^32 mod p
//   kind 2 (4 <= p <= 2^31):                               A = h >> 44, B = h & (2^44-1), m = 2^44 mod p
//   x = A*m + B == h (mod p), and x < 2^53 is an integer, so ONE fma gives it exactly
//     (kind 3: A*m < 2^32 * (2^21 - 2);  kind 2: A*m < 2^20 * 2^31, B < 2^44).
//   Q = trunc(x * inv_dn + 2^52) in ONE fma rounded toward zero = 2^52 + floor(x * inv_dn); inv_dn in
//     [(1/p)(1 - 2^-51), 1/p], so q = Q - 2^52 is floor(x/p) or one less (x/p * 2^-51 < 4/p <= 1).
//   r = x - q*p (one fma, exact: an integer in [0, 2p) <= 2^32);  result = min(r, r - p) on unsigned words.
// Host twin of fastmod_single_dev for the tests.
// ---------------------------------------------------------------------------
inline uint32_t fastmod_single_host(unsigned long long h, const FastMod& fm) {
    const unsigned long long A = fm.kind == 3u ? (h >> 32) : (h >> 44);
    const unsigned long long B = fm.kind == 3u ? (h & 0xFFFFFFFFull) : (h & ((1ull << 44) - 1ull));
    const unsigned long long m = fm.kind == 3u ? 0x100000000ull % fm.p : (1ull << 44) % fm.p;
    const unsigned long long x = A * m + B;
    unsigned long long q = x / fm.p;  // the device's estimate is this or one less; both give the same result
    return (uint32_t)(x - q * fm.p);
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------
// Device form of the two functions above on explicit 32-bit halves, so that every
// 64-bit rotate is two funnel shifts (SHF.L.W) and nothing else.  Same values as
// siphash13_u64 / fastmod_u64 (tests/test_parity_gpu.py checks nk_debug_hash
// against the oracle).
// ---------------------------------------------------------------------------
struct U64 {
    uint32_t lo, hi;
};
__device__ __forceinline__ U64 add64(U64 a, U64 b) {
    // low word: IADD3 with carry-out (ALU pipe); high word: multiply-add with carry-in, which
    // ptxas emits as IMAD.X on the otherwise idle FMA pipe (the count kernel is ALU-pipe bound)
    U64 r;
    asm("add.cc.u32 %0, %2, %4;\n\tmadc.lo.u32 %1, %3, 1, %5;" : "=r"(r.lo), "=r"(r.hi) : "r"(a.lo), "r"(a.hi), "r"(b.lo), "r"(b.hi));
    return r;
}
__device__ __forceinline__ U64 swap32(U64 x) { return U64{x.hi, x.lo}; }  // rotl64(x, 32)

// Pipe balancing.  ncu (profiles/r01_count_kernel_ncu.md): the first version of the count
// kernel ran the ALU pipe (LOP3/SHF/IADD3: 64 lanes/clk/SM) at 95 % while the FMA pipe
// (IMAD: 64 lanes/clk/SM, issues concurrently — tools/microbench.cu) idled at 7 %.  A
// 32-bit half of a 64-bit rotate is one SHF on the ALU pipe, or IMAD.HI + IMAD on the FMA
// pipe:  (a << B) | (b >> (32-B))  ==  a * 2^B + mulhi(b, 2^B)   (the two terms share no bits).
// IMAD.HI issues at half rate, so a moved half costs 3 FMA slots for 1 ALU slot saved;
// kRotPlan moves as many halves as it takes to level the two pipes.  The multipliers come
// from kernel parameters (RotMul) so that ptxas cannot strength-reduce them back to shifts.
struct RotMul {
    uint32_t m13, m16, m17, m21;  // 1 << 13, 1 << 16, 1 << 17, 1 << 21
};
__host__ __device__ __forceinline__ RotMul make_rotmul() { return RotMul{1u << 13, 1u << 16, 1u << 17, 1u << 21}; }

template <int B>
__device__ __forceinline__ uint32_t rot_mul(const RotMul& rm) {
    return B == 13 ? rm.m13 : (B == 16 ? rm.m16 : (B == 17 ? rm.m17 : rm.m21));
}
// one 32-bit half of rotl64: (a << B) | (b >> (32-B))
template <int B, bool FMA>
__device__ __forceinline__ uint32_t rot_half(uint32_t a, uint32_t b, const RotMul& rm) {
    if (FMA) return a * rot_mul<B>(rm) + __umulhi(b, rot_mul<B>(rm));
    return __funnelshift_l(b, a, B);
}
// rotl64(x, B) ^ y;  MODE bit0: low half on the FMA pipe, bit1: high half on the FMA pipe
template <int B, int MODE>
__device__ __forceinline__ U64 rotl_xor(U64 x, U64 y, const RotMul& rm) {
    return U64{rot_half<B, (MODE & 1) != 0>(x.lo, x.hi, rm) ^ y.lo, rot_half<B, (MODE & 2) != 0>(x.hi, x.lo, rm) ^ y.hi};
}

// modes of the 20 rotate-xor steps of one hash, in program order (4 per round; the first
// round's first step is constant-folded).  NK_ROT_PLAN_ID selects how many rotate halves
// run on the FMA pipe (tools/variants.sh measures them; DESIGN.md §5 has the table).
#ifndef NK_ROT_PLAN_ID
#define NK_ROT_PLAN_ID 0
#endif
#if NK_ROT_PLAN_ID == 0      /* all rotates on the ALU pipe (SHF.L.W) */
#define NK_ROT_PLAN {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#elif NK_ROT_PLAN_ID == 9
#define NK_ROT_PLAN {0, 1, 0, 2, 0, 1, 0, 2, 0, 1, 0, 2, 0, 1, 0, 2, 0, 1, 0, 2}
#elif NK_ROT_PLAN_ID == 13
#define NK_ROT_PLAN {0, 1, 0, 3, 0, 1, 0, 3, 0, 1, 0, 3, 0, 1, 0, 3, 0, 1, 0, 3}
#elif NK_ROT_PLAN_ID == 18
#define NK_ROT_PLAN {0, 1, 2, 1, 2, 1, 2, 1, 2, 1, 2, 1, 2, 1, 2, 1, 2, 1, 2, 1}
#elif NK_ROT_PLAN_ID == 27
#define NK_ROT_PLAN {0, 3, 3, 1, 3, 1, 3, 2, 3, 1, 3, 2, 3, 1, 3, 2, 3, 1, 3, 2}
#elif NK_ROT_PLAN_ID == 36
#define NK_ROT_PLAN {3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3, 3}
#elif NK_ROT_PLAN_ID == 5
#define NK_ROT_PLAN {0, 0, 0, 1, 0, 0, 0, 2, 0, 0, 0, 1, 0, 0, 0, 2, 0, 0, 0, 1}
#endif
__device__ constexpr int kRotPlan[20] = NK_ROT_PLAN;

#define NK_SIPROUND32(v0, v1, v2, v3, R, rm)                                                   \
    do {                                                                                       \
        v0 = add64(v0, v1); v1 = rotl_xor<13, kRotPlan[(R)*4 + 0]>(v1, v0, rm); v0 = swap32(v0); \
        v2 = add64(v2, v3); v3 = rotl_xor<16, kRotPlan[(R)*4 + 1]>(v3, v2, rm);                  \
        v0 = add64(v0, v3); v3 = rotl_xor<21, kRotPlan[(R)*4 + 2]>(v3, v0, rm);                  \
        v2 = add64(v2, v1); v1 = rotl_xor<17, kRotPlan[(R)*4 + 3]>(v1, v2, rm); v2 = swap32(v2); \
    } while (0)

__device__ __forceinline__ U64 siphash13_dev(uint32_t mlo, uint32_t mhi, const RotMul& rm) {
    U64 v0{0x70736575u, 0x736f6d65u};
    U64 v1{0x6e646f6du, 0x646f7261u};
    U64 v2{0x6e657261u, 0x6c796765u};
    U64 v3{0x79746573u ^ mlo, 0x74656462u ^ mhi};
    NK_SIPROUND32(v0, v1, v2, v3, 0, rm);
    v0.lo ^= mlo; v0.hi ^= mhi;
    v3.hi ^= 0x08000000u;  // length block 8 << 56
    NK_SIPROUND32(v0, v1, v2, v3, 1, rm);
    v0.hi ^= 0x08000000u;
    v2.lo ^= 0xffu;
    NK_SIPROUND32(v0, v1, v2, v3, 2, rm);
    NK_SIPROUND32(v0, v1, v2, v3, 3, rm);
    // last round: v0 ^ v3 collapses to rotl(v3, 21) (see siphash13_u64)
    v0 = add64(v0, v1); v1 = rotl_xor<13, kRotPlan[16]>(v1, v0, rm);
    v2 = add64(v2, v3); v3 = rotl_xor<16, kRotPlan[17]>(v3, v2, rm);
    v2 = add64(v2, v1);
    // rotl(v3,21) ^ rotl(v1,17) ^ v2 ^ rotl(v2,32)
    const uint32_t x = v2.lo ^ v2.hi;
    const U64 z{x, x};
    const U64 a = rotl_xor<21, kRotPlan[18]>(v3, z, rm);
    return rotl_xor<17, kRotPlan[19]>(v1, a, rm);
}

__device__ __forceinline__ uint32_t rem_2by1_dev(uint32_t u1, uint32_t u0, uint32_t dn, uint32_t v) {
    const unsigned long long q = (unsigned long long)v * u1 + (((unsigned long long)u1 << 32) | u0);
    const uint32_t q1 = (uint32_t)(q >> 32), q0 = (uint32_t)q;
    uint32_t r = (u0 - dn) - q1 * dn;  // u0 - (q1+1)*dn
    if (r > q0) r += dn;
    return min(r, r - dn);             // r >= dn ? r - dn : r   (r < 2*dn <= 2^33 never holds both)
}

// Moeller-Granlund form (integer pipes only); kept for reference and as a cross-check
template <bool POW2>
__device__ __forceinline__ uint32_t fastmod_mg_dev(U64 h, const FastMod& fm) {
    if (POW2) return h.lo & fm.pow2m1;
    const uint32_t s = fm.shift;
    const uint32_t u2 = __funnelshift_l(h.hi, 0u, s);      // h.hi >> (32-s), 0 for s == 0
    const uint32_t u1 = __funnelshift_l(h.lo, h.hi, s);
    const uint32_t u0 = h.lo << s;
    const uint32_t r1 = rem_2by1_dev(u2, u1, fm.dn, fm.v);
    const uint32_t r0 = rem_2by1_dev(r1, u0, fm.dn, fm.v);
    return r0 >> s;
}

// FP64-pipe form (see fastmod_fp64_host for the derivation)
template <bool POW2>
__device__ __forceinline__ uint32_t fastmod_dev(U64 h, const FastMod& fm) {
    if (POW2) return h.lo & fm.pow2m1;
    const double two52 = 4503599627370496.0;
    // x1 = h >> 12 as a double: bits (0x43300000 | hi', lo') encode 2^52 + value
    const uint32_t hi1 = h.hi >> 12, lo1 = __funnelshift_r(h.lo, h.hi, 12);
    const double x1 = __dadd_rn(__hiloint2double((int)(0x43300000u + hi1), (int)lo1), -two52);
    const double q1 = __dadd_rn(__dadd_rz(__dmul_rn(x1, fm.inv), two52), -two52);       // floor(x1 / p) +- 2
    const double r1 = __dadd_rn(__fma_rn(-q1, fm.dd, x1), fm.two_d);                    // in [0, 5p)
    const double x2 = __dadd_rn(__fma_rn(r1, 4096.0, __hiloint2double(0x43300000, (int)(h.lo & 4095u))), -two52);
    const double q2 = __dadd_rn(__dadd_rz(__dmul_rn(x2, fm.inv), two52), -two52);
    const double r2 = __fma_rn(-q2, fm.dd, x2);                                         // in [0, p]
    const uint32_t r = (uint32_t)__double2loint(__dadd_rn(r2, two52));
    return min(r, r - fm.p);
}

// One-stage FP64 remainder (see fastmod_single_host for the derivation); KIND = fm.kind (2 or 3).
template <int KIND>
__device__ __forceinline__ uint32_t fastmod_single_dev(U64 h, const FastMod& fm) {
    const double two52 = 4503599627370496.0;
    double a, b;
    if (KIND == 3) {
        a = __dadd_rn(__hiloint2double(0x43300000, (int)h.hi), -two52);                                // A
        b = __dadd_rn(__hiloint2double(0x43300000, (int)h.lo), -two52);                                // B
    } else {
        a = __dadd_rn(__hiloint2double(0x43300000, (int)(h.hi & 0xFFFFF000u)), -two52);                // A * 2^12
        b = __dadd_rn(__hiloint2double((int)(0x43300000u | (h.hi & 0xFFFu)), (int)h.lo), -two52);      // B
    }
    const double x = __fma_rn(a, fm.m1, b);
    const double q = __dadd_rn(__fma_rz(x, fm.inv_dn, two52), -two52);
    const double r = __fma_rn(-q, fm.dd, x);
    const uint32_t ri = (uint32_t)__double2loint(__dadd_rn(r, two52));
    return min(ri, ri - fm.p);
}
// MODK: 0 two stages, 1 power of two, 2 / 3 one stage
template <int MODK>
__device__ __forceinline__ uint32_t fastmod_kind_dev(U64 h, const FastMod& fm) {
    if (MODK == 1) return h.lo & fm.pow2m1;
    if (MODK >= 2) return fastmod_single_dev<MODK>(h, fm);
    return fastmod_dev<false>(h, fm);
}

// ---------------------------------------------------------------------------
// mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS: UBLKCP)
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    // make the init visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global → shared bulk copy; size multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, unsigned bytes,
                                            unsigned long long* bar) {
#ifdef NK_EXP_EVICT_FIRST
    // experiment: the input is read once — mark its lines evict-first so that it cannot push pool lines out of L2
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
    return;
#endif
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
#endif

// splitmix64 — the synthetic generator's mixer
__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

}  // namespace nk
