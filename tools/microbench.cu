// tools/microbench.cu — integer-pipe issue rates on sm_100a (B200), measured, not guessed.
// Each kernel runs 8 independent dependency chains per thread of one SASS instruction kind
// (or an interleaved pair) and reports thread-ops / clk / SM from clock64() and CUDA events.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <string>
#include <vector>

#define CH 8
#define UNROLL 8

template <int KIND>
__device__ __forceinline__ void body(unsigned (&a)[CH], unsigned (&b)[CH], unsigned c, unsigned d) {
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        if (KIND == 0) {  // LOP3
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c), "r"(d));
        } else if (KIND == 1) {  // SHF.L.W
            asm volatile("shf.l.wrap.b32 %0, %0, %1, 13;" : "+r"(a[i]) : "r"(b[i]));
        } else if (KIND == 2) {  // IADD3
            asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
        } else if (KIND == 3) {  // IMAD lo
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(d));
        } else if (KIND == 4) {  // IMAD.HI
            asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
        } else if (KIND == 5) {  // IMAD.WIDE
            unsigned long long t;
            asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(a[i]), "r"(c), "l"((unsigned long long)b[i] << 32 | d));
            a[i] = (unsigned)t; b[i] = (unsigned)(t >> 32);
        } else if (KIND == 6) {  // LOP3 + IMAD interleaved (two chains)
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c), "r"(d));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(c), "r"(d));
        } else if (KIND == 7) {  // LOP3 + IMAD.HI
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c), "r"(d));
            asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(c));
        } else if (KIND == 8) {  // LOP3 x2 + IMAD.WIDE
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c), "r"(d));
            unsigned long long t;
            asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(b[i]), "r"(c), "l"((unsigned long long)d));
            b[i] = (unsigned)t ^ (unsigned)(t >> 32);
        } else if (KIND == 9) {  // PRMT
            asm volatile("prmt.b32 %0, %0, %1, 0x1032;" : "+r"(a[i]) : "r"(b[i]));
        } else if (KIND == 10) {  // mul.lo by power of two (IMAD.SHL or SHF?)
            asm volatile("mul.lo.u32 %0, %0, 8192;" : "+r"(a[i]));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
        } else if (KIND == 11) {  // 2 LOP3 : 1 IMAD
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c), "r"(d));
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(d), "r"(c));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(c), "r"(d));
        } else if (KIND == 12) {  // IADD3 + IMAD (add via either pipe?)
            asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(b[i]) : "r"(c), "r"(d));
        } else if (KIND == 13) {  // add.cc / addc pair (64-bit add)
            asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(c), "r"(d));
        } else if (KIND == 14) {  // LOP3 + mad.wide (1:1)
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c), "r"(d));
            unsigned long long t;
            asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(t) : "r"(b[i]), "r"(c), "l"((unsigned long long)d));
            b[i] = (unsigned)(t >> 32);
        } else if (KIND == 15) {  // FFMA + LOP3 (is fma-lite separate from the IMAD path?)
            asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(c), "r"(d));
            float f = __uint_as_float(b[i]);
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f));
            b[i] = __float_as_uint(f);
        } else if (KIND == 16) {  // IMAD + FFMA (do they share a pipe?)
            asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(d));
            float f = __uint_as_float(b[i]);
            asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f));
            b[i] = __float_as_uint(f);
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(256) k(unsigned* out, unsigned iters, unsigned c, unsigned d, long long* clk) {
    unsigned a[CH], b[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = threadIdx.x * 7 + i + c; b[i] = blockIdx.x * 13 + i * 3 + d; }
    long long t0 = clock64();
    for (unsigned it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) body<KIND>(a, b, c, d);
    }
    long long t1 = clock64();
    unsigned x = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) x ^= a[i] ^ b[i];
    if (x == 0x1234567u) out[0] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = t1 - t0;
}

struct Case { const char* name; int ops_per_body; void (*fn)(unsigned*, unsigned, unsigned, unsigned, long long*); };

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned* out; long long* clk; cudaMalloc(&out, 64); cudaMalloc(&clk, 64);
    std::vector<Case> cases = {
        {"LOP3", 1, k<0>}, {"SHF.L.W", 1, k<1>}, {"IADD(add.u32)", 1, k<2>}, {"IMAD.lo", 1, k<3>}, {"IMAD.HI", 1, k<4>},
        {"IMAD.WIDE", 1, k<5>}, {"LOP3+IMAD 1:1", 2, k<6>}, {"LOP3+IMAD.HI 1:1", 2, k<7>}, {"LOP3+WIDE+LOP3(xor halves)", 3, k<8>},
        {"PRMT", 1, k<9>}, {"mul.lo 2^13 + add", 2, k<10>}, {"LOP3x2+IMAD", 3, k<11>}, {"IADD+IMAD 1:1", 2, k<12>},
        {"add.cc+addc", 2, k<13>}, {"LOP3+IMAD.WIDE 1:1", 2, k<14>}, {"LOP3+FFMA 1:1", 2, k<15>}, {"IMAD+FFMA 1:1", 2, k<16>},
    };
    const unsigned iters = 2048;
    const int blocks = sms * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%-32s %12s %12s %10s\n", "case", "Gops/s", "ops/clk/SM", "SM MHz");
    for (auto& cs : cases) {
        float best = 1e30f; long long cyc = 0;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            cs.fn<<<blocks, 256>>>(out, iters, 0x9e3779b9u, 0x85ebca6bu, clk);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) { best = ms; cudaMemcpy(&cyc, clk, 8, cudaMemcpyDeviceToHost); }
        }
        double ops = (double)blocks * 256 * iters * UNROLL * CH * cs.ops_per_body;
        // per-SM rate from the block-0 cycle count: 8 resident blocks/SM run concurrently
        double per_clk_sm = (double)8 * 256 * iters * UNROLL * CH * cs.ops_per_body / (double)cyc;
        printf("%-32s %12.1f %12.2f %10.0f\n", cs.name, ops / (best * 1e-3) / 1e9, per_clk_sm,
               ops / (best * 1e-3) / per_clk_sm / sms / 1e6);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
