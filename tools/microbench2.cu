// tools/microbench2.cu — does register-operand bandwidth (not the pipe) bound mixed ALU+FMA issue?
// Chains use DISTINCT registers per instruction (no operand reuse across instructions).
#include <cstdio>
#include <cuda_runtime.h>
#define CH 8
template <int KIND>
__global__ void __launch_bounds__(256) k(unsigned* out, unsigned iters, unsigned c, unsigned d) {
    unsigned a[CH], b[CH], e[CH], f[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = threadIdx.x * 7 + i + c; b[i] = blockIdx.x * 13 + i * 3 + d; e[i] = a[i] * 5 + 1; f[i] = b[i] * 3 + 7 + threadIdx.x * 11; }
    for (unsigned it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (KIND == 0) {          // LOP3, 3 distinct regs
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(e[i]));
                } else if (KIND == 1) {   // LOP3 (3 regs) + IMAD (3 regs), distinct
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(e[i]));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(f[i]) : "r"(b[(i + 1) % CH]), "r"(e[(i + 1) % CH]));
                } else if (KIND == 2) {   // LOP3 2 regs + imm
                    asm volatile("lop3.b32 %0, %0, %1, 0x12345, 0x96;" : "+r"(a[i]) : "r"(b[i]));
                } else if (KIND == 3) {   // LOP3 (2 regs+imm) + IMAD (2 regs + imm)
                    asm volatile("lop3.b32 %0, %0, %1, 0x12345, 0x96;" : "+r"(a[i]) : "r"(b[i]));
                    asm volatile("mad.lo.u32 %0, %0, 8193, %1;" : "+r"(f[i]) : "r"(e[i]));
                } else if (KIND == 4) {   // SHF funnel 2 regs + LOP3 2 regs + IADD 2 regs (ALU only, like SipHash)
                    asm volatile("shf.l.wrap.b32 %0, %0, %1, 13;" : "+r"(a[i]) : "r"(b[i]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x3c;" : "+r"(b[i]) : "r"(a[i]), "r"(a[i]));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(e[i]) : "r"(b[i]));
                } else if (KIND == 5) {   // 3 LOP3 : 1 IMAD.HI distinct
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(e[i]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[i]) : "r"(a[i]), "r"(e[i]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(e[i]) : "r"(b[i]), "r"(a[i]));
                    asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(f[i]) : "r"(b[(i + 1) % CH]));
                } else if (KIND == 7 || KIND == 8) {   // N LOP3 : 1 IMAD.WIDE (does the wide multiply block ALU issue?)
                    constexpr int NL = KIND == 7 ? 4 : 8;
#pragma unroll
                    for (int l = 0; l < NL; ++l)
                        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[(i + l) % CH]), "r"(e[(i + l + 1) % CH]));
                    unsigned long long t;
                    asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(f[i]), "r"(c));
                    f[i] = (unsigned)t ^ (unsigned)(t >> 32);
                } else if (KIND == 9 || KIND == 10) {  // N LOP3 : 1 DFMA (is the FP64 pipe independent of ALU issue?)
                    constexpr int NL = KIND == 9 ? 4 : 1;
#pragma unroll
                    for (int l = 0; l < NL; ++l)
                        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[(i + l) % CH]), "r"(e[(i + l + 1) % CH]));
                    double dd = __hiloint2double(0x43300000 | (f[i] & 0xFFFF), f[i]);
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd) : "d"(1.0000001), "d"(-4503599627370496.0));
                    f[i] = __double2loint(dd) ^ __double2hiint(dd);
                } else if (KIND == 11) {  // DFMA chain alone
                    double dd = __hiloint2double(0x43300000 | (f[i] & 0xFFFF), f[i]);
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd) : "d"(1.0000001), "d"(-4503599627370496.0));
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd) : "d"(1.0000001), "d"(0.5));
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd) : "d"(0.999), "d"(0.25));
                    asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd) : "d"(1.001), "d"(0.125));
                    f[i] = __double2loint(dd);
                } else if (KIND == 6) {   // 2 ALU : 1 IMAD : distinct (target mix)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(e[i]));
                    asm volatile("shf.l.wrap.b32 %0, %0, %1, 13;" : "+r"(b[i]) : "r"(a[i]));
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(f[i]) : "r"(b[(i + 1) % CH]), "r"(e[(i + 1) % CH]));
                }
            }
        }
    }
    unsigned x = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) x ^= a[i] ^ b[i] ^ e[i] ^ f[i];
    if (x == 0x1234567u) out[0] = x;
}
struct Case { const char* name; int ops; void (*fn)(unsigned*, unsigned, unsigned, unsigned); };
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned* out; cudaMalloc(&out, 64);
    Case cases[] = {{"LOP3 3 distinct regs", 1, k<0>}, {"LOP3+IMAD distinct 3-reg", 2, k<1>}, {"LOP3 2reg+imm", 1, k<2>},
                    {"LOP3+IMAD 2reg+imm", 2, k<3>}, {"SHF+LOP3+IADD (alu only)", 3, k<4>}, {"3 LOP3 : 1 IMAD.HI", 4, k<5>},
                    {"LOP3+SHF+IMAD distinct", 3, k<6>},
                    {"4 LOP3 : 1 IMAD.WIDE (+1 LOP3 merge)", 6, k<7>}, {"8 LOP3 : 1 IMAD.WIDE (+1 LOP3 merge)", 10, k<8>},
                    {"4 LOP3 : 1 DFMA (+2 glue ALU)", 7, k<9>}, {"1 LOP3 : 1 DFMA (+2 glue ALU)", 4, k<10>}, {"4 DFMA chain (+1 glue)", 5, k<11>}};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned iters = 2048; const int blocks = sms * 8;
    for (auto& cs : cases) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0); cs.fn<<<blocks, 256>>>(out, iters, 0x9e3779b9u, 0x85ebca6bu); cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
        }
        double ops = (double)blocks * 256 * iters * 8 * CH * cs.ops;
        printf("%-30s %10.1f Gops/s  %7.2f ops/clk/SM @1.95GHz\n", cs.name, ops / (best * 1e-3) / 1e9, ops / (best * 1e-3) / (sms * 1.95e9));
    }
    return cudaGetLastError() != cudaSuccess;
}
