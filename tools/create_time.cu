#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
#include "../include/neurokmer.h"
using clk = std::chrono::steady_clock;
static double ms(clk::time_point a) { return std::chrono::duration<double, std::milli>(clk::now() - a).count(); }
int main() {
    auto t0 = clk::now();
    int n = 0; cudaGetDeviceCount(&n); printf("cudaGetDeviceCount (driver init): %.1f ms, %d devices\n", ms(t0), n);
    auto t1 = clk::now(); cudaSetDevice(0); cudaFree(0); printf("cudaSetDevice + context: %.1f ms\n", ms(t1));
    auto t2 = clk::now(); void* p; cudaMalloc(&p, 8 << 20); printf("first cudaMalloc: %.1f ms\n", ms(t2));
    auto t3 = clk::now(); void* q; cudaMallocHost(&q, 64); printf("first cudaMallocHost: %.1f ms\n", ms(t3));
    auto t4 = clk::now(); nk_config cfg; nk_config_default(&cfg); cfg.pool_size = 2000000; cfg.use_canonical = 1; nk_counter* h = nullptr;
    int rc = nk_create(&cfg, &h); printf("nk_create (after context exists): %.1f ms rc=%d\n", ms(t4), rc);
    auto t5 = clk::now(); nk_counter* h2 = nullptr; nk_create(&cfg, &h2); printf("second nk_create: %.1f ms\n", ms(t5));
    nk_destroy(h); nk_destroy(h2);
    return 0;
}
