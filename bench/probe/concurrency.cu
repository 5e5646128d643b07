// concurrency.cu -- how many grids does the device run at once?  (bench/probe: measurement only)
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
__global__ void spin(long long cycles, int *sink) {
    extern __shared__ double sm[];
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { }
    if (sink && threadIdx.x == 0 && cycles < 0) sink[0] = (int)sm[0];
}
__global__ void tiny(int *q) { if (threadIdx.x == 0) q[0] = 0; }
static int g_var = 0;
static double g_sum = 0;
static double run(int S, int K, int ctas, int threads, size_t smem, int pre, long long cyc, int *d) {
    std::vector<cudaStream_t> st(S);
    for (auto &s : st) cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaFuncSetAttribute(spin, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) {
        cudaDeviceSynchronize();
        cudaEventRecord(e0, 0);
        for (auto &s : st) cudaStreamWaitEvent(s, e0, 0);
        for (int k = 0; k < K; k++) {
            cudaStream_t s = st[k % S];
            if (pre == 1) cudaMemsetAsync(d + (k % 1024) * 32, 0, 16, s);
            if (pre == 2) tiny<<<1, 1024, 0, s>>>(d + (k % 1024) * 32);
            long long c2 = cyc;
            if (g_var) { unsigned h = (unsigned)k * 2654435761u; h ^= h >> 15; const unsigned r = h % 100; c2 = r < 3 ? cyc * 6 : (cyc / 2 + (cyc * (h % 97)) / 97); }
            if (w == 1) g_sum += (double)c2 * ctas;
            spin<<<ctas, threads, smem, s>>>(c2, nullptr);
        }
        cudaEvent_t ev; cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        for (auto &s : st) { cudaEventRecord(ev, s); cudaStreamWaitEvent(0, ev, 0); }
        cudaEventRecord(e1, 0);
        cudaDeviceSynchronize();
    }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    for (auto &s : st) cudaStreamDestroy(s);
    return ms;
}
int main() {
    int *d; cudaMalloc(&d, 1024 * 128); 
    const long long cyc = 2000000;   // ~1 ms
    const double one = cyc / 1.965e6;
    struct { int S, K, ctas, threads; size_t smem; int pre; } cs[] = {
        {128, 1024, 1, 32, 0, 0}, {256, 1024, 1, 32, 0, 0}, {128, 1024, 1, 352, 219 * 1024, 0}, {128, 1024, 1, 352, 219 * 1024, 1},
        {128, 1024, 1, 352, 219 * 1024, 2}, {128, 1024, 2, 352, 219 * 1024, 0}, {128, 1024, 4, 352, 219 * 1024, 0}, {128, 1024, 4, 352, 219 * 1024, 1},
        {128, 1024, 4, 352, 100 * 1024, 0}, {128, 1024, 1, 352, 100 * 1024, 0}};
    for (auto &c : cs) {
        g_var = 0;
        const double ms = run(c.S, c.K, c.ctas, c.threads, c.smem, c.pre, cyc, d);
        printf("streams %3d kernels %d ctas %d threads %d smem %3zu KB pre %d: %.2f ms -> %.1f grids (%.1f CTAs) at once\n", c.S, c.K, c.ctas, c.threads,
               c.smem / 1024, c.pre, ms, c.K * one / ms, c.K * one / ms * c.ctas);
    }
    for (int S : {32, 128, 256, 512})
      for (int ctas : {2, 4}) {
        g_var = 1; g_sum = 0;
        const double ms = run(S, 2048, ctas, 352, 219 * 1024, 1, cyc, d);
        printf("variable durations, streams %d ctas %d: %.2f ms, SM occupancy %.1f of 148\n", S, ctas, ms, g_sum / 1.965e6 / ms);
      }
    return 0;
}
