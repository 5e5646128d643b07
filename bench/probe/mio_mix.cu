// mio_mix.cu -- how long does one "Riccati stage" worth of instructions take when W warps of an SM issue it at once?
// (measurement only)  mix A = scalar 4x4 sweep: 140 DFMA, 35 LDS.64, 10 STS.64 per stage, 16 lanes per warp;
// mix B = 4-lane column split: 70 DFMA, 36 SHFL.32, 28 SEL, 25 LDS.64 (4 lanes read one address), 4 STS.64, 32 lanes per warp.
#include <cstdio>
#include <cuda_runtime.h>
template <int NF, int NL, int NS, int NSH, int NSEL>
__device__ __forceinline__ void body(volatile double *sm, int ld, int lane, double &a0, double &a1, double &a2, double &a3, double &acc,
                                     double &b0, double &b1, double &b2, double &b3)
{
    double l[NL > 0 ? NL : 1];
#pragma unroll
    for (int i = 0; i < NL; i++) l[i] = sm[i * 32 + ld];
#pragma unroll
    for (int i = 0; i < NF / 8; i++) {      // 8 independent chains: throughput-bound
        const double c = l[i % (NL > 0 ? NL : 1)];
        a0 = fma(a0, c, c); a1 = fma(a1, c, c); a2 = fma(a2, c, c); a3 = fma(a3, c, c);
        b0 = fma(b0, c, c); b1 = fma(b1, c, c); b2 = fma(b2, c, c); b3 = fma(b3, c, c);
    }
#pragma unroll
    for (int i = 0; i < NSH / 2; i++) { acc += __shfl_xor_sync(0xffffffffu, (i & 1) ? a0 : a1, 1 + (i % 3)); }
#pragma unroll
    for (int i = 0; i < NSEL / 2; i++) { a2 = ((lane + i) & 2) ? a2 : a3; }
#pragma unroll
    for (int i = 0; i < NS; i++) sm[(40 + i) * 32 + lane] = a0 + i;
}
template <int NF, int NL, int NS, int NSH, int NSEL, int LANES>
__global__ void k(long long *out, int iters, int warps, double seed)
{
    __shared__ double smd[64 * 32];
    volatile double *sm = smd;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int i = tid; i < 64 * 32; i += blockDim.x) smd[i] = 1.0 + 1e-9 * i;
    __syncthreads();
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, acc = 0, b0 = seed, b1 = seed, b2 = seed, b3 = seed;
    long long t0 = 0, t1 = 0;
    if (w < warps) {
        const bool act = lane < LANES;
        const int ld = LANES == 16 ? lane : (lane >> 2);
        t0 = clock64();
        if (act)
            for (int it = 0; it < iters; it++) body<NF, NL, NS, NSH, NSEL>(sm, ld, lane, a0, a1, a2, a3, acc, b0, b1, b2, b3);
        t1 = clock64();
    }
    if (lane == 0 && w < warps && blockIdx.x == 0) out[w] = t1 - t0;
    if (a0 + a1 + a2 + a3 + acc + b0 + b1 + b2 + b3 == 12345.678) out[63] = 1;
}
template <int NF, int NL, int NS, int NSH, int NSEL, int LANES>
void run(const char *name, long long *d)
{
    long long h[64];
    const int iters = 2000;
    for (int warps = 1; warps <= 4; warps *= 2) {
        cudaMemset(d, 0, 64 * 8);
        k<NF, NL, NS, NSH, NSEL, LANES><<<148, 384>>>(d, iters, warps, 1e-3);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 64 * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < warps; i++) mx = h[i] > mx ? h[i] : mx;
        printf("%-44s warps %d: %7.1f cycles per stage\n", name, warps, (double)mx / iters);
    }
}
int main()
{
    long long *d;
    cudaMalloc(&d, 64 * 8);
    run<144, 35, 10, 0, 0, 16>("A scalar: 144 DFMA 35 LDS 10 STS (16 lanes)", d);
    run<144, 0, 0, 0, 0, 16>("144 DFMA only", d);
    run<0, 35, 0, 0, 0, 16>("35 LDS only (16 lanes)", d);
    run<0, 35, 0, 0, 0, 32>("35 LDS only (32 lanes, 4 per address)", d);
    run<0, 0, 10, 0, 0, 32>("10 STS only", d);
    run<0, 0, 0, 36, 0, 32>("36 SHFL.32 (18 double) only", d);
    run<0, 0, 0, 0, 28, 32>("28 SEL only", d);
    run<72, 25, 4, 36, 28, 32>("B split: 72 DFMA 25 LDS 4 STS 36 SHFL 28 SEL", d);
    run<72, 25, 4, 0, 28, 32>("B without SHFL", d);
    run<72, 45, 12, 6, 16, 32>("C smem exchange: 72 DFMA 45 LDS 12 STS 6 SHFL 16 SEL", d);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
