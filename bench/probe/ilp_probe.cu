// bench/probe/ilp_probe.cu -- how much does instruction-level parallelism inside a stage thread buy?
// 10 warps per CTA (the stage warps of nmpc_solve_kernel), one CTA per SM.  Every thread evaluates, per round, the
// transcendental core of two stage evaluations (2 x sincos(theta), sincos(etheta), log(slack product)) either one stage
// after the other (chained through a register dependency, as the kernel's two stage_eval calls are by their branches) or
// with the two stages' chains free to interleave.  Prints SM cycles per round.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mpc_ros_b200/csrc -o bench/probe/ilp_probe bench/probe/ilp_probe.cu
#include "nmpc_phases.cuh"
#include <cstdio>
using namespace nmpc;

template <int MODE>
__global__ void __launch_bounds__(384, 1) probe(double *out, int rounds, long long *cyc)
{
    const int t = threadIdx.x;
    double th0 = 0.1 + 1e-3 * t, e0 = -0.2 + 1e-3 * t, s0 = 1.5 + 1e-3 * t;
    double th1 = 0.3 + 1e-3 * t, e1 = 0.25 + 1e-3 * t, s1 = 2.5 + 1e-3 * t;
    double acc = 0.0;
    __syncthreads();
    const long long c0 = clock64();
    if (t >= 64) {
        for (int r = 0; r < rounds; r++) {
            double a, b, c, d, l, a2, b2, c2, d2, l2;
            if (MODE == 0) {
                sincos_d(th0, &a, &b); sincos_d(e0, &c, &d); l = log_pos(s0);
                double dep = (a + b) + (c + d) + l;
                // the second stage starts only after the first has finished
                th1 += 1e-300 * dep; e1 += 1e-300 * dep; s1 += 1e-300 * dep;
                sincos_d(th1, &a2, &b2); sincos_d(e1, &c2, &d2); l2 = log_pos(s1);
                acc += dep + (a2 + b2) + (c2 + d2) + l2;
            } else if (MODE == 1) {
                sincos_d(th0, &a, &b); sincos_d(e0, &c, &d); l = log_pos(s0);
                sincos_d(th1, &a2, &b2); sincos_d(e1, &c2, &d2); l2 = log_pos(s1);
                acc += (a + b) + (c + d) + l + (a2 + b2) + (c2 + d2) + l2;
            } else {
                // one logarithm for both stages
                sincos_d(th0, &a, &b); sincos_d(e0, &c, &d);
                sincos_d(th1, &a2, &b2); sincos_d(e1, &c2, &d2); l = log_pos(s0 * s1);
                acc += (a + b) + (c + d) + l + (a2 + b2) + (c2 + d2);
            }
            th0 += 1e-3 * acc * 1e-3; e0 -= 1e-6 * acc; s0 += 1e-9 * acc;
            th1 += 1e-6 * acc; e1 -= 1e-6 * acc; s1 += 1e-9 * acc;
        }
    }
    __syncthreads();
    const long long c1 = clock64();
    out[blockIdx.x * blockDim.x + t] = acc;
    if (t == 0 && blockIdx.x == 0) *cyc = c1 - c0;
}


template <int MODE> static void run(double *out, long long *cyc, int R, const char *name)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<148, 384>>>(out, R, cyc);
    cudaEventRecord(e0); probe<MODE><<<148, 384>>>(out, R, cyc); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h = -1; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-24s: %.1f cycles per round by events at 1.965 GHz, %.1f by clock64  (%s)\n", name, ms * 1e-3 * 1.965e9 / R, (double)h / R,
           cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    double *out; long long *cyc;
    cudaMalloc(&out, 148 * 384 * sizeof(double)); cudaMalloc(&cyc, 8);
    const int R = 20000;
    for (int rep = 0; rep < 2; rep++) {
        run<0>(out, cyc, R, "sequential stages");
        run<1>(out, cyc, R, "interleaved stages");
        run<2>(out, cyc, R, "interleaved, one log");
    }
    return 0;
}
