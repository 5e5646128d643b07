// named_barrier.cu -- does bar.arrive / bar.sync hand-over order the two sides as expected?  (measurement only)
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void nbar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void nbar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void spin(long long c) { const long long t0 = clock64(); while (clock64() - t0 < c) { } }
__global__ void k(long long *out, int cycles) {
    __shared__ volatile int box1, box2;
    int bad1 = 0, bad2 = 0;
    const int n = blockDim.x, tid = threadIdx.x;
    if (tid == 0) { box1 = 0; box2 = 0; }
    const bool ctl = tid >= n - 32;
    for (int c = 0; c < cycles; c++) {
        __syncthreads();
        if (ctl) {
            nbar_sync(1, n);                       // wait for the producers' first half
            if (box1 != c + 1) bad1++;
            if (tid == n - 32) out[c * 8 + 0] = clock64();
            spin(3000);
            __syncthreads();
            spin(5000);
            if (tid == n - 32) out[c * 8 + 1] = clock64();
            if (tid == n - 32) box2 = c + 1;
            __syncwarp();
            nbar_arrive(2, n);
            spin(2000);
            __syncthreads();
        } else {
            spin(1000);
            if (tid == 0) out[c * 8 + 2] = clock64();
            if (tid == 0) box1 = c + 1;
            __syncwarp();
            nbar_arrive(1, n);
            spin(1000);
            __syncthreads();
            nbar_sync(2, n);
            if (box2 != c + 1) bad2++;
            if (tid == 0) out[c * 8 + 3] = clock64();
            spin(500);
            __syncthreads();
        }
    }
    if (tid == n - 32) out[60] = bad1;
    if (tid == 0) out[61] = bad2;
}
int main() {
    long long *d, h[8 * 8];
    cudaMalloc(&d, sizeof(h)); cudaMemset(d, 0, sizeof(h));
    k<<<1, 352>>>(d, 8);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("status %s\n", cudaGetErrorString(e));
    for (int c = 0; c < 8; c++)
        printf("cycle %d: producers arrive(1) at %lld, consumer passed sync(1) at +%lld; control arrive(2) at +%lld, stage passed sync(2) at +%lld\n",
               c, h[c * 8 + 2] - h[2], h[c * 8 + 0] - h[c * 8 + 2], h[c * 8 + 1] - h[c * 8 + 2], h[c * 8 + 3] - h[c * 8 + 2]);
    printf("hand-over failures: consumer of barrier 1: %lld of 8, consumers of barrier 2: %lld of 8\n", h[60], h[61]);
    return 0;
}
