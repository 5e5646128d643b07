// bench/issue_loop.cpp -- the issue loop of bench.py's device-resident leg in C++ (the host side of this path is C++:
// north-star; a Python loop costs ~100 us per batch, which is what a batch of 4,096 takes on the device).
// Issues batches round-robin on S streams through the C ABI (mpc_b200_prestep_batch + mpc_b200_solve_batch, device
// pointers), keeps at most `depth` batches in flight per stream (waits on the event of the batch `depth` back) and
// brackets every `sample_every`-th solve launch with timing events on its own stream.  bench.py records its own CUDA
// events around the call and owns all buffers.  Built by __graft_entry__.build() into bench/libmpc_issue.so.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include "../include/mpc_b200.h"

struct IssueCtx {
    int S, depth;
    std::vector<cudaEvent_t> done;               // S x depth
    std::vector<long long> issued;               // per stream
    std::vector<cudaEvent_t> sx, sy;             // sampled launches
    size_t nsample;
};

extern "C" IssueCtx *mpcb_issue_create(int device, int S, int depth, int max_samples)
{
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    IssueCtx *c = new IssueCtx;
    c->S = S; c->depth = depth; c->nsample = 0;
    c->done.resize((size_t)S * depth); c->issued.assign((size_t)S, 0);
    for (auto &e : c->done) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    c->sx.resize((size_t)max_samples); c->sy.resize((size_t)max_samples);
    for (int i = 0; i < max_samples; i++) { cudaEventCreate(&c->sx[i]); cudaEventCreate(&c->sy[i]); }
    return c;
}

extern "C" void mpcb_issue_destroy(IssueCtx *c)
{
    if (!c) return;
    for (auto &e : c->done) cudaEventDestroy(e);
    for (auto &e : c->sx) cudaEventDestroy(e);
    for (auto &e : c->sy) cudaEventDestroy(e);
    delete c;
}

// Batches j0 .. j0 + n - 1; batch j runs on stream j % S with buffer set j % R.  ptr[q] = the 14 device pointers of set q:
// wx, wy, pose, vel, coeffs, state, u0, pred, obj, status, iters, kkt (12 used).
extern "C" int mpcb_issue_run(IssueCtx *c, mpc_b200_handle *h, long long j0, int n, void *const *streams, int R,
                              const uintptr_t *ptr, int B, int M, int sample_every)
{
    for (long long j = j0; j < j0 + n; j++) {
        const int s = (int)(j % c->S);
        const long long k = c->issued[s];
        cudaEvent_t ev = c->done[(size_t)s * c->depth + (size_t)(k % c->depth)];
        if (k >= c->depth && cudaEventSynchronize(ev) != cudaSuccess) return -100;
        const uintptr_t *p = ptr + (size_t)(j % R) * 12;
        cudaStream_t st = (cudaStream_t)streams[s];
        int rc = mpc_b200_prestep_batch(h, B, M, (const double *)p[0], (const double *)p[1], (const double *)p[2], (const double *)p[3],
                                        (double *)p[4], (double *)p[5], streams[s]);
        if (rc != MPC_B200_OK) return rc;
        const bool sample = sample_every > 0 && ((j - j0) % sample_every) == 0 && c->nsample < c->sx.size();
        if (sample) cudaEventRecord(c->sx[c->nsample], st);
        rc = mpc_b200_solve_batch(h, B, (const double *)p[5], (const double *)p[4], nullptr, nullptr, (double *)p[6], (double *)p[7],
                                  (double *)p[8], (int32_t *)p[9], (int32_t *)p[10], (double *)p[11], nullptr, streams[s]);
        if (rc != MPC_B200_OK) return rc;
        if (sample) { cudaEventRecord(c->sy[c->nsample], st); c->nsample++; }
        if (cudaEventRecord(ev, st) != cudaSuccess) return -101;
        c->issued[s] = k + 1;
    }
    return 0;
}

// Durations (ms) of the sampled solve launches; call after the device has been synchronised.  Returns the count; resets.
extern "C" int mpcb_issue_samples(IssueCtx *c, float *ms, int cap)
{
    int n = 0;
    for (size_t i = 0; i < c->nsample && n < cap; i++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, c->sx[i], c->sy[i]) == cudaSuccess) ms[n++] = t;
    }
    c->nsample = 0;
    return n;
}
