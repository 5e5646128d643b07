// nmpc_kernel.cuh -- the batched NMPC solve kernel for sm_100a.
//
// Persistent CTAs over a global work queue.  A CTA has PB <= 32 problem lanes.  Warp 0 is the control
// warp: lane p runs the interior-point logic and the serial Riccati sweeps of the problem currently in
// lane p -- one thread per problem, 32 different problems per warp instruction (no idle lanes, no
// shuffles in the recursions).  The other warps are stage threads: thread (g, p) owns stages
// g*SPT .. g*SPT+SPT-1 of lane p's problem and does everything that is parallel over the horizon
// (sin/cos, model derivatives, residual norms, line-search trial evaluation, step application).
// All exchange goes through shared memory [stage][slot][problem]; phases are separated by CTA
// barriers and both roles execute the same barrier sequence (B0..B9, V1, V2, R1, R2 below).
// When a lane's problem terminates its results are written out and the lane pops the next problem
// from the queue, so lanes never wait for the slowest problem of a batch (iteration counts range
// from 5 to the cap).  Lanes are independent: each is at its own interior-point iteration.
//
// Replaces, per problem: CppAD::ipopt::solve + Ipopt (mpc_ros/src/mpc_planner.cpp:373-375).
#pragma once
#include "nmpc_phases.cuh"

namespace nmpc {

struct SolveArgs {
    Params prm;
    int PB;
    int batch;
    const double *state;    // 6 x batch
    const double *coeffs;   // 4 x batch
    const double *ref_vel;  // batch or NULL
    double *u0;             // 2 x batch
    double *pred;           // 3N x batch
    double *obj;            // batch or NULL
    int *status;            // batch or NULL
    int *iters;             // batch or NULL
    double *kkt;            // batch or NULL
    double *warm_out;       // warm_size x batch or NULL
    int *queue;             // work-queue head (zeroed by the host before the launch)
    long long *prof;        // NMPC_PROFILE builds only: per-phase cycle counters of CTA 0
};

#ifdef NMPC_PROFILE
#define PROF_DECL long long prof_t = clock64(), prof_acc[12] = {0,0,0,0,0,0,0,0,0,0,0,0}
#define PROF_MARK(i) do { long long t_ = clock64(); prof_acc[i] += t_ - prof_t; prof_t = t_; } while (0)
#define PROF_FLUSH() do { if (a.prof && blockIdx.x == 0 && tid == 0) for (int q_ = 0; q_ < 12; q_++) a.prof[q_] = prof_acc[q_]; } while (0)
#else
#define PROF_DECL
#define PROF_MARK(i)
#define PROF_FLUSH()
#endif

template <int SPT, int CPB>
__global__ void __launch_bounds__(256, 1) nmpc_solve_kernel(const SolveArgs a)
{
    extern __shared__ double smem_raw[];
    const Params &prm = a.prm;
    const int N = prm.N, PB = (CPB > 0) ? CPB : a.PB, batch = a.batch;
    SmemT<CPB> sm;
    sm.st = smem_raw;
    sm.ps = smem_raw + (size_t)N * NSLOTS * PB;
    sm.pi = reinterpret_cast<int *>(sm.ps + (size_t)NPS * PB);
    sm.PB = PB;

    const int tid = threadIdx.x;

    if (tid < 32) {
        // ------------------------------------------------------------ control warp
        const int p = tid;
        const bool lane = p < PB;
        Ctrl c;
        c.status = 0; c.iter = 0; c.E0 = 0.0; c.obj = 0.0;
        // initial problems: one queue pop for the whole CTA
        int base = 0;
        if (p == 0) base = atomicAdd(a.queue, PB);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane) {
            const int idx = base + p;
            sm.I(PI_MODE, p) = MODE_IDLE; sm.I(PI_STATUS, p) = 0; sm.I(PI_LSQ, p) = 0;
            sm.I(PI_PROB, p) = idx; sm.I(PI_NEXT, p) = (idx < batch) ? idx : -1;
        }
        __syncthreads();  // R1 (initial)
        if (lane && sm.I(PI_NEXT, p) >= 0) {
            const int idx = sm.I(PI_NEXT, p);
            double s6[6];
            for (int i = 0; i < 6; i++) s6[i] = a.state[(size_t)i * batch + idx];
            const double rv = a.ref_vel ? a.ref_vel[idx] : prm.ref_vel;
            ctrl_init(prm, sm, c, p, s6, rv);
        }
        __syncthreads();  // R2 (initial) == B0
        PROF_DECL;
        for (;;) {
            __syncthreads();  // B1: residual partials written
            PROF_MARK(0);
            int done = 0;
            if (lane) {
                const int md = sm.I(PI_MODE, p);
                if (md == MODE_RESID || md == MODE_ACCEPT) {
                    if (sm.I(PI_LSQ, p) || ctrl_check(prm, sm, c, p)) { sm.I(PI_MODE, p) = MODE_COEF; sm.P(PS_DW, p) = 0.0; }
                    else { sm.I(PI_MODE, p) = MODE_DONE; sm.I(PI_STATUS, p) = c.status; done = 1; }
                } else if (md == MODE_DONE) {
                    done = 1;   // failed in the previous cycle's sweep / line search
                }
            }
            PROF_MARK(1);
            if (__syncthreads_or(done)) {  // V1: some lane finished -> flush results, refill
                // (a lane can also be DONE because a sweep / line search failed in the previous cycle)
                const bool fin = lane && sm.I(PI_MODE, p) == MODE_DONE;
                if (fin) {
                    const size_t i = (size_t)sm.I(PI_PROB, p);
                    if (a.obj) a.obj[i] = c.obj;
                    if (a.status) a.status[i] = c.status;
                    if (a.iters) a.iters[i] = c.iter;
                    if (a.kkt) a.kkt[i] = c.E0;
                }
                const unsigned m = __ballot_sync(0xffffffffu, fin);
                int nb = 0;
                if (p == 0 && m) nb = atomicAdd(a.queue, __popc(m));
                nb = __shfl_sync(0xffffffffu, nb, 0);
                int nidx = -1;
                if (fin) { nidx = nb + __popc(m & ((1u << p) - 1u)); if (nidx >= batch) nidx = -1; }
                if (lane) sm.I(PI_NEXT, p) = nidx;
                __syncthreads();  // R1: results flushed by the stage threads, next problems known
                if (fin) {
                    if (nidx >= 0) {
                        sm.I(PI_PROB, p) = nidx;
                        double s6[6];
                        for (int i = 0; i < 6; i++) s6[i] = a.state[(size_t)i * batch + nidx];
                        const double rv = a.ref_vel ? a.ref_vel[nidx] : prm.ref_vel;
                        ctrl_init(prm, sm, c, p, s6, rv);
                    } else {
                        sm.I(PI_MODE, p) = MODE_IDLE;
                    }
                }
                __syncthreads();  // R2: new problems initialised
            }
            int running = 0;
            if (lane) running = sm.I(PI_MODE, p) != MODE_IDLE;
            if (!__syncthreads_or(running)) break;  // V2
            for (;;) {
                __syncthreads();  // B3: coefficients written
                PROF_MARK(2);
                int retry = 0;
                if (lane && sm.I(PI_MODE, p) == MODE_COEF) {
                    const int lsq = sm.I(PI_LSQ, p);
                    const double dw = sm.P(PS_DW, p);
                    const HessDiag hd = hess_diag(prm, sm.P(PS_SF, p), dw, lsq);
                    const int okb = riccati_backward(prm, sm, p, hd);
                    PROF_MARK(3);
                    if (okb || lsq) {
                        riccati_forward(prm, sm, p);
                        if (dw > 0.0) c.dw_last = dw;
                        sm.I(PI_MODE, p) = MODE_STEP;
                    } else {
                        const double nd = next_dw(c, dw);
                        if (nd > NMPC_DW_MAX) { c.status = 10; sm.I(PI_STATUS, p) = 10; sm.I(PI_MODE, p) = MODE_DONE; }
                        else { sm.P(PS_DW, p) = nd; retry = 1; }
                    }
                }
                PROF_MARK(4);
                if (!__syncthreads_or(retry)) break;  // B4
            }
            __syncthreads();  // B5: step partials written
            PROF_MARK(5);
            if (lane && sm.I(PI_MODE, p) == MODE_STEP) {
                if (sm.I(PI_LSQ, p)) { ctrl_lsq_finish(prm, sm, c, p); sm.I(PI_MODE, p) = MODE_ACCEPT; }
                else { ctrl_step(prm, sm, c, p); sm.I(PI_MODE, p) = MODE_TRIAL; }
            }
            PROF_MARK(6);
            __syncthreads();  // B6
            for (;;) {
                __syncthreads();  // B7: trial partials written
                PROF_MARK(7);
                int again = 0;
                if (lane && sm.I(PI_MODE, p) == MODE_TRIAL) {
                    const int r = ctrl_linesearch(prm, sm, c, p);
                    if (r == 1) sm.I(PI_MODE, p) = MODE_ACCEPT;
                    else if (r < 0) { c.status = 9; sm.I(PI_STATUS, p) = 9; sm.I(PI_MODE, p) = MODE_DONE; }
                    else again = 1;
                }
                PROF_MARK(8);
                if (!__syncthreads_or(again)) break;  // B8
            }
            if (lane && sm.I(PI_MODE, p) == MODE_ACCEPT && !sm.I(PI_LSQ, p)) ctrl_accept(sm, c, p);
            __syncthreads();  // B9: step applied
            PROF_MARK(9);
            if (lane) sm.I(PI_LSQ, p) = 0;
        }
        PROF_FLUSH();
    } else {
        // ------------------------------------------------------------ stage threads
        const int t = tid - 32;
        const int p = t % PB;
        const int g = t / PB;
        const int k0 = g * SPT;
        const bool mine = k0 < N;
        StageRegs r[SPT];
        __syncthreads();  // R1 (initial)
        if (mine && sm.I(PI_NEXT, p) >= 0) {
            const int idx = sm.I(PI_NEXT, p);
            double s6[6], c4[4];
            for (int i = 0; i < 6; i++) s6[i] = a.state[(size_t)i * batch + idx];
            for (int i = 0; i < 4; i++) c4[i] = a.coeffs[(size_t)i * batch + idx];
#pragma unroll
            for (int j = 0; j < SPT; j++)
                if (k0 + j < N) stage_init(prm, sm, r[j], k0 + j, p, s6, c4);
        }
        __syncthreads();  // R2 (initial) == B0
        for (;;) {
            if (mine) {
                const int md = sm.I(PI_MODE, p);
                if (md == MODE_RESID || md == MODE_ACCEPT) {
                    ResidPart acc;
                    part_reset(acc);
#pragma unroll
                    for (int j = 0; j < SPT; j++)
                        if (k0 + j < N) stage_residuals(prm, sm, r[j], k0 + j, p, acc);
                    part_store(sm, k0, p, acc);
                }
            }
            __syncthreads();  // B1
            if (__syncthreads_or(0)) {  // V1
                // ---- flush the finished problem: the last iterate whatever the status (mpc_planner.cpp:378-401)
                if (mine && sm.I(PI_MODE, p) == MODE_DONE) {
                    const size_t i = (size_t)sm.I(PI_PROB, p);
#pragma unroll
                    for (int j = 0; j < SPT; j++) {
                        const int k = k0 + j;
                        if (k >= N) continue;
                        a.pred[((size_t)0 * N + k) * batch + i] = sm.at(k, S_X, p);
                        a.pred[((size_t)1 * N + k) * batch + i] = sm.at(k, S_Y, p);
                        a.pred[((size_t)2 * N + k) * batch + i] = sm.at(k, S_T, p);
                        if (k == 0) { a.u0[i] = r[j].uw; a.u0[(size_t)batch + i] = r[j].ua; }
                        if (a.warm_out) {
                            // primal in the reference's variable layout (mpc_planner.cpp:232-239), then equality
                            // multipliers (row layout of :153-158, unscaled), then zL, zU of w and a.
                            double *wo = a.warm_out;
                            const double sf = sm.P(PS_SF, p);
                            for (int cc = 0; cc < 6; cc++) wo[((size_t)cc * N + k) * batch + i] = sm.at(k, S_X + cc, p);
                            const size_t offl = (size_t)(8 * N - 2);
                            if (k < N - 1) {
                                wo[((size_t)6 * N + k) * batch + i] = r[j].uw;
                                wo[((size_t)7 * N - 1 + k) * batch + i] = r[j].ua;
                                for (int cc = 0; cc < 6; cc++)
                                    wo[(offl + (size_t)cc * N + k + 1) * batch + i] = sm.at(k, L_X + cc, p) / sf;
                                const size_t offz = offl + (size_t)6 * N;
                                const int nu = N - 1;
                                wo[(offz + k) * batch + i] = r[j].zlw / sf;
                                wo[(offz + nu + k) * batch + i] = r[j].zla / sf;
                                wo[(offz + 2 * nu + k) * batch + i] = r[j].zuw / sf;
                                wo[(offz + 3 * nu + k) * batch + i] = r[j].zua / sf;
                            }
                            if (k == 0)
                                for (int cc = 0; cc < 6; cc++)
                                    wo[(offl + (size_t)cc * N) * batch + i] = sm.P(PS_L0X + cc, p) / sf;
                        }
                    }
                }
                __syncthreads();  // R1
                if (mine && sm.I(PI_NEXT, p) >= 0) {
                    const int idx = sm.I(PI_NEXT, p);
                    double s6[6], c4[4];
                    for (int i = 0; i < 6; i++) s6[i] = a.state[(size_t)i * batch + idx];
                    for (int i = 0; i < 4; i++) c4[i] = a.coeffs[(size_t)i * batch + idx];
#pragma unroll
                    for (int j = 0; j < SPT; j++)
                        if (k0 + j < N) stage_init(prm, sm, r[j], k0 + j, p, s6, c4);
                }
                __syncthreads();  // R2
            }
            if (!__syncthreads_or(0)) break;  // V2
            for (;;) {
                if (mine && sm.I(PI_MODE, p) == MODE_COEF) {
                    const int lsq = sm.I(PI_LSQ, p);
#pragma unroll
                    for (int j = 0; j < SPT; j++)
                        if (k0 + j < N) stage_coeffs(prm, sm, r[j], k0 + j, p, lsq);
                }
                __syncthreads();  // B3
                if (!__syncthreads_or(0)) break;  // B4
            }
            if (mine && sm.I(PI_MODE, p) == MODE_STEP) {
                const int lsq = sm.I(PI_LSQ, p);
                const HessDiag hd = hess_diag(prm, sm.P(PS_SF, p), sm.P(PS_DW, p), lsq);
                StepPart acc;
                part_reset(acc);
#pragma unroll
                for (int j = 0; j < SPT; j++)
                    if (k0 + j < N) stage_step(prm, sm, r[j], k0 + j, p, hd, lsq, acc);
                part_store(sm, k0, p, acc);
            }
            __syncthreads();  // B5
            __syncthreads();  // B6
            for (;;) {
                if (mine && sm.I(PI_MODE, p) == MODE_TRIAL) {
                    TrialPart acc;
                    part_reset(acc);
#pragma unroll
                    for (int j = 0; j < SPT; j++)
                        if (k0 + j < N) stage_trial(prm, sm, r[j], k0 + j, p, acc);
                    part_store(sm, k0, p, acc);
                }
                __syncthreads();  // B7
                if (!__syncthreads_or(0)) break;  // B8
            }
            if (mine && sm.I(PI_MODE, p) == MODE_ACCEPT) {
                const int lsq = sm.I(PI_LSQ, p);
                // stage k reads ds_k from stage k-1's D slots and writes only its own S/L slots
#pragma unroll
                for (int j = 0; j < SPT; j++)
                    if (k0 + j < N) stage_accept(prm, sm, r[j], k0 + j, p, lsq);
            }
            __syncthreads();  // B9
        }
    }
}

}  // namespace nmpc
