// tests/emu/nmpc_emu.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Runs the solver's phase functions (mpc_ros_b200/csrc/nmpc_phases.cuh, the code the CUDA
// kernel executes) sequentially on the host, one emulated CTA of PB lanes at a time, with the
// kernel's cycle / barrier structure (nmpc_kernel.cuh) turned into plain loops -- including the
// per-lane state machine.  (The kernel's work queue is not emulated: lanes get their problems
// statically; results do not depend on the lane, which the tests check.)  It exists so that the
// solver LOGIC can be checked against the oracle in the CPU-only test tier (this container has no
// GPU).  It is built only by tests/ into tests/emu/libnmpc_emu.so, is never linked into
// libmpc_b200.so, and nothing in the product path can reach it.
#include "../../mpc_ros_b200/csrc/nmpc_phases.cuh"
#include <vector>
#include <cstring>

using namespace nmpc;

static long long g_lane_cycles = 0;
static std::vector<int> g_ages, g_kinds;   // g_kinds: 4 per problem: backtracks, corrections, resumes, inertia retries
static std::vector<int> g_kind_lane;
static std::vector<int> g_ages_unused;      // global cycles each problem of the last run took (tail studies)

template <bool RATE, int NC = 4>
static int emu_run(int N, const double *prm14, double tol, int max_iter, int PB, int batch, int ncoef,
                   const double *state, const double *coeffs, const double *ref_vel,
                   double *u0, double *pred, double *obj, int *status, int *iters, double *kkt,
                   double *lam_out, int *n_reg)
{
    Params prm;
    prm.N = N;
    prm.dt = prm14[0]; prm.ref_cte = prm14[1]; prm.ref_etheta = prm14[2]; prm.ref_vel = prm14[3];
    prm.w_cte = prm14[4]; prm.w_etheta = prm14[5]; prm.w_vel = prm14[6]; prm.w_angvel = prm14[7];
    prm.w_accel = prm14[8]; prm.max_angvel = prm14[9]; prm.max_throttle = prm14[10];
    prm.tol = tol; prm.max_iter = max_iter;
    prm.warm_mu = 1e-3;
    prm.w_angvel_d = prm14[11]; prm.w_accel_d = prm14[12];
    const int SPT = 2;   // same grouping of partial sums as the kernel's stage threads
    prm.grp = SPT;
    prm.idt = 1.0 / prm.dt;
    prm.i_mnb = 1.0 / (double)(6 * N + 4 * (N - 1)); prm.i_nb = 1.0 / (double)(4 * (N - 1));
    prm.bound_chk = (prm14[13] > 0.0 ? prm14[13] : 1e3) * (1.0 - 1e-3);   // prm14[13]: bound_value (0 = the 1e3 default)
    const int NG = (N + SPT - 1) / SPT;
    g_ages.assign((size_t)batch, 0); g_kinds.assign((size_t)4 * batch, 0); g_kind_lane.assign((size_t)4 * PB, 0);

    const int NS = RATE ? NSLOTS_RATE : NSLOTS;
    std::vector<double> mem(smem_bytes(N, NG, PB, NS) / sizeof(double) + 8);
    SmemT<0, RATE ? NSLOTS_RATE : NSLOTS> sm; sm.PB = PB; sm.carve(mem.data(), N, NG);
    std::vector<StageRegs> regs((size_t)N * PB);
    std::vector<double> cfs((size_t)NC * PB);
    std::vector<Ctrl> ctrl(PB);
    std::vector<int> nreg(PB);
#define REG(k, p) regs[(size_t)(k) * PB + (p)]

    for (int base = 0; base < batch; base += PB) {
        const int np = (batch - base < PB) ? batch - base : PB;
        std::fill(mem.begin(), mem.end(), 0.0);
        for (int p = 0; p < PB; p++) { sm.I(PI_MODE, p) = MODE_IDLE; sm.I(PI_FLAGS, p) = 0; sm.I(PI_PROB, p) = -1; sm.I(PI_NEXT, p) = -1; }
        // ---- "refill": every lane starts its problem
        for (int p = 0; p < np; p++) {
            double s6[6];
            for (int c = 0; c < 6; c++) s6[c] = state[(size_t)c * batch + base + p];
            const double rv = ref_vel ? ref_vel[base + p] : prm.ref_vel;
            ctrl_init(prm, sm, ctrl[p], p, s6, rv);
            sm.I(PI_NEXT, p) = base + p;
            sm.I(PI_MODE, p) = MODE_NEWTON; sm.I(PI_FLAGS, p) = FL_LSQ;
            nreg[p] = 0; for (int q = 0; q < 4; q++) g_kind_lane[4 * p + q] = 0;
        }
        for (;;) {
            int active = 0;
            for (int p = 0; p < np; p++) active |= (sm.I(PI_MODE, p) != MODE_IDLE) || (sm.I(PI_FLAGS, p) & FL_FLUSH);
            if (!active) break;
            for (int p = 0; p < np; p++) g_lane_cycles += (sm.I(PI_MODE, p) != MODE_IDLE);
            // ---- P3a
            for (int p = 0; p < np; p++) {
                const int fl = sm.I(PI_FLAGS, p);
                if (fl & FL_APPLY) for (int k = 0; k < N; k++) stage_apply<RATE>(prm, sm, REG(k, p), k, p);
                if (fl & FL_FLUSH) {
                    const size_t i = (size_t)sm.I(PI_PROB, p);
                    if (status)
                        for (int k = 0; k < N; k++)
                            if (stage_bound_hit(prm, sm, k, p) && (status[i] == 1 || status[i] == 4)) status[i] = NMPC_STATUS_BOUND_ACTIVE;
                    u0[i] = REG(0, p).uw; u0[(size_t)batch + i] = REG(0, p).ua;
                    for (int k = 0; k < N; k++) {
                        pred[((size_t)0 * N + k) * batch + i] = sm.at(k, S_X, p);
                        pred[((size_t)1 * N + k) * batch + i] = sm.at(k, S_Y, p);
                        pred[((size_t)2 * N + k) * batch + i] = sm.at(k, S_T, p);
                    }
                    if (lam_out) {
                        const double sf = sm.P(PS_AP_SF, p);
                        for (int c = 0; c < 6; c++)
                            for (int k = 0; k < N - 1; k++)
                                lam_out[((size_t)c * N + k + 1) * batch + i] = sm.at(k, L_X + c, p) / sf;
                    }
                }
                const int idx = sm.I(PI_NEXT, p);
                if (idx >= 0) {
                    double s6[6], c4[4];
                    for (int c = 0; c < 6; c++) s6[c] = state[(size_t)c * batch + idx];
                    for (int c = 0; c < 4; c++) c4[c] = coeffs[(size_t)c * batch + idx];
                    for (int c = 0; c < NC; c++) cfs[(size_t)NC * p + c] = c < ncoef ? coeffs[(size_t)c * batch + idx] : 0.0;
                    for (int c = 0; c < 4; c++) sm.P(PS_NX6 + c, p) = c4[c];        // (the kernel's refill stages them here)
                    for (int c = 4; c < NC; c++) sm.P(PS_NXC4 + (c - 4), p) = c < ncoef ? coeffs[(size_t)c * batch + idx] : 0.0;
                    for (int k = 0; k < N; k++) stage_init<RATE>(prm, sm, REG(k, p), k, p, s6, c4);
                }
            }
            // second-order correction: next right-hand side from the rejected trial point, before any stage's coefficients
            // overwrite the step slots
            for (int p = 0; p < np; p++)
                if (sm.I(PI_MODE, p) == MODE_NEWTON && (sm.I(PI_FLAGS, p) & (FL_SOC | FL_APPLY)) == FL_SOC)
                    for (int k = 0; k < N; k++) stage_soc_rhs<NC>(prm, sm, REG(k, p), k, p, &cfs[(size_t)NC * p]);
            for (int p = 0; p < np; p++) {
                if (sm.I(PI_FLAGS, p) & FL_APPLY) sm.I(PI_FLAGS, p) &= ~FL_SOC;
                sm.I(PI_FLAGS, p) &= ~(FL_APPLY | FL_FLUSH);
                if (sm.I(PI_NEXT, p) >= 0) { sm.I(PI_PROB, p) = sm.I(PI_NEXT, p); sm.I(PI_NEXT, p) = -1; }
            }
            // ---- P3b
            for (int p = 0; p < np; p++)
                if (sm.I(PI_MODE, p) == MODE_NEWTON)
                    for (int k = 0; k < N; k++) stage_coeffs<RATE, NC>(prm, sm, REG(k, p), k, p, sys_kind(sm.I(PI_FLAGS, p)), &cfs[(size_t)NC * p], sm.I(PI_FLAGS, p) & FL_SOC);
            // ---- P4
            for (int p = 0; p < np; p++) {
                if (sm.I(PI_MODE, p) != MODE_NEWTON) continue;
                Ctrl &c = ctrl[p];
                const int lsq = sys_kind(sm.I(PI_FLAGS, p));
                const double dw = sm.P(PS_DW, p);
                const HessDiag hd = hess_diag(prm, sm.P(PS_SF, p), dw, lsq);
                if (riccati_backward<RATE>(prm, sm, p, hd) || lsq) {
                    riccati_forward<RATE>(prm, sm, p, hd);
                    if (dw > 0.0) c.dw_last = dw;
                    sm.I(PI_MODE, p) = MODE_STEP;
                } else {
                    const double nd = next_dw(c, dw);
                    nreg[p]++;
                    if (nd > NMPC_DW_MAX) { c.status = 10; sm.I(PI_MODE, p) = MODE_FAIL; }
                    else sm.P(PS_DW, p) = nd;
                }
            }
            // ---- P5
            for (int p = 0; p < np; p++) {
                if (sm.I(PI_MODE, p) != MODE_STEP) continue;
                const int lsq = sys_kind(sm.I(PI_FLAGS, p));
                const HessDiag hd = hess_diag(prm, sm.P(PS_SF, p), sm.P(PS_DW, p), lsq);
                for (int g = 0; g < NG; g++) {
                    StepPart acc; part_reset(acc);
                    double gk[SPT][6];
                    for (int k = g * SPT; k < g * SPT + SPT && k < N; k++) stage_step<RATE, NC>(prm, sm, REG(k, p), k, p, hd, lsq, acc, gk[k - g * SPT], &cfs[(size_t)NC * p]);
                    for (int k = g * SPT; k < g * SPT + SPT && k < N; k++)
                        for (int q = 0; q < 6; q++) sm.at(k, W_0 + q, p) = gk[k - g * SPT][q];
                    part_store(sm, g, p, acc);
                }
            }
            // ---- P6
            for (int p = 0; p < np; p++) {
                if (sm.I(PI_MODE, p) != MODE_STEP) continue;
                const int big = adjoint_sweep(prm, sm, p);
                if (sm.I(PI_FLAGS, p) & FL_LSQ) {
                    const int keep = ctrl_lsq_finish(prm, sm, p, big);
                    sm.I(PI_FLAGS, p) = FL_ADOPT | keep;
                } else {
                    const int mode = sm.I(PI_FLAGS, p) & (FL_SOC | FL_RESUME | FL_RESTO);
                    ctrl_step(prm, sm, ctrl[p], p, NG, mode);
                    ctrl_step_late(ctrl[p]);
                    sm.I(PI_FLAGS, p) = FL_LS | (mode & (FL_SOC | FL_RESTO));
                }
                sm.I(PI_MODE, p) = MODE_EVAL;
            }
            // ---- P1
            for (int p = 0; p < np; p++) {
                if (sm.I(PI_MODE, p) != MODE_EVAL) continue;
                const int fl = sm.I(PI_FLAGS, p);
                // all stages read their neighbours' slots before FL_ADOPT overwrites the L slots of a stage:
                // stage k writes only its own L slots and reads lambda^+ of stage k-1 from W, as in the kernel
                if (fl & FL_ADOPT) for (int k = 0; k < N; k++) stage_adopt(prm, sm, k, p, fl);
                for (int g = 0; g < NG; g++) {
                    EvalPart acc; part_reset(acc);
                    for (int k = g * SPT; k < g * SPT + SPT && k < N; k++) stage_eval<RATE, NC>(prm, sm, REG(k, p), k, p, fl, acc, &cfs[(size_t)NC * p]);
                    part_store(sm, g, p, acc);
                }
            }
            // ---- P2
            for (int p = 0; p < np; p++) {
                Ctrl &c = ctrl[p];
                const int md = sm.I(PI_MODE, p);
                int term = 0;
                if (md == MODE_EVAL) {
                    const int fl = sm.I(PI_FLAGS, p);
                    const int r = ctrl_decide(prm, sm, c, p, fl, NG);
                    if (r == 0) g_kind_lane[4 * p]++; else if (r == 3) g_kind_lane[4 * p + 1]++; else if (r == 4) g_kind_lane[4 * p + 2]++;
                    if (r == 5 || r == 6) g_kind_lane[4 * p + 1]++;      // (restoration steps are counted with the corrections)
                    if (r == 0) sm.I(PI_FLAGS, p) = FL_LS | (fl & FL_RESTO);
                    else if (r == 3) { sm.I(PI_MODE, p) = MODE_NEWTON; sm.I(PI_FLAGS, p) = FL_SOC; }
                    else if (r == 4) { sm.I(PI_MODE, p) = MODE_NEWTON; sm.I(PI_FLAGS, p) = FL_RESUME; }
                    else if (r == 5) { sm.I(PI_MODE, p) = MODE_NEWTON; sm.I(PI_FLAGS, p) = FL_RESTO; sm.P(PS_DW, p) = 0.0; }
                    else if (r == 6) {
                        sm.P(PS_AP_ALPHA, p) = sm.P(PS_ALPHA, p); sm.P(PS_AP_AZ, p) = NMPC_AZ_RESTO; sm.P(PS_AP_MU, p) = sm.P(PS_MU_STEP, p);
                        sm.I(PI_MODE, p) = MODE_NEWTON; sm.I(PI_FLAGS, p) = FL_APPLY | FL_RESTO; sm.P(PS_DW, p) = 0.0;
                    } else {
                        int nf = 0;
                        // the evaluated point becomes the iterate: P3a applies the step (a plain evaluation is
                        // a step of length 0 -- it still adopts the sin/cos computed at the evaluated point)
                        nf = FL_APPLY;
                        if ((fl & (FL_LS | FL_RESTO)) == (FL_LS | FL_RESTO)) {
                            sm.P(PS_AP_ALPHA, p) = sm.P(PS_ALPHA, p); sm.P(PS_AP_AZ, p) = NMPC_AZ_RESTO_END;
                        } else if (fl & FL_LS) {
                            sm.P(PS_AP_ALPHA, p) = sm.P(PS_ALPHA, p); sm.P(PS_AP_AZ, p) = sm.P(PS_ALPHA_Z, p);
                            ctrl_apply(sm, p);
                        } else {
                            sm.P(PS_AP_ALPHA, p) = 0.0; sm.P(PS_AP_AZ, p) = 0.0;
                        }
                        sm.P(PS_AP_MU, p) = sm.P(PS_MU_STEP, p);
                        if (r == 1) { sm.I(PI_MODE, p) = MODE_NEWTON; sm.P(PS_DW, p) = 0.0; sm.I(PI_FLAGS, p) = nf; }
                        else { term = 1; sm.I(PI_FLAGS, p) = nf | FL_FLUSH; }
                    }
                } else if (md == MODE_FAIL) {
                    term = 1; sm.I(PI_FLAGS, p) = FL_FLUSH;
                }
                // watchdog: retries and backtracks are bounded, this only guards against the unforeseen
                if (!term && md != MODE_IDLE && ++c.age > 8 * prm.max_iter + 64) {
                    if (c.status == 0) c.status = 2;
                    term = 1; sm.I(PI_FLAGS, p) = FL_FLUSH;
                }
                if (term) {
                    const size_t i = (size_t)sm.I(PI_PROB, p);
                    if (obj) obj[i] = c.obj;
                    if (status) status[i] = c.status;
                    if (iters) iters[i] = c.iter;
                    if (kkt) kkt[i] = c.E0;
                    if (n_reg) n_reg[i] = nreg[p];
                    g_ages[i] = c.age + 1; g_kind_lane[4 * p + 3] = nreg[p];
                    for (int q = 0; q < 4; q++) g_kinds[4 * i + q] = g_kind_lane[4 * p + q];
                    sm.P(PS_AP_SF, p) = sm.P(PS_SF, p);
                    if (lam_out)
                        for (int cc = 0; cc < 6; cc++)
                            lam_out[((size_t)cc * N) * batch + i] = sm.P(PS_L0X + cc, p) / sm.P(PS_SF, p);
                    sm.I(PI_MODE, p) = MODE_IDLE;
                }
            }
        }
    }
    return 0;
}

extern "C" int nmpc_emu_last_ages(int *out, int n) { int m = (int)g_ages.size() < n ? (int)g_ages.size() : n; for (int i = 0; i < m; i++) out[i] = g_ages[i]; return m; }

extern "C" int nmpc_emu_last_kinds(int *out, int n) { int m = (int)g_kinds.size() < 4 * n ? (int)g_kinds.size() : 4 * n; for (int i = 0; i < m; i++) out[i] = g_kinds[i]; return m; }

// lane-cycles (one lane busy for one global cycle) spent since the last reset: utilisation studies
extern "C" long long nmpc_emu_lane_cycles(int reset) { const long long v = g_lane_cycles; if (reset) g_lane_cycles = 0; return v; }

extern "C" int nmpc_emu_solve(int N, const double *prm14, double tol, int max_iter, int PB, int batch,
                              const double *state, const double *coeffs, const double *ref_vel,
                              double *u0, double *pred, double *obj, int *status, int *iters, double *kkt,
                              double *lam_out /* 6N x batch, optional */, int *n_reg /* batch, optional */)
{
    // prm14[11], prm14[12]: w_angvel_d, w_accel_d (rate penalties) select the augmented-Riccati variant
    if (prm14[11] != 0.0 || prm14[12] != 0.0)
        return emu_run<true>(N, prm14, tol, max_iter, PB, batch, 4, state, coeffs, ref_vel, u0, pred, obj, status, iters, kkt,
                             lam_out, n_reg);
    return emu_run<false>(N, prm14, tol, max_iter, PB, batch, 4, state, coeffs, ref_vel, u0, pred, obj, status, iters, kkt,
                          lam_out, n_reg);
}

// Path polynomial of order ncoef - 1 in 4 .. NMPC_MAX_COEFFS - 1 (coeffs: ncoef x batch): the instantiation the
// kernel uses for orders above 3.
extern "C" int nmpc_emu_solve_poly(int N, const double *prm14, double tol, int max_iter, int PB, int batch, int ncoef,
                                   const double *state, const double *coeffs, double *u0, double *pred, double *obj,
                                   int *status, int *iters, double *kkt)
{
    if (ncoef < 4 || ncoef > NMPC_MAX_COEFFS) return -1;
    if (prm14[11] != 0.0 || prm14[12] != 0.0)
        return emu_run<true, NMPC_MAX_COEFFS>(N, prm14, tol, max_iter, PB, batch, ncoef, state, coeffs, nullptr, u0, pred, obj,
                                              status, iters, kkt, nullptr, nullptr);
    return emu_run<false, NMPC_MAX_COEFFS>(N, prm14, tol, max_iter, PB, batch, ncoef, state, coeffs, nullptr, u0, pred, obj,
                                           status, iters, kkt, nullptr, nullptr);
}
