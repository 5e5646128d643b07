// tests/emu/nmpc_emu.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Runs the solver's phase functions (mpc_ros_b200/csrc/nmpc_phases.cuh, the code the CUDA
// kernel executes) sequentially on the host, one emulated CTA of PB problems at a time, with
// the kernel's barrier structure turned into plain loops.  It exists so that the solver LOGIC
// can be checked against the oracle in the CPU-only test tier (this container has no GPU).
// It is built only by tests/ (tests/emu/Makefile) into tests/emu/libnmpc_emu.so, is never
// linked into libmpc_b200.so, and nothing in the product path can reach it.
#include "../../mpc_ros_b200/csrc/nmpc_phases.cuh"
#include <vector>
#include <cstring>

using namespace nmpc;

extern "C" int nmpc_emu_solve(int N, const double *prm14, double tol, int max_iter, int PB, int batch,
                              const double *state, const double *coeffs, const double *ref_vel,
                              double *u0, double *pred, double *obj, int *status, int *iters, double *kkt,
                              double *lam_out /* 6N x batch, optional */, int *n_reg /* batch, optional */)
{
    Params prm;
    prm.N = N;
    prm.dt = prm14[0]; prm.ref_cte = prm14[1]; prm.ref_etheta = prm14[2]; prm.ref_vel = prm14[3];
    prm.w_cte = prm14[4]; prm.w_etheta = prm14[5]; prm.w_vel = prm14[6]; prm.w_angvel = prm14[7];
    prm.w_accel = prm14[8]; prm.max_angvel = prm14[9]; prm.max_throttle = prm14[10];
    prm.tol = tol; prm.max_iter = max_iter;
    prm.grp = 3;   // same grouping of partial sums as the kernel's stage threads (SPT = 3)

    std::vector<double> st((size_t)N * NSLOTS * PB), ps((size_t)NPS * PB);
    std::vector<int> pi((size_t)NPI * PB);
    Smem sm; sm.st = st.data(); sm.ps = ps.data(); sm.pi = pi.data(); sm.PB = PB;
    std::vector<StageRegs> regs((size_t)N * PB);
    std::vector<Ctrl> ctrl(PB);
    std::vector<int> nreg(PB);

    for (int base = 0; base < batch; base += PB) {
        const int np = (batch - base < PB) ? batch - base : PB;
        std::fill(st.begin(), st.end(), 0.0);
        for (int p = 0; p < PB; p++) sm.I(PI_MODE, p) = MODE_IDLE;
        // ---- init
        for (int p = 0; p < np; p++) {
            double s6[6], c4[4];
            for (int c = 0; c < 6; c++) s6[c] = state[(size_t)c * batch + base + p];
            for (int c = 0; c < 4; c++) c4[c] = coeffs[(size_t)c * batch + base + p];
            const double rv = ref_vel ? ref_vel[base + p] : prm.ref_vel;
            for (int k = 0; k < N; k++) stage_init(prm, sm, regs[(size_t)k * PB + p], k, p, s6, c4);
            ctrl_init(prm, sm, ctrl[p], p, s6, rv);
            nreg[p] = 0;
        }
        // ---- cycles (same phase / barrier structure as nmpc_solve_kernel)
        for (;;) {
            for (int p = 0; p < np; p++) {
                const int md = sm.I(PI_MODE, p);
                if (md == MODE_RESID || md == MODE_ACCEPT)
                    for (int k0 = 0; k0 < N; k0 += prm.grp) {
                        ResidPart acc; part_reset(acc);
                        for (int k = k0; k < k0 + prm.grp && k < N; k++) stage_residuals(prm, sm, regs[(size_t)k * PB + p], k, p, acc);
                        part_store(sm, k0, p, acc);
                    }
            }
            int any_run = 0;
            for (int p = 0; p < np; p++) {
                const int md = sm.I(PI_MODE, p);
                if (md == MODE_RESID || md == MODE_ACCEPT) {
                    if (sm.I(PI_LSQ, p) || ctrl_check(prm, sm, ctrl[p], p)) { sm.I(PI_MODE, p) = MODE_COEF; sm.P(PS_DW, p) = 0.0; }
                    else { sm.I(PI_MODE, p) = MODE_IDLE; sm.I(PI_STATUS, p) = ctrl[p].status; }
                }
                any_run |= sm.I(PI_MODE, p) != MODE_IDLE;
            }
            if (!any_run) break;
            // Newton system, with inertia-correction retries
            for (;;) {
                for (int p = 0; p < np; p++)
                    if (sm.I(PI_MODE, p) == MODE_COEF)
                        for (int k = 0; k < N; k++) stage_coeffs(prm, sm, regs[(size_t)k * PB + p], k, p, sm.I(PI_LSQ, p));
                int any = 0;
                for (int p = 0; p < np; p++) {
                    if (sm.I(PI_MODE, p) != MODE_COEF) continue;
                    Ctrl &c = ctrl[p];
                    const double dw = sm.P(PS_DW, p);
                    HessDiag hd = hess_diag(prm, sm.P(PS_SF, p), dw, sm.I(PI_LSQ, p));
                    if (riccati_backward(prm, sm, p, hd) || sm.I(PI_LSQ, p)) {
                        riccati_forward(prm, sm, p);
                        if (dw > 0.0) c.dw_last = dw;
                        sm.I(PI_MODE, p) = MODE_STEP;
                    } else {
                        const double nd = next_dw(c, dw);
                        nreg[p]++;
                        if (nd > NMPC_DW_MAX) { c.status = 10; sm.I(PI_STATUS, p) = 10; sm.I(PI_MODE, p) = MODE_IDLE; }
                        else { sm.P(PS_DW, p) = nd; any = 1; }
                    }
                }
                if (!any) break;
            }
            for (int p = 0; p < np; p++)
                if (sm.I(PI_MODE, p) == MODE_STEP) {
                    HessDiag hd = hess_diag(prm, sm.P(PS_SF, p), sm.P(PS_DW, p), sm.I(PI_LSQ, p));
                    for (int k0 = 0; k0 < N; k0 += prm.grp) {
                        StepPart acc; part_reset(acc);
                        for (int k = k0; k < k0 + prm.grp && k < N; k++) stage_step(prm, sm, regs[(size_t)k * PB + p], k, p, hd, sm.I(PI_LSQ, p), acc);
                        part_store(sm, k0, p, acc);
                    }
                }
            for (int p = 0; p < np; p++)
                if (sm.I(PI_MODE, p) == MODE_STEP) {
                    if (sm.I(PI_LSQ, p)) { ctrl_lsq_finish(prm, sm, ctrl[p], p); sm.I(PI_MODE, p) = MODE_ACCEPT; }
                    else { ctrl_step(prm, sm, ctrl[p], p); sm.I(PI_MODE, p) = MODE_TRIAL; }
                }
            // line search
            for (;;) {
                for (int p = 0; p < np; p++)
                    if (sm.I(PI_MODE, p) == MODE_TRIAL)
                        for (int k0 = 0; k0 < N; k0 += prm.grp) {
                            TrialPart acc; part_reset(acc);
                            for (int k = k0; k < k0 + prm.grp && k < N; k++) stage_trial(prm, sm, regs[(size_t)k * PB + p], k, p, acc);
                            part_store(sm, k0, p, acc);
                        }
                int any = 0;
                for (int p = 0; p < np; p++) {
                    if (sm.I(PI_MODE, p) != MODE_TRIAL) continue;
                    const int r = ctrl_linesearch(prm, sm, ctrl[p], p);
                    if (r == 1) sm.I(PI_MODE, p) = MODE_ACCEPT;
                    else if (r < 0) { ctrl[p].status = 9; sm.I(PI_STATUS, p) = 9; sm.I(PI_MODE, p) = MODE_IDLE; }
                    else any = 1;
                }
                if (!any) break;
            }
            for (int p = 0; p < np; p++)
                if (sm.I(PI_MODE, p) == MODE_ACCEPT) {
                    for (int k = 0; k < N; k++) stage_accept(prm, sm, regs[(size_t)k * PB + p], k, p, sm.I(PI_LSQ, p));
                    if (!sm.I(PI_LSQ, p)) ctrl_accept(sm, ctrl[p], p);
                }
            for (int p = 0; p < np; p++) sm.I(PI_LSQ, p) = 0;
        }
        // ---- outputs (the reference returns the last iterate whatever the status, mpc_planner.cpp:378-401)
        for (int p = 0; p < np; p++) {
            const size_t i = (size_t)base + p;
            u0[i] = regs[p].uw; u0[(size_t)batch + i] = regs[p].ua;
            for (int k = 0; k < N; k++) {
                pred[((size_t)0 * N + k) * batch + i] = sm.at(k, S_X, p);
                pred[((size_t)1 * N + k) * batch + i] = sm.at(k, S_Y, p);
                pred[((size_t)2 * N + k) * batch + i] = sm.at(k, S_T, p);
            }
            if (obj) obj[i] = ctrl[p].obj;
            if (status) status[i] = ctrl[p].status;
            if (iters) iters[i] = ctrl[p].iter;
            if (kkt) kkt[i] = ctrl[p].E0;
            if (n_reg) n_reg[i] = nreg[p];
            if (lam_out) {
                // reference row layout: component-major, row comp*N + k (mpc_planner.cpp:153-158); unscaled
                const double sf = sm.P(PS_SF, p);
                for (int c = 0; c < 6; c++) {
                    lam_out[((size_t)c * N + 0) * batch + i] = sm.P(PS_L0X + c, p) / sf;
                    for (int k = 0; k < N - 1; k++)
                        lam_out[((size_t)c * N + k + 1) * batch + i] = sm.at(k, L_X + c, p) / sf;
                }
            }
        }
    }
    return 0;
}
