// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C entry points around the reference's UNMODIFIED MPC class
// (mpc_ros/include/mpc_planner.h:26-47, compiled from
// /root/reference/mpc_ros/src/mpc_planner.cpp where it lies; see
// oracle/Makefile).  Built into oracle/_ref/libmpc_ref.so.  The NLP solver
// behind CppAD::ipopt::solve is oracle/ipm.c, NOT Ipopt 3.12.8.
#include "mpc_planner.h"
#include "shim/ip_standin.h"
#include <cstring>
#include <iostream>
#include <sstream>
#include <map>
#include <string>

namespace {
// keys of driving_state.cpp:65-79, in that order
const char *kKeys[15] = { "DT", "STEPS", "REF_CTE", "REF_ETHETA", "REF_V", "W_CTE", "W_EPSI", "W_V",
                          "W_ANGVEL", "W_A", "W_DANGVEL", "W_DA", "ANGVEL", "MAXTHR", "BOUND" };

MPC *make_mpc(const double *params15)
{
    // MPC::MPC prints "init mpc" (mpc_planner.cpp:225); keep stdout clean for the harness.
    std::streambuf *old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    MPC *mpc = new MPC();
    std::cout.rdbuf(old);
    std::map<std::string, double> p;
    for (int i = 0; i < 15; i++) p[kKeys[i]] = params15[i];
    mpc->LoadParams(p);
    return mpc;
}
}  // namespace

extern "C" {

void *ref_mpc_create(const double *params15) { return make_mpc(params15); }
void ref_mpc_destroy(void *h) { delete static_cast<MPC *>(h); }
void ref_set_cpu_time_override(double s) { standin_set_cpu_time_override(s); }
void ref_set_dense_ldl(int on) { standin_set_dense_ldl(on); }

// One MPC::Solve.  u0[2]; pred[3*N] = mpc_x, mpc_y, mpc_theta; info[10] =
// status, iters, obj, kkt_error, dual_inf, constr_viol, compl_inf, n_inertia, n_resto, n_fact;
// sol (8N-2), lambda (6N), zl, zu (8N-2) optional.
int ref_mpc_solve(void *h, int N, const double *state6, const double *coeffs, int ncoef,
                  double *u0, double *pred, double *info, double *sol, double *lambda, double *zl, double *zu)
{
    MPC *mpc = static_cast<MPC *>(h);
    Eigen::VectorXd st(6), co(ncoef);
    for (int i = 0; i < 6; i++) st[i] = state6[i];
    for (int i = 0; i < ncoef; i++) co[i] = coeffs[i];
    standin_probe().active = false;
    std::vector<double> r = mpc->Solve(st, co);
    u0[0] = r[0]; u0[1] = r[1];
    if (pred) for (int i = 0; i < N; i++) { pred[i] = mpc->mpc_x[i]; pred[N + i] = mpc->mpc_y[i]; pred[2 * N + i] = mpc->mpc_theta[i]; }
    const StandinLast &L = standin_last();
    if (info) {
        info[0] = L.status; info[1] = L.iters; info[2] = L.obj; info[3] = L.kkt_error; info[4] = L.dual_inf;
        info[5] = L.constr_viol; info[6] = L.compl_inf; info[7] = L.n_inertia; info[8] = L.n_resto; info[9] = L.n_fact;
    }
    if (sol) std::memcpy(sol, L.x.data(), sizeof(double) * L.x.size());
    if (lambda) std::memcpy(lambda, L.lambda.data(), sizeof(double) * L.lambda.size());
    if (zl) std::memcpy(zl, L.zl.data(), sizeof(double) * L.zl.size());
    if (zu) std::memcpy(zu, L.zu.data(), sizeof(double) * L.zu.size());
    return L.status;
}

// Evaluate what the reference's FG_eval + CppAD hand to Ipopt at a given point:
// f, grad f (n), g (m), dense Jacobian (m*n), dense symmetric Hessian of sigma f + lambda'g (n*n).
// nnz[2] = {nnz_jac, nnz_hess_lower}.
int ref_fg_eval(void *h, int N, const double *state6, const double *coeffs, int ncoef,
                const double *x, const double *lambda, double sigma,
                double *f, double *grad, double *g, double *jac_dense, double *hess_dense, int *nnz)
{
    MPC *mpc = static_cast<MPC *>(h);
    const int n = 8 * N - 2, m = 6 * N;
    StandinProbe &p = standin_probe();
    p.active = true;
    p.x.assign(x, x + n); p.lambda.assign(lambda, lambda + m); p.sigma = sigma;
    Eigen::VectorXd st(6), co(ncoef);
    for (int i = 0; i < 6; i++) st[i] = state6[i];
    for (int i = 0; i < ncoef; i++) co[i] = coeffs[i];
    mpc->Solve(st, co);
    p.active = false;
    if (p.n != n || p.m != m) return -1;
    *f = p.f;
    std::memcpy(grad, p.grad.data(), sizeof(double) * n);
    std::memcpy(g, p.g.data(), sizeof(double) * m);
    std::memset(jac_dense, 0, sizeof(double) * (size_t)m * n);
    for (size_t k = 0; k < p.jac_vals.size(); k++) jac_dense[(size_t)p.jac_row[k] * n + p.jac_col[k]] += p.jac_vals[k];
    std::memset(hess_dense, 0, sizeof(double) * (size_t)n * n);
    for (size_t k = 0; k < p.hess_vals.size(); k++) {
        int i = p.hess_row[k], j = p.hess_col[k];
        hess_dense[(size_t)i * n + j] += p.hess_vals[k];
        if (i != j) hess_dense[(size_t)j * n + i] += p.hess_vals[k];
    }
    nnz[0] = (int)p.jac_vals.size(); nnz[1] = (int)p.hess_vals.size();
    return 0;
}

}  // extern "C"
