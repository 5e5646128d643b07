// oracle/shim/ip_standin.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// IpoptApplication::OptimizeTNLP for the stand-in headers: adapts the
// Ipopt::TNLP the reference's CppAD glue builds
// (mpc_ros/include/cppad/ipopt/solve_callback.hpp) to oracle/ipm.c and hands
// the result back through finalize_solution (solve_callback.hpp:1091-1186).
#include "coin/IpIpoptApplication.hpp"
#include "ip_standin.h"
#include "../ipm.h"
#include <cstring>

static thread_local StandinProbe g_probe;
static thread_local StandinLast g_last;
static thread_local double g_cpu_override = 0.0;
static thread_local int g_dense = 0;

StandinProbe &standin_probe() { return g_probe; }
StandinLast &standin_last() { return g_last; }
void standin_set_cpu_time_override(double s) { g_cpu_override = s; }
void standin_set_dense_ldl(int on) { g_dense = on; }

namespace {
struct Ctx {
    Ipopt::TNLP *t;
    int n, m, nj, nh;
};
int cb_bounds(void *u, double *xl, double *xu, double *gl, double *gu)
{ Ctx *c = (Ctx *)u; return c->t->get_bounds_info(c->n, xl, xu, c->m, gl, gu); }
int cb_start(void *u, double *x0)
{ Ctx *c = (Ctx *)u; return c->t->get_starting_point(c->n, true, x0, false, NULL, NULL, c->m, false, NULL); }
int cb_f(void *u, const double *x, double *f)
{ Ctx *c = (Ctx *)u; Ipopt::Number v = 0; bool ok = c->t->eval_f(c->n, x, true, v); *f = v; return ok; }
int cb_grad(void *u, const double *x, double *g)
{ Ctx *c = (Ctx *)u; return c->t->eval_grad_f(c->n, x, true, g); }
int cb_g(void *u, const double *x, double *g)
{ Ctx *c = (Ctx *)u; return c->t->eval_g(c->n, x, true, c->m, g); }
int cb_js(void *u, int *ir, int *jc)
{ Ctx *c = (Ctx *)u; return c->t->eval_jac_g(c->n, NULL, false, c->m, c->nj, ir, jc, NULL); }
int cb_jac(void *u, const double *x, double *v)
{ Ctx *c = (Ctx *)u; return c->t->eval_jac_g(c->n, x, true, c->m, c->nj, NULL, NULL, v); }
int cb_hs(void *u, int *ir, int *jc)
{ Ctx *c = (Ctx *)u; return c->t->eval_h(c->n, NULL, false, 1.0, c->m, NULL, false, c->nh, ir, jc, NULL); }
int cb_hess(void *u, const double *x, double sigma, const double *lam, double *v)
{ Ctx *c = (Ctx *)u; return c->t->eval_h(c->n, x, true, sigma, c->m, lam, true, c->nh, NULL, NULL, v); }
}  // namespace

namespace Ipopt {

ApplicationReturnStatus IpoptApplication::OptimizeTNLP(const SmartPtr<TNLP> &tnlp)
{
    Ctx c;
    c.t = tnlp.get();
    TNLP::IndexStyleEnum style;
    Index n, m, nj, nh;
    if (!c.t->get_nlp_info(n, m, nj, nh, style)) return Invalid_Problem_Definition;
    c.n = n; c.m = m; c.nj = nj; c.nh = nh;

    if (g_probe.active) {
        StandinProbe &p = g_probe;
        p.n = n; p.m = m;
        p.grad.assign(n, 0.0); p.g.assign(m, 0.0);
        p.jac_vals.assign(nj, 0.0); p.hess_vals.assign(nh, 0.0);
        p.jac_row.assign(nj, 0); p.jac_col.assign(nj, 0); p.hess_row.assign(nh, 0); p.hess_col.assign(nh, 0);
        cb_js(&c, p.jac_row.data(), p.jac_col.data());
        cb_hs(&c, p.hess_row.data(), p.hess_col.data());
        cb_f(&c, p.x.data(), &p.f);
        cb_grad(&c, p.x.data(), p.grad.data());
        cb_g(&c, p.x.data(), p.g.data());
        cb_jac(&c, p.x.data(), p.jac_vals.data());
        cb_hess(&c, p.x.data(), p.sigma, p.lambda.data(), p.hess_vals.data());
        std::vector<double> z(n, 0.0);
        c.t->finalize_solution(USER_REQUESTED_STOP, n, p.x.data(), z.data(), z.data(), m, p.g.data(),
                               p.lambda.data(), p.f, NULL, NULL);
        return User_Requested_Stop;
    }

    ipm_nlp nlp;
    nlp.n = n; nlp.m = m; nlp.nnz_jac = nj; nlp.nnz_hess = nh; nlp.user = &c;
    nlp.get_bounds = cb_bounds; nlp.get_start = cb_start; nlp.eval_f = cb_f; nlp.eval_grad_f = cb_grad;
    nlp.eval_g = cb_g; nlp.jac_struct = cb_js; nlp.eval_jac = cb_jac; nlp.hess_struct = cb_hs; nlp.eval_hess = cb_hess;

    ipm_options opt;
    ipm_default_options(&opt);
    Number v; Index iv;
    if (options_->GetNumericValue("tol", v)) opt.tol = v;
    if (options_->GetNumericValue("max_cpu_time", v)) opt.max_cpu_time = v;
    if (options_->GetNumericValue("acceptable_tol", v)) opt.acceptable_tol = v;
    if (options_->GetIntegerValue("max_iter", iv)) opt.max_iter = iv;
    if (options_->GetIntegerValue("print_level", iv)) opt.print_level = iv > 4 ? 1 : 0;
    if (g_cpu_override > 0.0) opt.max_cpu_time = g_cpu_override;
    opt.use_dense_ldl = g_dense;

    StandinLast &L = g_last;
    L.x.assign(n, 0.0); L.lambda.assign(m, 0.0); L.zl.assign(n, 0.0); L.zu.assign(n, 0.0);
    std::vector<double> g(m, 0.0);
    ipm_result r;
    std::memset(&r, 0, sizeof(r));
    r.x = L.x.data(); r.lambda = L.lambda.data(); r.zl = L.zl.data(); r.zu = L.zu.data(); r.g = g.data();
    int st = ipm_solve(&nlp, &opt, &r);
    L.status = st; L.iters = r.iters; L.obj = r.obj; L.kkt_error = r.kkt_error;
    L.dual_inf = r.dual_inf; L.constr_viol = r.constr_viol; L.compl_inf = r.compl_inf; L.mu = r.mu;
    L.n_inertia = r.n_inertia_corrections; L.n_resto = r.n_restorations; L.n_fact = r.n_factorizations;

    SolverReturn sr; ApplicationReturnStatus ar;
    switch (st) {
        case IPM_SUCCESS: sr = SUCCESS; ar = Solve_Succeeded; break;
        case IPM_MAXITER_EXCEEDED: sr = MAXITER_EXCEEDED; ar = Maximum_Iterations_Exceeded; break;
        case IPM_STOP_AT_TINY_STEP: sr = STOP_AT_TINY_STEP; ar = Search_Direction_Becomes_Too_Small; break;
        case IPM_STOP_AT_ACCEPTABLE_POINT: sr = STOP_AT_ACCEPTABLE_POINT; ar = Solved_To_Acceptable_Level; break;
        case IPM_LOCAL_INFEASIBILITY: sr = LOCAL_INFEASIBILITY; ar = Infeasible_Problem_Detected; break;
        case IPM_RESTORATION_FAILURE: sr = RESTORATION_FAILURE; ar = Restoration_Failed; break;
        case IPM_ERROR_IN_STEP_COMPUTATION: sr = ERROR_IN_STEP_COMPUTATION; ar = Error_In_Step_Computation; break;
        case IPM_INVALID_NUMBER_DETECTED: sr = INVALID_NUMBER_DETECTED; ar = Invalid_Number_Detected; break;
        case 15: sr = CPUTIME_EXCEEDED; ar = Maximum_CpuTime_Exceeded; break;
        default: sr = INTERNAL_ERROR; ar = Internal_Error; break;
    }
    c.t->finalize_solution(sr, n, r.x, r.zl, r.zu, m, r.g, r.lambda, r.obj, NULL, NULL);
    return ar;
}

}  // namespace Ipopt
