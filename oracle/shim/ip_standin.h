// oracle/shim/ip_standin.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Side channel between the stand-in IpoptApplication and oracle/ref_driver.cpp.
#ifndef ORACLE_IP_STANDIN_H
#define ORACLE_IP_STANDIN_H
#include <vector>

struct StandinProbe {
    bool active;                 // when set, OptimizeTNLP evaluates the callbacks at (x, lambda, sigma) and returns
    std::vector<double> x, lambda;
    double sigma;
    // outputs
    int n, m;
    double f;
    std::vector<double> grad, g, jac_vals, hess_vals;
    std::vector<int> jac_row, jac_col, hess_row, hess_col;
    StandinProbe() : active(false), sigma(1.0), n(0), m(0), f(0.0) {}
};

struct StandinLast {
    int status, iters, n_inertia, n_resto, n_fact;
    double obj, kkt_error, dual_inf, constr_viol, compl_inf, mu;
    std::vector<double> x, lambda, zl, zu;
    StandinLast() : status(0), iters(0), n_inertia(0), n_resto(0), n_fact(0), obj(0), kkt_error(0),
                    dual_inf(0), constr_viol(0), compl_inf(0), mu(0) {}
};

StandinProbe &standin_probe();
StandinLast &standin_last();
// Overrides applied on top of the options the caller set (tests use this to lift max_cpu_time).
void standin_set_cpu_time_override(double seconds /* <=0: none */);
void standin_set_dense_ldl(int on);

#endif
