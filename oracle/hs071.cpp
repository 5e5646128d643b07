// oracle/hs071.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// The Ipopt-manual problem HS071 driven through CppAD::ipopt::solve and the
// stand-in solver, the same way the reference's shipped example does
// (assets/document/example/CppAD_Ipopt.cpp:63-165): known answer
// x = (1, 4.743, 3.82115, 1.379408), zl[0] = 1.087871 (:146-147).
// Restated here, not copied: min x1 x4 (x1+x2+x3) + x3  s.t.  x1 x2 x3 x4 >= 25,
// sum x_i^2 = 40, 1 <= x <= 5, start (1,5,5,1).
#include <cppad/ipopt/solve.hpp>
#include "shim/ip_standin.h"

namespace {
struct Hs071 {
    typedef CPPAD_TESTVECTOR(CppAD::AD<double>) ADvector;
    void operator()(ADvector &fg, const ADvector &x)
    {
        fg[0] = x[0] * x[3] * (x[0] + x[1] + x[2]) + x[2];
        fg[1] = x[0] * x[1] * x[2] * x[3];
        CppAD::AD<double> s = 0.0;
        for (int i = 0; i < 4; i++) s += x[i] * x[i];
        fg[2] = s;
    }
};
}  // namespace

extern "C" int ref_hs071(double tol, double *x4, double *zl4, double *zu4, double *obj, int *iters)
{
    typedef CPPAD_TESTVECTOR(double) Dvector;
    Dvector xi(4), xl(4), xu(4), gl(2), gu(2);
    xi[0] = 1.0; xi[1] = 5.0; xi[2] = 5.0; xi[3] = 1.0;
    for (int i = 0; i < 4; i++) { xl[i] = 1.0; xu[i] = 5.0; }
    gl[0] = 25.0; gu[0] = 1.0e19;
    gl[1] = 40.0; gu[1] = 40.0;
    Hs071 fg;
    std::string options;
    options += "Integer print_level  0\n";
    options += "Sparse  true         reverse\n";
    char buf[64];
    snprintf(buf, sizeof(buf), "Numeric tol          %g\n", tol);
    options += buf;
    CppAD::ipopt::solve_result<Dvector> sol;
    standin_probe().active = false;
    CppAD::ipopt::solve<Dvector, Hs071>(options, xi, xl, xu, gl, gu, fg, sol);
    for (int i = 0; i < 4; i++) { x4[i] = sol.x[i]; zl4[i] = sol.zl[i]; zu4[i] = sol.zu[i]; }
    *obj = sol.obj_value;
    *iters = standin_last().iters;
    return (int)sol.status;
}
