// oracle/shim/coin/IpTNLP.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Stand-in for the part of Ipopt 3.12's public C++ interface that the
// reference's vendored glue uses (mpc_ros/include/cppad/ipopt/solve.hpp:445-586,
// solve_callback.hpp:42-57,461-1186).  Written from the published TNLP
// interface; the solver behind it is oracle/ipm.c.
#ifndef ORACLE_SHIM_IPTNLP_HPP
#define ORACLE_SHIM_IPTNLP_HPP
#include <cstddef>

namespace Ipopt {

typedef double Number;
typedef int Index;

class ReferencedObject {
public:
    ReferencedObject() : refs_(0) {}
    virtual ~ReferencedObject() {}
    void AddRef() const { ++refs_; }
    int ReleaseRef() const { return --refs_; }
private:
    mutable int refs_;
};

template <class T>
class SmartPtr {
public:
    SmartPtr() : p_(NULL) {}
    SmartPtr(T *p) : p_(p) { if (p_) p_->AddRef(); }
    SmartPtr(const SmartPtr<T> &o) : p_(o.p_) { if (p_) p_->AddRef(); }
    template <class U>
    SmartPtr(const SmartPtr<U> &o) : p_(o.get()) { if (p_) p_->AddRef(); }
    ~SmartPtr() { release(); }
    SmartPtr<T> &operator=(const SmartPtr<T> &o)
    {
        if (o.p_) o.p_->AddRef();
        release();
        p_ = o.p_;
        return *this;
    }
    T *operator->() const { return p_; }
    T &operator*() const { return *p_; }
    T *get() const { return p_; }
private:
    void release() { if (p_ && p_->ReleaseRef() == 0) delete p_; p_ = NULL; }
    T *p_;
};

template <class T> inline T *GetRawPtr(const SmartPtr<T> &p) { return p.get(); }
template <class T> inline bool IsValid(const SmartPtr<T> &p) { return p.get() != NULL; }

enum SolverReturn {
    SUCCESS,
    MAXITER_EXCEEDED,
    CPUTIME_EXCEEDED,
    STOP_AT_TINY_STEP,
    STOP_AT_ACCEPTABLE_POINT,
    LOCAL_INFEASIBILITY,
    USER_REQUESTED_STOP,
    FEASIBLE_POINT_FOUND,
    DIVERGING_ITERATES,
    RESTORATION_FAILURE,
    ERROR_IN_STEP_COMPUTATION,
    INVALID_NUMBER_DETECTED,
    TOO_FEW_DEGREES_OF_FREEDOM,
    INVALID_OPTION,
    OUT_OF_MEMORY,
    INTERNAL_ERROR,
    UNASSIGNED
};

class IpoptData;
class IpoptCalculatedQuantities;

class TNLP : public ReferencedObject {
public:
    enum IndexStyleEnum { C_STYLE = 0, FORTRAN_STYLE = 1 };
    virtual ~TNLP() {}
    virtual bool get_nlp_info(Index &n, Index &m, Index &nnz_jac_g, Index &nnz_h_lag,
                              IndexStyleEnum &index_style) = 0;
    virtual bool get_bounds_info(Index n, Number *x_l, Number *x_u, Index m, Number *g_l, Number *g_u) = 0;
    virtual bool get_starting_point(Index n, bool init_x, Number *x, bool init_z, Number *z_L, Number *z_U,
                                    Index m, bool init_lambda, Number *lambda) = 0;
    virtual bool eval_f(Index n, const Number *x, bool new_x, Number &obj_value) = 0;
    virtual bool eval_grad_f(Index n, const Number *x, bool new_x, Number *grad_f) = 0;
    virtual bool eval_g(Index n, const Number *x, bool new_x, Index m, Number *g) = 0;
    virtual bool eval_jac_g(Index n, const Number *x, bool new_x, Index m, Index nele_jac, Index *iRow,
                            Index *jCol, Number *values) = 0;
    virtual bool eval_h(Index n, const Number *x, bool new_x, Number obj_factor, Index m,
                        const Number *lambda, bool new_lambda, Index nele_hess, Index *iRow, Index *jCol,
                        Number *values) = 0;
    virtual void finalize_solution(SolverReturn status, Index n, const Number *x, const Number *z_L,
                                   const Number *z_U, Index m, const Number *g, const Number *lambda,
                                   Number obj_value, const IpoptData *ip_data,
                                   IpoptCalculatedQuantities *ip_cq) = 0;
};

}  // namespace Ipopt
#endif
