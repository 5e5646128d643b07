// mpc_planner.cpp -- the MPC adapter class over the C ABI (see ../include/mpc_planner.h).
#include "mpc_planner.h"
#include "mpc_b200.h"

#include <cstdlib>
#include <iostream>

MPC::MPC() : handle_(nullptr), device_(0), handle_steps_(0), dirty_(true), status_(0), iters_(0), obj_(0.0), kkt_(0.0)
{
    std::cout << "init mpc" << std::endl;     // the reference announces itself the same way
    if (const char *d = std::getenv("MPC_B200_DEVICE")) device_ = std::atoi(d);
}

MPC::~MPC()
{
    if (handle_) mpc_b200_destroy(handle_);
}

MPC::MPC(const MPC &o)
    : mpc_x(o.mpc_x), mpc_y(o.mpc_y), mpc_theta(o.mpc_theta), params_(o.params_), handle_(nullptr), device_(o.device_),
      handle_steps_(0), dirty_(true), status_(o.status_), iters_(o.iters_), obj_(o.obj_), kkt_(o.kkt_), pred_(o.pred_)
{
}

MPC &MPC::operator=(const MPC &o)
{
    if (this == &o) return *this;
    mpc_x = o.mpc_x; mpc_y = o.mpc_y; mpc_theta = o.mpc_theta;
    params_ = o.params_;
    if (handle_ && device_ != o.device_) { mpc_b200_destroy(handle_); handle_ = nullptr; handle_steps_ = 0; }
    device_ = o.device_;
    dirty_ = true;                 // the own handle (if any) is re-parameterised on the next Solve
    status_ = o.status_; iters_ = o.iters_; obj_ = o.obj_; kkt_ = o.kkt_; pred_ = o.pred_;
    return *this;
}

void MPC::LoadParams(const std::map<std::string, double> &params)
{
    // merge: the reference re-reads every key on each call and keeps the old value of a key
    // that is absent (mpc_planner.cpp:73-85, :247-250)
    for (std::map<std::string, double>::const_iterator it = params.begin(); it != params.end(); ++it)
        params_[it->first] = it->second;
    dirty_ = true;
}

void MPC::ensure_handle()
{
    if (handle_ && !dirty_) return;
    mpc_b200_params p;
    mpc_b200_params_default(&p);               // MPC::MPC / FG_eval::FG_eval defaults
    for (std::map<std::string, double>::const_iterator it = params_.begin(); it != params_.end(); ++it)
        mpc_b200_params_set(&p, it->first.c_str(), it->second);   // unknown keys are ignored, as in the reference
    int rc;
    if (!handle_) {
        rc = mpc_b200_create(&handle_, &p, 1, device_);
    } else {
        rc = mpc_b200_set_params(handle_, &p);
    }
    if (rc != MPC_B200_OK) {
        std::cerr << "[mpc_b200] cannot configure the GPU solver: " << mpc_b200_strerror(rc);
        if (handle_) std::cerr << " (" << mpc_b200_last_cuda_error(handle_) << ")";
        std::cerr << std::endl;
        if (rc == MPC_B200_ERR_CUDA || !handle_) {
            // no CPU path: without a device there is nothing to fall back to
            std::cerr << "[mpc_b200] no usable CUDA device; aborting" << std::endl;
            std::abort();
        }
    }
    handle_steps_ = p.mpc_steps;
    dirty_ = false;
}

std::vector<double> MPC::Solve(Eigen::VectorXd state, Eigen::VectorXd coeffs)
{
    ensure_handle();
    const int N = handle_steps_;
    double st[6] = { 0, 0, 0, 0, 0, 0 };
    for (int i = 0; i < 6 && i < state.size(); i++) st[i] = state[i];
    // the only caller fits a cubic (driving_state.cpp:210); lower orders are padded with zeros, orders 4..7 go through
    // the library's higher-order instantiation (option "poly_coeffs"), anything beyond is refused loudly rather than
    // solved as a different problem
    double co[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
    int nco = 4;
    for (int i = 0; i < coeffs.size(); i++) {
        if (i < 8) { co[i] = coeffs[i]; if (i >= 4 && coeffs[i] != 0.0) nco = i + 1; }
        else if (coeffs[i] != 0.0) {
            std::cerr << "[mpc_b200] fatal: path polynomial of order " << coeffs.size() - 1
                      << " (non-zero coefficient " << i << "); the GPU path takes orders up to 7" << std::endl;
            std::abort();
        }
    }
    if (nco > 4 && mpc_b200_set_option(handle_, "poly_coeffs", (double)nco) != MPC_B200_OK) std::abort();
    double u0[2] = { 0.0, 0.0 };
    pred_.assign(3 * (size_t)N, 0.0);
    int32_t status = 0, iters = 0;
    double obj = 0.0, kkt = 0.0;
    const int rc = mpc_b200_solve_batch(handle_, 1, st, co, nullptr, nullptr, u0, pred_.data(), &obj, &status, &iters,
                                        &kkt, nullptr, nullptr);
    if (nco > 4) mpc_b200_set_option(handle_, "poly_coeffs", 4.0);
    if (rc != MPC_B200_OK)
        std::cerr << "[mpc_b200] solve failed: " << mpc_b200_strerror(rc) << " (" << mpc_b200_last_cuda_error(handle_)
                  << ")" << std::endl;
    status_ = status; iters_ = iters; obj_ = obj; kkt_ = kkt;
    mpc_x.assign(pred_.begin(), pred_.begin() + N);
    mpc_y.assign(pred_.begin() + N, pred_.begin() + 2 * N);
    mpc_theta.assign(pred_.begin() + 2 * N, pred_.begin() + 3 * N);
    std::vector<double> result;
    result.push_back(u0[0]);
    result.push_back(u0[1]);
    return result;
}
