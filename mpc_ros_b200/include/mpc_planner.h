// mpc_planner.h -- drop-in replacement for the reference's MPC class
// (OkDoky/mpc_ros, mpc_ros/include/mpc_planner.h:26-47): same class name, same public members,
// same call semantics, so driving_state.cpp / mpc_planner_ros.cpp compile against it unchanged.
// Behind Solve() sits the B200 solver's C ABI (include/mpc_b200.h) with a batch of one.
//
//   MPC::MPC()                     prints "init mpc" like the reference (mpc_planner.cpp:225)
//   MPC::LoadParams(map)           string keys of mpc_planner.cpp:73-85, :247-250; a missing key keeps
//                                  the previous value
//   MPC::Solve(state, coeffs)      returns {w_0, throttle_0}; fills mpc_x, mpc_y, mpc_theta
//                                  (mpc_planner.cpp:388-401).  Like the reference it never throws and
//                                  returns the last iterate whatever the solver status (:378); the
//                                  status is kept in last_status() for callers that care.
#ifndef MPC_B200_MPC_PLANNER_H
#define MPC_B200_MPC_PLANNER_H

#include <map>
#include <string>
#include <vector>
#include <Eigen/Core>

struct mpc_b200_handle;

// The reference's header does this at mpc_planner.h:24 and code that includes it may lean on it.
using namespace std;

class MPC {
public:
    MPC();
    ~MPC();
    // Copyable like the reference's class (DrivingStateContext::getMpc() returns an MPC by value,
    // driving_state.h:80-82): a copy takes the parameters, the device and the last prediction, and lazily
    // creates its OWN solver handle on its first Solve.
    MPC(const MPC &other);
    MPC &operator=(const MPC &other);

    std::vector<double> Solve(Eigen::VectorXd state, Eigen::VectorXd coeffs);
    std::vector<double> mpc_x;
    std::vector<double> mpc_y;
    std::vector<double> mpc_theta;

    void LoadParams(const std::map<std::string, double> &params);

    // additions (not in the reference): solver diagnostics of the last Solve
    int last_status() const { return status_; }
    int last_iterations() const { return iters_; }
    double last_objective() const { return obj_; }
    double last_kkt_error() const { return kkt_; }
    // CUDA device used by this object (default 0, or $MPC_B200_DEVICE)
    int device() const { return device_; }

private:
    void ensure_handle();
    std::map<std::string, double> params_;
    mpc_b200_handle *handle_;
    int device_;
    int handle_steps_;
    bool dirty_;
    int status_, iters_;
    double obj_, kkt_;
    std::vector<double> pred_;
};

#endif
