// TEST INFRASTRUCTURE (see ros/ros.h in this directory).
#ifndef ORACLE_STUB_BLP_TRAJECTORY_H
#define ORACLE_STUB_BLP_TRAJECTORY_H
namespace base_local_planner { class Trajectory { public: double xv_ = 0, yv_ = 0, thetav_ = 0, cost_ = -1; }; }
#endif
