// TEST INFRASTRUCTURE (see ros/ros.h in this directory).  base_local_planner::LocalPlannerUtil reduced to a plan holder:
// getLocalPlan returns the plan set by setPlan unchanged (no tf, no costmap clipping, no pruning).
#ifndef ORACLE_STUB_BLP_UTIL_H
#define ORACLE_STUB_BLP_UTIL_H
#include <string>
#include <vector>
#include "geometry_msgs/PoseStamped.h"
namespace tf2_ros { class Buffer; }
namespace costmap_2d { class Costmap2D; }
namespace base_local_planner {
struct LocalPlannerLimits {
    double max_vel_trans = 0, min_vel_trans = 0, max_vel_x = 0, min_vel_x = 0, max_vel_y = 0, min_vel_y = 0,
           max_vel_theta = 0, min_vel_theta = 0, acc_lim_x = 0, acc_lim_y = 0, acc_lim_theta = 0, acc_lim_trans = 0,
           xy_goal_tolerance = 0.1, yaw_goal_tolerance = 0.1, trans_stopped_vel = 0.1, theta_stopped_vel = 0.1;
    bool prune_plan = false, restore_defaults = false;
};
class LocalPlannerUtil {
public:
    void initialize(tf2_ros::Buffer *, costmap_2d::Costmap2D *, std::string frame) { frame_ = frame; }
    void reconfigureCB(LocalPlannerLimits &l, bool) { limits_ = l; }
    bool setPlan(const std::vector<geometry_msgs::PoseStamped> &p) { plan_ = p; return true; }
    bool getGoal(geometry_msgs::PoseStamped &g) { if (plan_.empty()) return false; g = plan_.back(); return true; }
    bool getLocalPlan(const geometry_msgs::PoseStamped &, std::vector<geometry_msgs::PoseStamped> &out) { out = plan_; return !plan_.empty(); }
    LocalPlannerLimits getCurrentLimits() { return limits_; }
private:
    std::string frame_;
    std::vector<geometry_msgs::PoseStamped> plan_;
    LocalPlannerLimits limits_;
};
}
#endif
