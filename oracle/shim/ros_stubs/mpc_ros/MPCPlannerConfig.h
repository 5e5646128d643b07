// TEST INFRASTRUCTURE (see ros/ros.h in this directory).  What dynamic_reconfigure generates from the reference's
// mpc_ros/cfg/MPCPlanner.cfg:13-39 (+ base_local_planner's generic local-planner limits, :11): names and defaults.
#ifndef ORACLE_STUB_MPCPLANNERCONFIG_H
#define ORACLE_STUB_MPCPLANNERCONFIG_H
namespace mpc_ros {
struct MPCPlannerConfig {
    // add_generic_localplanner_params (base_local_planner/cfg/LocalPlannerLimits)
    double max_vel_trans = 0.55, min_vel_trans = 0.1, max_vel_x = 0.55, min_vel_x = 0.0, max_vel_y = 0.1, min_vel_y = -0.1,
           max_vel_theta = 1.0, min_vel_theta = 0.4, acc_lim_x = 2.5, acc_lim_y = 2.5, acc_lim_theta = 3.2, acc_lim_trans = 0.1,
           xy_goal_tolerance = 0.1, yaw_goal_tolerance = 0.1, trans_stopped_vel = 0.1, theta_stopped_vel = 0.1;
    bool prune_plan = false, restore_defaults = false;
    // MPCPlanner.cfg:13-39
    bool debug_info = true, delay_mode = true;
    double max_speed = 0.50, default_max_speed = 0.50, waypoints_dist = -1, path_length = 5.0, controller_freq = 10.0;
    double steps = 20.0, ref_cte = 0.0, ref_vel = 1.0, ref_etheta = 0.0;
    double w_cte = 1000.0, w_etheta = 1000.0, w_vel = 100.0, w_angvel = 100.0, w_angvel_d = 0.0, w_accel = 50.0, w_accel_d = 10.0;
    double max_angvel = 1.0, max_throttle = 1.0, bound_value = 1000.0;
    double heading_yaw_error_threshold = 0.1;
    static MPCPlannerConfig __getDefault__() { return MPCPlannerConfig(); }
};
}
#endif
