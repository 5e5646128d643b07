// TEST INFRASTRUCTURE (see ros/ros.h in this directory).
#ifndef ORACLE_STUB_BLP_GOAL_FUNCTIONS_H
#define ORACLE_STUB_BLP_GOAL_FUNCTIONS_H
#include <vector>
#include "nav_msgs/Path.h"
#include "ros/ros.h"
namespace base_local_planner {
inline void publishPlan(const std::vector<geometry_msgs::PoseStamped> &path, const ros::Publisher &pub)
{
    nav_msgs::Path p; p.poses = path; pub.publish(p);
}
}
#endif
