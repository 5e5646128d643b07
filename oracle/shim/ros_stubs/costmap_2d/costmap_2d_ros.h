// TEST INFRASTRUCTURE (see ros/ros.h in this directory).  The costmap is reduced to "where is the robot".
#ifndef ORACLE_STUB_COSTMAP_2D_ROS_H
#define ORACLE_STUB_COSTMAP_2D_ROS_H
#include <string>
#include "geometry_msgs/PoseStamped.h"
namespace costmap_2d {
class Costmap2D {};
class Costmap2DROS {
public:
    Costmap2D *getCostmap() { return &map_; }
    std::string getGlobalFrameID() { return "odom"; }
    std::string getBaseFrameID() { return "base_link"; }
    bool getRobotPose(geometry_msgs::PoseStamped &p) const { p = pose; return true; }
    geometry_msgs::PoseStamped pose;
private:
    Costmap2D map_;
};
}
#endif
