// TEST INFRASTRUCTURE (see ros/ros.h in this directory).
#ifndef ORACLE_STUB_NAV_CORE_BLP_H
#define ORACLE_STUB_NAV_CORE_BLP_H
#include <string>
#include <vector>
#include "geometry_msgs/PoseStamped.h"
#include "geometry_msgs/Twist.h"
namespace tf2_ros { class Buffer; }
namespace costmap_2d { class Costmap2DROS; }
namespace nav_core {
class BaseLocalPlanner {
public:
    virtual bool computeVelocityCommands(geometry_msgs::Twist &cmd_vel) = 0;
    virtual bool isGoalReached() = 0;
    virtual bool setPlan(const std::vector<geometry_msgs::PoseStamped> &plan) = 0;
    virtual void initialize(std::string name, tf2_ros::Buffer *tf, costmap_2d::Costmap2DROS *costmap_ros) = 0;
    virtual ~BaseLocalPlanner() {}
protected:
    BaseLocalPlanner() {}
};
}
#endif
