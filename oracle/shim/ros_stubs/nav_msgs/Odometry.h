// TEST INFRASTRUCTURE (see ros/ros.h in this directory).
#ifndef ORACLE_STUB_NAV_MSGS_ODOMETRY_H
#define ORACLE_STUB_NAV_MSGS_ODOMETRY_H
#include "geometry_msgs/PoseStamped.h"
#include "geometry_msgs/Twist.h"
namespace nav_msgs { struct Odometry { std_msgs::Header header; geometry_msgs::Pose pose; geometry_msgs::Twist twist; }; }
#endif
