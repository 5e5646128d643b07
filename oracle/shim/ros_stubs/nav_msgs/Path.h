// TEST INFRASTRUCTURE (see ros/ros.h in this directory).
#ifndef ORACLE_STUB_NAV_MSGS_PATH_H
#define ORACLE_STUB_NAV_MSGS_PATH_H
#include <vector>
#include "geometry_msgs/PoseStamped.h"
namespace nav_msgs { struct Path { std_msgs::Header header; std::vector<geometry_msgs::PoseStamped> poses; }; }
#endif
