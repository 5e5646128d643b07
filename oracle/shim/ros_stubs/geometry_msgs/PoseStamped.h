// TEST INFRASTRUCTURE (see ros/ros.h in this directory).
#ifndef ORACLE_STUB_GEOMETRY_MSGS_POSESTAMPED_H
#define ORACLE_STUB_GEOMETRY_MSGS_POSESTAMPED_H
#include "std_msgs/Header.h"
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
}
#endif
