// TEST INFRASTRUCTURE (see ros/ros.h in this directory).
#ifndef ORACLE_STUB_TF2_ROS_BUFFER_H
#define ORACLE_STUB_TF2_ROS_BUFFER_H
namespace tf2_ros { class Buffer {}; }
#endif
