// TEST INFRASTRUCTURE (see ros/ros.h in this directory).
#ifndef ORACLE_STUB_STD_MSGS_HEADER_H
#define ORACLE_STUB_STD_MSGS_HEADER_H
#include <string>
#include "ros/ros.h"
namespace std_msgs { struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; }; }
#endif
