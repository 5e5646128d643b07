// TEST INFRASTRUCTURE (see ros/ros.h in this directory).
#ifndef ORACLE_STUB_STD_MSGS_STRING_H
#define ORACLE_STUB_STD_MSGS_STRING_H
#include <string>
namespace std_msgs { struct String { std::string data; }; }
#endif
