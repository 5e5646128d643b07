// TEST INFRASTRUCTURE (see ros/ros.h in this directory).
#ifndef ORACLE_STUB_TF2_GEOMETRY_MSGS_H
#define ORACLE_STUB_TF2_GEOMETRY_MSGS_H
#include "tf2/utils.h"
#endif
