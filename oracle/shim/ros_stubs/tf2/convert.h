// TEST INFRASTRUCTURE (see ros/ros.h in this directory).
#ifndef ORACLE_STUB_TF2_CONVERT_H
#define ORACLE_STUB_TF2_CONVERT_H
#include "tf2/utils.h"
#endif
