// TEST INFRASTRUCTURE (see ros/ros.h in this directory).
#ifndef ORACLE_STUB_GEOMETRY_MSGS_TWIST_H
#define ORACLE_STUB_GEOMETRY_MSGS_TWIST_H
namespace geometry_msgs {
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Twist { Vector3 linear, angular; };
}
#endif
