// TEST INFRASTRUCTURE (see ros/ros.h in this directory).
#ifndef ORACLE_STUB_ANGLES_H
#define ORACLE_STUB_ANGLES_H
#include <cmath>
namespace angles {
inline double normalize_angle_positive(double a) { const double r = std::fmod(a, 2.0 * M_PI); return r < 0.0 ? r + 2.0 * M_PI : r; }
inline double normalize_angle(double a) { const double r = normalize_angle_positive(a); return r > M_PI ? r - 2.0 * M_PI : r; }
inline double shortest_angular_distance(double from, double to) { return normalize_angle(to - from); }
}
#endif
