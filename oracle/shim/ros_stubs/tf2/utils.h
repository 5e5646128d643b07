// TEST INFRASTRUCTURE (see ros/ros.h in this directory).
#ifndef ORACLE_STUB_TF2_UTILS_H
#define ORACLE_STUB_TF2_UTILS_H
#include <cmath>
#include "geometry_msgs/PoseStamped.h"
namespace tf2 {
class Quaternion {
public:
    Quaternion() { q_[0] = q_[1] = q_[2] = 0.0; q_[3] = 1.0; }
    void setRPY(double roll, double pitch, double yaw)
    {
        const double cy = std::cos(0.5 * yaw), sy = std::sin(0.5 * yaw), cp = std::cos(0.5 * pitch), sp = std::sin(0.5 * pitch),
                     cr = std::cos(0.5 * roll), sr = std::sin(0.5 * roll);
        q_[0] = sr * cp * cy - cr * sp * sy; q_[1] = cr * sp * cy + sr * cp * sy;
        q_[2] = cr * cp * sy - sr * sp * cy; q_[3] = cr * cp * cy + sr * sp * sy;
    }
    double operator[](int i) const { return q_[i]; }
private:
    double q_[4];
};
inline double getYaw(const geometry_msgs::Quaternion &q)
{
    return std::atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z));
}
}
#endif
