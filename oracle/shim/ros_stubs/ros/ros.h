// oracle/shim/ros_stubs -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Stand-in for the slice of roscpp the reference's ROS sources touch (mpc_ros/src/mpc_planner_ros.cpp,
// mpc_ros/src/driving_state.cpp): ROS is absent from this image.  Lets the UNMODIFIED reference sources compile
// (and the parts that need no live ROS graph run) against this repository's MPC class or the reference's own.
#ifndef ORACLE_STUB_ROS_H
#define ORACLE_STUB_ROS_H
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <functional>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#define ROS_WARN(...) do { if (::ros::stub_verbose()) { std::fprintf(stderr, "[WARN] " __VA_ARGS__); std::fprintf(stderr, "\n"); } } while (0)
#define ROS_ERROR(...) do { if (::ros::stub_verbose()) { std::fprintf(stderr, "[ERROR] " __VA_ARGS__); std::fprintf(stderr, "\n"); } } while (0)
#define ROS_INFO(...) do { if (::ros::stub_verbose()) { std::fprintf(stderr, "[INFO] " __VA_ARGS__); std::fprintf(stderr, "\n"); } } while (0)
namespace boost {
template <class S> using function = std::function<S>;
template <class... A> auto bind(A &&... a) -> decltype(std::bind(std::forward<A>(a)...)) { return std::bind(std::forward<A>(a)...); }
}  // namespace boost
using namespace std::placeholders;   // boost's global _1, _2
namespace ros {
inline bool &stub_verbose() { static bool v = false; return v; }
struct Time {
    double t;
    Time() : t(0.0) {}
    static Time now() { return Time(); }
    double toSec() const { return t; }
};
class Publisher {
public:
    template <class M> void publish(const M &) const { count_++; }
    mutable long count_ = 0;
};
class Subscriber {};
class NodeHandle {
public:
    NodeHandle() {}
    explicit NodeHandle(const std::string &ns) : ns_(ns) {}
    static std::map<std::string, double> &params() { static std::map<std::string, double> p; return p; }
    bool searchParam(const std::string &key, std::string &result) const
    {
        if (params().count(key)) { result = key; return true; }
        return false;
    }
    template <class T> bool param(const std::string &key, T &val, const T &def) const
    {
        std::map<std::string, double>::const_iterator it = params().find(key);
        if (it == params().end()) { val = def; return false; }
        val = static_cast<T>(it->second);
        return true;
    }
    template <class M> Publisher advertise(const std::string &, uint32_t, bool = false) { return Publisher(); }
    template <class M, class T> Subscriber subscribe(const std::string &, uint32_t, void (T::*)(const M &), T *) { return Subscriber(); }
private:
    std::string ns_;
};
}  // namespace ros
#endif
