// TEST INFRASTRUCTURE (see ros/ros.h in this directory).  setCallback fires the callback once with the
// schema's defaults, as the real server does on start-up.
#ifndef ORACLE_STUB_DYNRECONF_SERVER_H
#define ORACLE_STUB_DYNRECONF_SERVER_H
#include "ros/ros.h"
namespace dynamic_reconfigure {
template <class C> class Server {
public:
    typedef boost::function<void(C &, uint32_t)> CallbackType;
    explicit Server(const ros::NodeHandle &) {}
    void setCallback(const CallbackType &cb) { cb_ = cb; C c = C::__getDefault__(); cb_(c, ~0u); }
    void update(C &c) { if (cb_) cb_(c, 0u); }
private:
    CallbackType cb_;
};
}
#endif
