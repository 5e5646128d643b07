// TEST INFRASTRUCTURE (see ros/ros.h in this directory).  The real macro registers a factory with class_loader; here it
// only has to prove that the class is concrete and derives from the base.
#ifndef ORACLE_STUB_PLUGINLIB_H
#define ORACLE_STUB_PLUGINLIB_H
#define PLUGINLIB_EXPORT_CLASS(cls, base) \
    extern "C" base *oracle_stub_make_plugin() { return new cls(); }
#endif
