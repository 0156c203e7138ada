// ros:: stand-in for the UNCHANGED ros/packages/caffe_ros sources (ROS is not in this image): the logging / assertion macros,
// ros::shutdown() and ros::Time that tensor_net.{h,cpp} and int8_calibrator.cpp use.  Drop-in infrastructure only (tools/dropin).
#pragma once
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <memory>
#include <string>
#include <vector>

#define ROS_STUB_LOG(tag, ...) do { std::fprintf(stderr, "[%s] ", tag); std::fprintf(stderr, __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_DEBUG(...) do { if (std::getenv("ROS_STUB_DEBUG")) ROS_STUB_LOG("DEBUG", __VA_ARGS__); } while (0)
#define ROS_INFO(...) ROS_STUB_LOG("INFO", __VA_ARGS__)
#define ROS_WARN(...) ROS_STUB_LOG("WARN", __VA_ARGS__)
#define ROS_ERROR(...) ROS_STUB_LOG("ERROR", __VA_ARGS__)
#define ROS_FATAL(...) ROS_STUB_LOG("FATAL", __VA_ARGS__)
#define ROS_ASSERT(cond) do { if (!(cond)) { std::fprintf(stderr, "ROS_ASSERT failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); std::abort(); } } while (0)

namespace ros {
// A node that calls ros::shutdown() after ROS_FATAL stops; here the process exits with an error.
inline void shutdown() { std::fflush(stderr); std::exit(3); }
struct Duration {
    double s = 0;
    double toSec() const { return s; }
};
struct Time {
    double t = 0;
    static Time now()
    {
        Time r;
        r.t = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
        return r;
    }
    Duration operator-(const Time& o) const { Duration d; d.s = t - o.t; return d; }
};
}  // namespace ros
