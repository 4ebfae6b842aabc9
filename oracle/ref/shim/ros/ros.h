// Stand-in for roscpp: only the types voxel_map.h / KILO.h name (publisher, rate, time).
#pragma once
#include <cstdint>
#include <string>
namespace ros {
struct Time {
    double t = 0.0;
    Time() {}
    explicit Time(double s) : t(s) {}
    double toSec() const { return t; }
    Time& fromSec(double s) { t = s; return *this; }
};
struct Duration {
    double d = 0.0;
    Duration() {}
    explicit Duration(double s) : d(s) {}
};
struct Rate {
    explicit Rate(double) {}
    void sleep() {}
};
struct Publisher {
    template <class M>
    void publish(const M&) const {}
};
struct NodeHandle {};
}  // namespace ros
namespace std_msgs {
struct Header {
    uint32_t seq = 0;
    ros::Time stamp;
    std::string frame_id;
};
}  // namespace std_msgs
