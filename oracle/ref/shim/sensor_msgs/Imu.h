#pragma once
#include <memory>
#include <ros/ros.h>
#include <geometry_msgs/types.h>
namespace sensor_msgs {
struct Imu {
    std_msgs::Header header;
    geometry_msgs::Quaternion orientation;
    geometry_msgs::Vector3 angular_velocity;
    geometry_msgs::Vector3 linear_acceleration;
    typedef std::shared_ptr<Imu> Ptr;
    typedef std::shared_ptr<const Imu> ConstPtr;
};
typedef std::shared_ptr<Imu> ImuPtr;
typedef std::shared_ptr<const Imu> ImuConstPtr;
}  // namespace sensor_msgs
