// Stand-in for PCL's point types (test infrastructure; see ../Eigen/Dense for the rationale).
#pragma once
#include <cstdint>
#ifndef DEG2RAD
#define DEG2RAD(x) ((x)*0.017453293)  // pcl/pcl_macros.h
#endif
#ifndef RAD2DEG
#define RAD2DEG(x) ((x)*57.29578)
#endif
namespace pcl {
struct PointXYZINormal {
    float x = 0, y = 0, z = 0;
    float intensity = 0;
    float normal_x = 0, normal_y = 0, normal_z = 0;
    float curvature = 0;
};
struct PointXYZI {
    float x = 0, y = 0, z = 0, intensity = 0;
};
}  // namespace pcl
