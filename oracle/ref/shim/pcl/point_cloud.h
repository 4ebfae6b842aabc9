#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl {
template <class PointT>
class PointCloud {
   public:
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void clear() { points.clear(); width = height = 0; }
    void push_back(const PointT& p) { points.push_back(p); }
    PointT& operator[](size_t i) { return points[i]; }
    const PointT& operator[](size_t i) const { return points[i]; }
};
}  // namespace pcl
