// Stand-in for pcl::VoxelGrid. The harness hands KILO::process a cloud that is ALREADY downsampled (the oracle's
// preprocess step is a separate row of SURVEY §8), so filter() is the identity here.
#pragma once
#include <pcl/point_cloud.h>
namespace pcl {
template <class PointT>
class VoxelGrid {
   public:
    void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
    void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { in_ = c; }
    void filter(PointCloud<PointT>& out) { out = *in_; }
   private:
    float leaf_[3] = {0, 0, 0};
    typename PointCloud<PointT>::ConstPtr in_;
};
}  // namespace pcl
