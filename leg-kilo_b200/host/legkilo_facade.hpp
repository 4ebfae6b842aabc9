// legkilo_facade.hpp — the reference's class names over the C ABI.
//
// Header-only, needs Eigen, i.e. it is compiled INSIDE the reference's catkin workspace, not in this
// repository's build image (no Eigen / PCL / ROS here): everything below is guarded by __has_include.
// tests/test_facade_compiles.py type-checks it and instantiates its templates against a small stand-in
// for <Eigen/Dense> (tests/stubs/); the POD C ABI underneath (include/legkilo_b200.h) is what the parity
// tests cover.
//
// Mirrors: legkilo::State / ESKF (legkilo/src/core/slam/eskf.h:15-109), VoxelMapManager
// (legkilo/src/core/slam/voxel_map.h:180-244), and the per-scan entry KILO::process
// (legkilo/src/core/slam/KILO.h:28).
#pragma once
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>

#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "legkilo_b200.h"

namespace legkilo {
namespace b200 {

using Mat3D = Eigen::Matrix3d;
using Vec3D = Eigen::Vector3d;
using StateCov = Eigen::Matrix<double, 30, 30>;
using RowMat3 = Eigen::Matrix<double, 3, 3, Eigen::RowMajor>;
using RowCov = Eigen::Matrix<double, 30, 30, Eigen::RowMajor>;

// legkilo::State <-> lk_state (rot is row-major in the ABI).
template <class StateT>
inline lk_state toAbi(const StateT& s) {
    lk_state x;
    RowMat3 R = s.rot_;
    std::memcpy(x.rot, R.data(), sizeof(x.rot));
    const Vec3D* v[9] = {&s.pos_, &s.vel_, &s.ba_, &s.bw_, &s.grav_, &s.imu_a_, &s.imu_w_, &s.bv_, &s.contact_};
    double* d[9] = {x.pos, x.vel, x.ba, x.bw, x.grav, x.imu_a, x.imu_w, x.bv, x.contact};
    for (int i = 0; i < 9; ++i) std::memcpy(d[i], v[i]->data(), 24);
    return x;
}
template <class StateT>
inline void fromAbi(const lk_state& x, StateT& s) {
    s.rot_ = Eigen::Map<const RowMat3>(x.rot);
    Vec3D* v[9] = {&s.pos_, &s.vel_, &s.ba_, &s.bw_, &s.grav_, &s.imu_a_, &s.imu_w_, &s.bv_, &s.contact_};
    const double* d[9] = {x.pos, x.vel, x.ba, x.bw, x.grav, x.imu_a, x.imu_w, x.bv, x.contact};
    for (int i = 0; i < 9; ++i) *v[i] = Eigen::Map<const Vec3D>(d[i]);
}

// Owns the device context; what KILO keeps instead of unique_ptr<ESKF> + unique_ptr<VoxelMapManager>.
class Core {
   public:
    template <class EskfConfig, class VoxelMapConfig>
    Core(const EskfConfig& ec, const VoxelMapConfig& mc, const Mat3D& ext_rot, const Vec3D& ext_t, int device = 0) {
        static_assert(sizeof(EskfConfig) == sizeof(lk_eskf_cfg), "ESKF::Config layout (eskf.h:49-65)");
        std::memcpy(&ec_, &ec, sizeof(ec_));
        lk_map_cfg m{};
        m.max_voxel_size = mc.max_voxel_size_; m.planner_threshold = mc.planner_threshold_; m.beam_err = mc.beam_err_;
        m.dept_err = mc.dept_err_; m.sigma_num = mc.sigma_num_; m.max_layer = mc.max_layer_; m.max_points_num = mc.max_points_num_;
        for (int i = 0; i < 5 && i < (int)mc.layer_init_num_.size(); ++i) m.layer_init_num[i] = mc.layer_init_num_[i];
        RowMat3 Re = ext_rot;
        if (lk_create(&ec_, &m, Re.data(), ext_t.data(), device, &h_) != LK_OK) throw std::runtime_error(lk_last_error(nullptr));
        Q_.resize(900);
        lk_init_process_cov(&ec_, Q_.data());  // ESKF::initProcessCovQ (eskf.cc:47-62)
    }
    ~Core() { lk_destroy(h_); }
    Core(const Core&) = delete;
    Core& operator=(const Core&) = delete;

    // VoxelMapManager::BuildVoxelMap (voxel_map.cc:287-334). xyz arrays: n x 3 floats.
    void BuildVoxelMap(const float* xyz_world, const float* xyz_body, size_t n, const Mat3D& rot, const Mat3D& rot_cov,
                       const Mat3D& pos_cov) {
        RowMat3 R = rot, Cr = rot_cov, Cp = pos_cov;
        check(lk_map_build(h_, xyz_world, xyz_body, n, R.data(), Cr.data(), Cp.data()));
    }

    // The second lambda of KILO::process (KILO.cc:367-396) for one scan whose points are already in the
    // canonical (stable, ascending curvature) order. Exactly one of imu / kin may be non-empty.
    template <class StateT>
    size_t processScan(StateT& state, StateCov& cov, double& last_predict_time, double& last_update_time,
                       const std::vector<float>& xyzt, const std::vector<uint32_t>& bucket_offsets,
                       const std::vector<double>& bucket_times, std::vector<lk_imu_meas>& imu, std::vector<lk_kinimu_meas>& kin,
                       double gravity, double acc_norm, std::vector<float>& world_xyzi, int iters = 1, bool update_map = true) {
        lk_state x = toAbi(state);
        RowCov P = cov;
        lk_stream_clock clk{last_predict_time, last_update_time};
        uint32_t n_eff = 0, used = 0;
        const uint32_t n = (uint32_t)(xyzt.size() / 4);
        world_xyzi.resize(xyzt.size());
        check(lk_process_scan(h_, &x, P.data(), Q_.data(), &clk, xyzt.data(), n, bucket_offsets.data(), bucket_times.data(),
                              (uint32_t)bucket_times.size(), imu.empty() ? nullptr : imu.data(), kin.empty() ? nullptr : kin.data(),
                              (uint32_t)(imu.empty() ? kin.size() : imu.size()), gravity, acc_norm, iters, update_map ? 1 : 0,
                              world_xyzi.data(), &n_eff, &used));
        fromAbi(x, state);
        cov = P;
        last_predict_time = clk.last_predict_time;
        last_update_time = clk.last_update_time;
        if (!imu.empty()) imu.erase(imu.begin(), imu.begin() + used);  // the deque pop_front of KILO.cc:382, :388
        if (!kin.empty()) kin.erase(kin.begin(), kin.begin() + used);
        return n_eff;  // success_pts_size_out
    }

    // VoxelMapManager::mapSliding (voxel_map.cc:552-571): drop the root voxels that left the +-half_map_size window.
    bool mapSliding(const Vec3D& position_last, uint64_t* removed = nullptr) {
        int32_t slid = 0;
        check(lk_map_slide(h_, position_last.data(), &slid, removed));
        return slid != 0;
    }

    // TrajectorySaver::write (trajectory_saver.hpp:43-50): one TUM line of the current pose.
    template <class StateT>
    std::string tumLine(double timestamp, const StateT& state) const {
        RowMat3 R = state.rot_;
        char buf[256];
        const int n = lk_tum_line(timestamp, R.data(), state.pos_.data(), buf, sizeof(buf));
        if (n < 0) throw std::runtime_error("lk_tum_line");
        return std::string(buf, (size_t)n);
    }

    lk_handle handle() const { return h_; }

   private:
    void check(int rc) const {
        if (rc != LK_OK) throw std::runtime_error(lk_last_error(h_));
    }
    lk_handle h_ = nullptr;
    lk_eskf_cfg ec_{};
    std::vector<double> Q_;
};

}  // namespace b200
}  // namespace legkilo
#endif  // __has_include(<Eigen/Dense>)
