// oracle/lko_core.hpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of Leg-KILO's per-scan ESKF LiDAR point-to-plane measurement update, written
// from the reference's behaviour (not copied): every function cites the reference lines it
// follows. Float temporaries, thresholds and quirks of the reference are reproduced on purpose
// (SURVEY.md §8a). The reference ships no tests, golden vectors or fixtures; this restatement is
// pinned against oracle/_ref — the reference's own eskf.cc / voxel_map.cc / KILO.cc compiled from
// /root/reference over the stand-in third-party headers of oracle/ref/shim
// (tests/test_oracle_vs_reference.py) — and by the literal-vs-information gain check, the
// independent numpy mirror in tests/, analytic cases and property tests (oracle/README.md).
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <deque>
#include <unordered_map>
#include <vector>

#include "lko_linalg.hpp"

namespace lko {

constexpr int DIM = 30;
using StateVec = Mat<DIM, 1>;
using StateCov = Mat<DIM, DIM>;

// ---- SO(3) helpers: legkilo/src/common/math_utils.hpp:13-76 --------------------------------
M3 skew(const V3& v);
M3 Exp_vec(const V3& ang);                     // math_utils.hpp:20-32  (threshold 1e-7)
M3 Exp_vel_dt(const V3& ang_vel, double dt);   // math_utils.hpp:35-52  (threshold 1e-7)
M3 Exp3(double v1, double v2, double v3);      // math_utils.hpp:55-68  (threshold 1e-5)
V3 Log(const M3& R);                           // math_utils.hpp:72-76

// ---- State: legkilo/src/core/slam/eskf.h:15-32, eskf.cc:5-45 --------------------------------
struct State {
    M3 rot;
    V3 pos, vel, ba, bw, grav, imu_a, imu_w, bv, contact;
    State();
    void boxplus(const StateVec& delta);        // operator+=  eskf.cc:18-29
    StateVec boxminus(const State& other) const;  // operator-   eskf.cc:31-45
};

struct EskfConfig {  // eskf.h:49-65
    double vel_process_cov, imu_acc_process_cov, imu_gyr_process_cov, contact_process_cov,
        acc_bias_process_cov, gyr_bias_process_cov, kin_bias_process_cov, imu_acc_meas_noise,
        imu_acc_z_meas_noise, imu_gyr_meas_noise, kin_meas_noise, chd_meas_noise,
        contact_meas_noise, lidar_point_meas_ratio;
};

struct ObsPoints {  // ObsShared::pt_* eskf.h:34-44
    std::vector<double> h;  // N x 6
    std::vector<double> z;  // N
    std::vector<double> R;  // N
    int n() const { return (int)z.size(); }
};

enum GainMode { GAIN_LITERAL = 0, GAIN_INFORMATION = 1 };

struct PointGain {  // what an update applies: delta_x and the covariance decrement factors
    StateVec delta;
    Mat<DIM, 6> KH;  // K*h  (30x6): P <- P - KH * P[0:6,:]
};

class ESKF {  // eskf.h:46-109
   public:
    explicit ESKF(const EskfConfig& c) : config(c) {}
    EskfConfig config;
    State state;
    StateCov cov;
    StateCov Q;
    void initProcessCovQ();                                  // eskf.cc:47-62
    StateVec getFunctionf(double dt) const;                  // eskf.cc:64-70
    StateCov getFx(double dt) const;                         // eskf.cc:72-81
    void predict(double dt, bool prop_state, bool prop_cov);  // eskf.cc:83-89
    // eskf.cc:91-113 split in two so that the iterated variant can reuse it:
    PointGain pointGain(const ObsPoints& obs, GainMode mode) const;
    void applyPointGain(const PointGain& g, bool update_state, bool update_cov);
    void updateByPoints(const ObsPoints& obs, GainMode mode) {
        PointGain g = pointGain(obs, mode);
        applyPointGain(g, true, true);
    }
    void updateByImu(const double z[6], const double R[6]);                       // eskf.cc:125-135
    void updateByKinImu(int m, const std::vector<double>& H, const std::vector<double>& z,
                        const std::vector<double>& R);                            // eskf.cc:137-145
    M3 getRot() const { return state.rot; }
    V3 getPos() const { return state.pos; }
    V3 getVel() const { return state.vel; }
    M3 getRotCov() const { return cov.block<3, 3>(0, 0); }
    M3 getPosCov() const { return cov.block<3, 3>(3, 3); }
};

// ---- voxel map: legkilo/src/core/slam/voxel_map.h / voxel_map.cc ----------------------------
struct VoxelMapConfig {  // voxel_map.h:41-57
    double max_voxel_size_;
    int max_layer_;
    int max_iterations_;
    std::vector<int> layer_init_num_;
    int max_points_num_;
    double planner_threshold_;
    double beam_err_;
    double dept_err_;
    double sigma_num_;
};

struct pointWithVar {  // voxel_map.h:59-78
    V3 point_b, point_i, point_w;
    M3 body_var, var, point_crossmat;
    V3 normal;
};

struct PointToPlane {  // voxel_map.h:80-94
    V3 point_b_, point_w_, normal_, center_;
    M3 point_crossmat_;
    M6 plane_var_;
    M3 body_cov_;
    int layer_ = 0;
    double d_ = 0;
    float dis_to_plane_ = 0;
};

struct VoxelPlane {  // voxel_map.h:96-119
    V3 center_, normal_, y_normal_, x_normal_;
    M3 covariance_;
    M6 plane_var_;
    float radius_ = 0, min_eigen_value_ = 1, mid_eigen_value_ = 1, max_eigen_value_ = 1, d_ = 0;
    int points_size_ = 0;
    bool is_plane_ = false, is_init_ = false, is_update_ = false;
    int id_ = 0;
};

void calcBodyCov(V3& pb, const float range_inc, const float degree_inc, M3& cov);  // voxel_map.cc:22-40

class VoxelOctoTree {  // voxel_map.h:129-176
   public:
    std::vector<pointWithVar> temp_points_;
    VoxelPlane* plane_ptr_;
    int layer_;
    int octo_state_;
    VoxelOctoTree* leaves_[8];
    double voxel_center_[3];
    std::vector<int> layer_init_num_;
    float quater_length_;
    float planer_threshold_;
    int points_size_threshold_;
    int update_size_threshold_;
    int max_points_num_;
    int max_layer_;
    int new_points_;
    bool init_octo_;
    bool update_enable_;
    VoxelOctoTree(int max_layer, int layer, int points_size_threshold, int max_points_num,
                  float planer_threshold);
    ~VoxelOctoTree();
    void init_plane(const std::vector<pointWithVar>& points, VoxelPlane* plane);  // voxel_map.cc:42-117
    void init_octo_tree();                                                        // :119-137
    void cut_octo_tree();                                                         // :139-183
    void UpdateOctoTree(const pointWithVar& pv);                                  // :185-241
};

struct KeyHash {
    size_t operator()(const std::array<int, 3>& k) const;
};

class VoxelMapManager {  // voxel_map.h:180-244
   public:
    explicit VoxelMapManager(const VoxelMapConfig& c) : config_setting_(c) {}
    ~VoxelMapManager();
    VoxelMapConfig config_setting_;
    std::unordered_map<std::array<int, 3>, VoxelOctoTree*, KeyHash> voxel_map_;
    M3 extR_;
    V3 extT_;
    // feats_down_world_/feats_down_body_ of the reference, as float xyz triplets
    void BuildVoxelMap(const float* xyz_world, const float* xyz_body, size_t n, const M3& rot,
                       const M3& rot_cov, const M3& pos_cov);  // :287-334
    void UpdateVoxelMap(const std::vector<pointWithVar>& input_points);  // :336-361
    void build_single_residual(pointWithVar& pv, const VoxelOctoTree* current_octo,
                               const int current_layer, bool& is_success, double& prob,
                               PointToPlane& single_ptpl) const;  // :363-427
};

// ---- KILO orchestration: legkilo/src/core/slam/KILO.cc --------------------------------------
struct ImuMeas { double stamp, acc[3], gyr[3]; };
struct KinImuMeas {  // sensor_types.hpp:19-26
    double time_stamp_, foot_pos_[4][3], foot_vel_[4][3];
    bool contact_[4];
    double acc_[3], gyr_[3];
};

struct PointXYZT { float x, y, z, curvature; };
struct PointXYZI { float x, y, z, intensity; };

struct BucketDebug {  // per-point rows of one bucket (loop KILO.cc:122-210)
    std::vector<uint8_t> ok;
    std::vector<double> h, z, R;
    std::vector<int> key;
};

class KILO {
   public:
    KILO(const EskfConfig& ec, std::shared_ptr<VoxelMapManager> map, const M3& ext_rot, const V3& ext_t);
    ESKF eskf_;
    std::shared_ptr<VoxelMapManager> map_manager_;  // shared read-only across batch replicas
    bool imu_mode_only_ = true;
    double gravity_ = 9.81;
    double acc_norm_ = 1.0;
    double last_state_predict_time_ = 0.0;
    double last_state_update_time_ = 0.0;
    M3 ext_rot_;
    V3 ext_t_;
    // harness parameters that the reference fixes implicitly
    GainMode gain_mode_ = GAIN_LITERAL;
    int iters_ = 1;
    bool update_map_ = true;

    // KILO.cc:108-233. cloud_down_body/world indexed [idx_i, idx_j).
    bool predictUpdatePoint(double current_time, size_t idx_i, size_t idx_j,
                            const std::vector<PointXYZT>& cloud_down_body,
                            std::vector<PointXYZI>& cloud_down_world, size_t& success_pts_size_out,
                            BucketDebug* dbg = nullptr);
    bool predictUpdateImu(const ImuMeas& imu);          // KILO.cc:235-258
    bool predictUpdateKinImu(const KinImuMeas& kin_imu);  // KILO.cc:260-314
    // second lambda of KILO::process, KILO.cc:367-396; points must already be in canonical
    // (stable, ascending curvature) order — std::sort there is unstable.
    void processSorted(double begin_time, const std::vector<PointXYZT>& pts,
                       std::vector<PointXYZI>& world, std::deque<ImuMeas>& imus,
                       std::deque<KinImuMeas>& kin_imus, size_t& success_pts_size_out);
};

}  // namespace lko
