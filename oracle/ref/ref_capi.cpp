// oracle/ref/ref_capi.cpp — TEST INFRASTRUCTURE ONLY.
//
// C entry points (ctypes) over the REFERENCE's own sources for the hot path: legkilo/src/core/slam/eskf.cc,
// voxel_map.cc and KILO.cc are compiled unmodified from /root/reference by oracle/ref/Makefile and linked with this
// file into oracle/_ref/liblkref.so. The third-party libraries those sources include (Eigen, PCL, ROS messages, glog,
// yaml-cpp) are not in this image; oracle/ref/shim/ stands in for the slice of each the three files touch.
// This harness only moves data in and out: it fills the option registry KILO::initializeFromYaml reads, sets / gets the
// filter, and calls KILO's own handlers. The private members it needs are reached by compiling THIS translation unit
// with `private` spelled `public` (class layout is unchanged, so it links against the untouched objects).
// Uses the product's POD structs (include/legkilo_b200.h), like oracle/lko_capi.cpp, so tests hand both the same buffers.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <deque>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include <Eigen/Dense>
#include <pcl/filters/voxel_grid.h>
#include <pcl/point_types.h>
#include <ros/ros.h>
#include <sensor_msgs/Imu.h>
#include <visualization_msgs/MarkerArray.h>

#include "common/yaml_helper.hpp"

#define private public
#include "core/slam/KILO.h"
#include "core/slam/eskf.h"
#include "core/slam/voxel_map.h"
#undef private

#include "../../include/legkilo_b200.h"

using namespace legkilo;

namespace {

struct Ctx {
    std::unique_ptr<KILO> kilo;
};

void reg(const char* key, double v) { yaml_registry()[key] = {v}; }

Mat3D to_m3(const double* a) {
    Mat3D m;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) m(i, j) = a[3 * i + j];
    return m;
}

sensor_msgs::ImuPtr imu_from(const lk_imu_meas& m) {
    sensor_msgs::ImuPtr p(new sensor_msgs::Imu());
    p->header.stamp = ros::Time(m.stamp);
    p->linear_acceleration.x = m.acc[0];
    p->linear_acceleration.y = m.acc[1];
    p->linear_acceleration.z = m.acc[2];
    p->angular_velocity.x = m.gyr[0];
    p->angular_velocity.y = m.gyr[1];
    p->angular_velocity.z = m.gyr[2];
    return p;
}

common::KinImuMeas kin_from(const lk_kinimu_meas& k) {
    common::KinImuMeas m;
    m.time_stamp_ = k.stamp;
    std::memcpy(m.foot_pos_, k.foot_pos, sizeof(m.foot_pos_));
    std::memcpy(m.foot_vel_, k.foot_vel, sizeof(m.foot_vel_));
    for (int i = 0; i < 4; ++i) m.contact_[i] = k.contact[i] != 0;
    std::memcpy(m.acc_, k.acc, sizeof(m.acc_));
    std::memcpy(m.gyr_, k.gyr, sizeof(m.gyr_));
    return m;
}

// The lk_map blob (include/legkilo_b200.h) of the reference's unordered_map<Vector3i, VoxelOctoTree*>: roots by
// ascending key, children of a node as eight contiguous slots — the layout oracle/lko_capi.cpp writes for its own map.
struct Exporter {
    std::vector<lk_map_root> roots;
    std::vector<lk_map_node> nodes;
    std::vector<lk_map_aux> aux;
    std::vector<lk_map_point> points;

    void fill(int idx, const VoxelOctoTree* t, int parent) {
        lk_map_node n;
        lk_map_aux a;
        std::memset(&n, 0, sizeof(n));
        std::memset(&a, 0, sizeof(a));
        const VoxelPlane& p = *t->plane_ptr_;
        for (int k = 0; k < 3; ++k) {
            n.center[k] = p.center_[k];
            n.normal[k] = p.normal_[k];
            a.voxel_center[k] = t->voxel_center_[k];
        }
        int q = 0;
        for (int i = 0; i < 6; ++i)
            for (int j = i; j < 6; ++j) n.plane_var[q++] = p.plane_var_(i, j);
        n.d = p.d_;
        n.radius = p.radius_;
        uint32_t mask = 0;
        for (int i = 0; i < 8; ++i)
            if (t->leaves_[i]) mask |= 1u << i;
        n.flags = (p.is_plane_ ? LK_NODE_IS_PLANE : 0u) | (t->init_octo_ ? LK_NODE_INIT_OCTO : 0u) |
                  (t->update_enable_ ? LK_NODE_UPDATE_ENABLE : 0u) | ((uint32_t)t->layer_ << LK_NODE_LAYER_SHIFT) |
                  (mask << LK_NODE_CHILDMASK_SHIFT);
        n.child_base = -1;
        a.quater_length = t->quater_length_;
        a.pts_base = (uint32_t)points.size();
        a.pts_count = (int32_t)t->temp_points_.size();
        a.pts_cap = a.pts_count;
        a.new_points = t->new_points_;
        a.parent = parent;
        for (const auto& pv : t->temp_points_) {
            lk_map_point mp;
            for (int k = 0; k < 3; ++k) mp.pw[k] = pv.point_w[k];
            mp.var[0] = pv.var(0, 0); mp.var[1] = pv.var(0, 1); mp.var[2] = pv.var(0, 2);
            mp.var[3] = pv.var(1, 1); mp.var[4] = pv.var(1, 2); mp.var[5] = pv.var(2, 2);
            points.push_back(mp);
        }
        if (mask) {
            const int base = (int)nodes.size();
            n.child_base = base;
            nodes.resize(nodes.size() + 8);
            aux.resize(aux.size() + 8);
            for (int i = 0; i < 8; ++i) {
                std::memset(&nodes[base + i], 0, sizeof(lk_map_node));
                std::memset(&aux[base + i], 0, sizeof(lk_map_aux));
                nodes[base + i].child_base = -1;
                aux[base + i].parent = idx;
            }
        }
        nodes[idx] = n;
        aux[idx] = a;
        if (mask)
            for (int i = 0; i < 8; ++i)
                if (t->leaves_[i]) fill(n.child_base + i, t->leaves_[i], idx);
    }
};

}  // namespace

extern "C" {

// KILO::initializeFromYaml (KILO.cc:25-84) reads these keys; the names are the reference's YAML keys.
void* lkref_create(const lk_eskf_cfg* ec, const lk_map_cfg* mc, const double extR[9], const double extT[3],
                   int imu_mode_only, double gravity) {
    yaml_registry().clear();
    reg("only_imu_use", imu_mode_only ? 1.0 : 0.0);
    reg("vel_process_cov", ec->vel_process_cov);
    reg("imu_acc_process_cov", ec->imu_acc_process_cov);
    reg("imu_gyr_process_cov", ec->imu_gyr_process_cov);
    reg("acc_bias_process_cov", ec->acc_bias_process_cov);
    reg("gyr_bias_process_cov", ec->gyr_bias_process_cov);
    reg("kin_bias_process_cov", ec->kin_bias_process_cov);
    reg("contact_process_cov", ec->contact_process_cov);
    reg("imu_acc_meas_noise", ec->imu_acc_meas_noise);
    reg("imu_acc_z_meas_noise", ec->imu_acc_z_meas_noise);
    reg("imu_gyr_meas_noise", ec->imu_gyr_meas_noise);
    reg("kin_meas_noise", ec->kin_meas_noise);
    reg("chd_meas_noise", ec->chd_meas_noise);
    reg("contact_meas_noise", ec->contact_meas_noise);
    reg("lidar_point_meas_ratio", ec->lidar_point_meas_ratio);
    reg("gravity", gravity);
    reg("pub_plane_en", 0.0);
    reg("max_layer", mc->max_layer);
    reg("voxel_size", mc->max_voxel_size);
    reg("min_eigen_value", mc->planner_threshold);
    reg("sigma_num", mc->sigma_num);
    reg("beam_err", mc->beam_err);
    reg("dept_err", mc->dept_err);
    yaml_registry()["layer_init_num"] = std::vector<double>(mc->layer_init_num, mc->layer_init_num + 5);
    reg("max_points_num", mc->max_points_num);
    reg("map_sliding_en", mc->map_sliding_en);
    reg("half_map_size", mc->half_map_size);
    reg("sliding_thresh", mc->sliding_thresh);
    yaml_registry()["extrinsic_T"] = std::vector<double>(extT, extT + 3);
    yaml_registry()["extrinsic_R"] = std::vector<double>(extR, extR + 9);
    reg("voxel_grid_resolution", 0.0);  // the stand-in VoxelGrid passes the (already downsampled) cloud through
    Ctx* c = new Ctx;
    c->kilo.reset(new KILO("registry"));
    c->kilo->eskf_->cov_.setZero();
    c->kilo->eskf_->Q_.setZero();
    return c;
}

void lkref_destroy(void* h) { delete (Ctx*)h; }

// P and Q are row-major 30 x 30 (the ABI's convention).
void lkref_set_filter(void* h, const lk_state* x, const double* P, const double* Q, const lk_stream_clock* clk) {
    KILO& k = *((Ctx*)h)->kilo;
    if (x) {
        State& s = k.eskf_->state_;
        s.rot_ = to_m3(x->rot);
        Vec3D* dst[9] = {&s.pos_, &s.vel_, &s.ba_, &s.bw_, &s.grav_, &s.imu_a_, &s.imu_w_, &s.bv_, &s.contact_};
        const double* src[9] = {x->pos, x->vel, x->ba, x->bw, x->grav, x->imu_a, x->imu_w, x->bv, x->contact};
        for (int b = 0; b < 9; ++b)
            for (int j = 0; j < 3; ++j) (*dst[b])(j) = src[b][j];
    }
    for (int i = 0; i < 30; ++i)
        for (int j = 0; j < 30; ++j) {
            if (P) k.eskf_->cov_(i, j) = P[30 * i + j];
            if (Q) k.eskf_->Q_(i, j) = Q[30 * i + j];
        }
    if (clk) {
        k.last_state_predict_time_ = clk->last_predict_time;
        k.last_state_update_time_ = clk->last_update_time;
    }
}

void lkref_get_filter(void* h, lk_state* x, double* P, double* Q, lk_stream_clock* clk) {
    KILO& k = *((Ctx*)h)->kilo;
    if (x) {
        const State& s = k.eskf_->state_;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) x->rot[3 * i + j] = s.rot_(i, j);
        const Vec3D* src[9] = {&s.pos_, &s.vel_, &s.ba_, &s.bw_, &s.grav_, &s.imu_a_, &s.imu_w_, &s.bv_, &s.contact_};
        double* dst[9] = {x->pos, x->vel, x->ba, x->bw, x->grav, x->imu_a, x->imu_w, x->bv, x->contact};
        for (int b = 0; b < 9; ++b)
            for (int j = 0; j < 3; ++j) dst[b][j] = (*src[b])(j);
    }
    for (int i = 0; i < 30; ++i)
        for (int j = 0; j < 30; ++j) {
            if (P) P[30 * i + j] = k.eskf_->cov_(i, j);
            if (Q) Q[30 * i + j] = k.eskf_->Q_(i, j);
        }
    if (clk) {
        clk->last_predict_time = k.last_state_predict_time_;
        clk->last_update_time = k.last_state_update_time_;
    }
}

// What the first frame of KILO::process leaves behind besides the filter (KILO.cc:351-354).
void lkref_set_runtime(void* h, double acc_norm, int initialised) {
    KILO& k = *((Ctx*)h)->kilo;
    k.acc_norm_ = acc_norm;
    k.init_flag_ = !initialised;
}
double lkref_acc_norm(void* h) { return ((Ctx*)h)->kilo->acc_norm_; }

void lkref_init_process_cov(void* h) { ((Ctx*)h)->kilo->eskf_->initProcessCovQ(); }

void lkref_predict(void* h, double dt, int prop_state, int prop_cov) {
    ((Ctx*)h)->kilo->eskf_->predict(dt, prop_state != 0, prop_cov != 0);
}

// VoxelMapManager::BuildVoxelMap (voxel_map.cc:287-334) on explicit world / body clouds (n x 3 floats).
void lkref_build_voxel_map(void* h, const float* xyz_world, const float* xyz_body, size_t n, const double R[9],
                           const double rot_cov[9], const double pos_cov[9]) {
    KILO& k = *((Ctx*)h)->kilo;
    CloudPtr w(new PointCloudType()), b(new PointCloudType());
    w->points.resize(n);
    b->points.resize(n);
    for (size_t i = 0; i < n; ++i) {
        w->points[i].x = xyz_world[3 * i]; w->points[i].y = xyz_world[3 * i + 1]; w->points[i].z = xyz_world[3 * i + 2];
        b->points[i].x = xyz_body[3 * i]; b->points[i].y = xyz_body[3 * i + 1]; b->points[i].z = xyz_body[3 * i + 2];
    }
    k.map_manager_->feats_down_body_ = b;
    k.map_manager_->feats_down_world_ = w;
    k.map_manager_->BuildVoxelMap(to_m3(R), to_m3(rot_cov), to_m3(pos_cov));
}

// One bucket: KILO::predictUpdatePoint (KILO.cc:108-233). pts4 = (x, y, z, curvature) per point.
int lkref_predict_update_point(void* h, double t, const float* pts4, uint32_t n, float* world4_out,
                               uint32_t* n_eff_inout) {
    KILO& k = *((Ctx*)h)->kilo;
    PointCloudType body, world;
    body.points.resize(n);
    world.points.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
        body.points[i].x = pts4[4 * i]; body.points[i].y = pts4[4 * i + 1]; body.points[i].z = pts4[4 * i + 2];
        body.points[i].curvature = pts4[4 * i + 3];
    }
    size_t succ = n_eff_inout ? *n_eff_inout : 0;
    const bool upd = k.predictUpdatePoint(t, 0, n, body, world, succ);
    if (world4_out)
        for (uint32_t i = 0; i < n; ++i) {
            world4_out[4 * i] = world.points[i].x; world4_out[4 * i + 1] = world.points[i].y;
            world4_out[4 * i + 2] = world.points[i].z; world4_out[4 * i + 3] = world.points[i].intensity;
        }
    if (n_eff_inout) *n_eff_inout = (uint32_t)succ;
    return upd ? 1 : 0;
}

void lkref_obs_imu(void* h, const lk_imu_meas* imu, uint32_t n) {
    KILO& k = *((Ctx*)h)->kilo;
    for (uint32_t i = 0; i < n; ++i) k.predictUpdateImu(imu_from(imu[i]));
}

void lkref_obs_kinimu(void* h, const lk_kinimu_meas* kin, uint32_t n) {
    KILO& k = *((Ctx*)h)->kilo;
    for (uint32_t i = 0; i < n; ++i) k.predictUpdateKinImu(kin_from(kin[i]));
}

// A whole KILO::process call (KILO.cc:316-399): first frame = initialisation + BuildVoxelMap, later frames = sort +
// bucket loop. body4_out receives the cloud in the order process() left it in (std::sort by curvature is not stable,
// so the test feeds the oracle THIS order), world4_out the matching world cloud. Returns process()'s result.
int lkref_process(void* h, double begin_time, double end_time, const float* pts4, uint32_t n, const lk_imu_meas* imu,
                  uint32_t n_imu, const lk_kinimu_meas* kin, uint32_t n_kin, float* body4_out, float* world4_out,
                  uint32_t* n_eff_out) {
    KILO& k = *((Ctx*)h)->kilo;
    common::MeasGroup mg;
    mg.lidar_scan_.lidar_begin_time_ = begin_time;
    mg.lidar_scan_.lidar_end_time_ = end_time;
    mg.lidar_scan_.cloud_.reset(new PointCloudType());
    mg.lidar_scan_.cloud_->points.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
        PointType& p = mg.lidar_scan_.cloud_->points[i];
        p.x = pts4[4 * i]; p.y = pts4[4 * i + 1]; p.z = pts4[4 * i + 2]; p.curvature = pts4[4 * i + 3];
    }
    for (uint32_t i = 0; i < n_imu; ++i) mg.imus_.push_back(imu_from(imu[i]));
    for (uint32_t i = 0; i < n_kin; ++i) mg.kin_imus_.push_back(kin_from(kin[i]));
    const bool first = k.init_flag_;
    CloudPtr body, world;
    size_t succ = 0;
    const bool ok = k.process(mg, body, world, succ);
    if (ok) {
        const PointCloudType& b = first ? *mg.lidar_scan_.cloud_ : *body;
        for (uint32_t i = 0; i < n && i < b.points.size(); ++i) {
            if (body4_out) {
                body4_out[4 * i] = b.points[i].x; body4_out[4 * i + 1] = b.points[i].y;
                body4_out[4 * i + 2] = b.points[i].z; body4_out[4 * i + 3] = b.points[i].curvature;
            }
            if (world4_out) {
                world4_out[4 * i] = world->points[i].x; world4_out[4 * i + 1] = world->points[i].y;
                world4_out[4 * i + 2] = world->points[i].z; world4_out[4 * i + 3] = world->points[i].intensity;
            }
        }
    }
    if (n_eff_out) *n_eff_out = (uint32_t)succ;
    return ok ? 1 : 0;
}

// VoxelMapManager::mapSliding (voxel_map.cc:552-571).
int lkref_map_slide(void* h, const double position_last[3]) {
    KILO& k = *((Ctx*)h)->kilo;
    k.map_manager_->position_last_ = Vec3D(position_last[0], position_last[1], position_last[2]);
    return k.map_manager_->mapSliding() ? 1 : 0;
}

void lkref_calc_body_cov(double pb[3], float range_inc, float degree_inc, double cov[9]) {
    Eigen::Vector3d p(pb[0], pb[1], pb[2]);
    Eigen::Matrix3d cv;
    calcBodyCov(p, range_inc, degree_inc, cv);
    for (int k = 0; k < 3; ++k) pb[k] = p[k];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) cov[3 * i + j] = cv(i, j);
}

// VoxelOctoTree::init_plane (voxel_map.cc:44-122) on explicit points: pw[n*3], var[n*9] row-major.
void lkref_init_plane(uint32_t n, const double* pw, const double* var, float planer_threshold, double* center,
                      double* normal, double* plane_var36, float* d, float* radius, int* is_plane, float* eig3) {
    VoxelOctoTree t(2, 0, 5, 50, planer_threshold);
    std::vector<pointWithVar> pts(n);
    for (uint32_t i = 0; i < n; ++i) {
        pts[i].point_w = Eigen::Vector3d(pw[3 * i], pw[3 * i + 1], pw[3 * i + 2]);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) pts[i].var(r, c) = var[9 * i + 3 * r + c];
    }
    t.init_plane(pts, t.plane_ptr_);
    const VoxelPlane& p = *t.plane_ptr_;
    for (int k = 0; k < 3; ++k) {
        center[k] = p.center_[k];
        normal[k] = p.normal_[k];
    }
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) plane_var36[6 * r + c] = p.plane_var_(r, c);
    *d = p.d_;
    *radius = p.radius_;
    *is_plane = p.is_plane_ ? 1 : 0;
    eig3[0] = p.min_eigen_value_;
    eig3[1] = p.mid_eigen_value_;
    eig3[2] = p.max_eigen_value_;
}

void lkref_boxminus(const lk_state* a, const lk_state* b, double out30[30]) {
    State x, y;
    auto load = [](const lk_state* s, State& o) {
        o.rot_ = to_m3(s->rot);
        Vec3D* dst[9] = {&o.pos_, &o.vel_, &o.ba_, &o.bw_, &o.grav_, &o.imu_a_, &o.imu_w_, &o.bv_, &o.contact_};
        const double* src[9] = {s->pos, s->vel, s->ba, s->bw, s->grav, s->imu_a, s->imu_w, s->bv, s->contact};
        for (int q = 0; q < 9; ++q)
            for (int j = 0; j < 3; ++j) (*dst[q])(j) = src[q][j];
    };
    load(a, x);
    load(b, y);
    StateVec dlt = x - y;
    for (int i = 0; i < 30; ++i) out30[i] = dlt(i);
}

uint64_t lkref_map_num_roots(void* h) { return ((Ctx*)h)->kilo->map_manager_->voxel_map_.size(); }

int lkref_map_export(void* h, void* buf, size_t cap, size_t* bytes_out) {
    KILO& k = *((Ctx*)h)->kilo;
    Exporter ex;
    std::vector<std::pair<std::array<int, 3>, const VoxelOctoTree*>> items;
    for (auto& kv : k.map_manager_->voxel_map_)
        items.push_back({std::array<int, 3>{kv.first[0], kv.first[1], kv.first[2]}, kv.second});
    std::sort(items.begin(), items.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    ex.nodes.resize(items.size());
    ex.aux.resize(items.size());
    int idx = 0;
    for (auto& it : items) {
        lk_map_root r;
        for (int q = 0; q < 3; ++q) r.key[q] = it.first[q];
        r.node = idx;
        ex.roots.push_back(r);
        ex.fill(idx, it.second, -1);
        for (int q = 0; q < 3; ++q) ex.aux[idx].key[q] = it.first[q];
        ++idx;
    }
    lk_map_blob_header hd;
    std::memset(&hd, 0, sizeof(hd));
    hd.magic = LK_MAP_MAGIC;
    hd.version = 1;
    hd.n_roots = (uint32_t)ex.roots.size();
    hd.n_nodes = (uint32_t)ex.nodes.size();
    hd.n_points = ex.points.size();
    const size_t bytes = sizeof(hd) + ex.roots.size() * sizeof(lk_map_root) + ex.nodes.size() * sizeof(lk_map_node) +
                         ex.aux.size() * sizeof(lk_map_aux) + ex.points.size() * sizeof(lk_map_point);
    if (bytes_out) *bytes_out = bytes;
    if (!buf) return 0;
    if (cap < bytes) return -1;
    char* p = (char*)buf;
    auto put = [&p](const void* src, size_t nb) {
        if (nb) std::memcpy(p, src, nb);
        p += nb;
    };
    put(&hd, sizeof(hd));
    put(ex.roots.data(), ex.roots.size() * sizeof(lk_map_root));
    put(ex.nodes.data(), ex.nodes.size() * sizeof(lk_map_node));
    put(ex.aux.data(), ex.aux.size() * sizeof(lk_map_aux));
    put(ex.points.data(), ex.points.size() * sizeof(lk_map_point));
    return 0;
}

}  // extern "C"
