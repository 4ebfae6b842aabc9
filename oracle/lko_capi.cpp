// oracle/lko_capi.cpp — TEST INFRASTRUCTURE ONLY. C entry points (ctypes) over the CPU
// restatement in lko_core.cpp. Uses the product's POD structs (include/legkilo_b200.h) so that
// tests can hand identical buffers to the oracle and to the CUDA library.
#include <algorithm>
#include <array>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstring>
#include <malloc.h>
#include <pthread.h>
#include <sched.h>
#include <thread>
#include <unistd.h>

#include "../include/legkilo_b200.h"
#include "lko_core.hpp"

using namespace lko;

namespace {

struct Ctx {
    std::shared_ptr<VoxelMapManager> map;
    std::unique_ptr<KILO> kilo;
    EskfConfig ec;
    VoxelMapConfig mc;
    M3 extR;
    V3 extT;
};

EskfConfig to_ec(const lk_eskf_cfg* c) {
    EskfConfig e;
    static_assert(sizeof(EskfConfig) == sizeof(lk_eskf_cfg), "layout");
    std::memcpy(&e, c, sizeof(e));
    return e;
}

VoxelMapConfig to_mc(const lk_map_cfg* c) {
    VoxelMapConfig m;
    m.max_voxel_size_ = c->max_voxel_size;
    m.max_layer_ = c->max_layer;
    m.max_iterations_ = c->max_iterations;
    m.layer_init_num_.assign(c->layer_init_num, c->layer_init_num + 5);
    m.max_points_num_ = c->max_points_num;
    m.planner_threshold_ = c->planner_threshold;
    m.beam_err_ = c->beam_err;
    m.dept_err_ = c->dept_err;
    m.sigma_num_ = c->sigma_num;
    return m;
}

M3 to_m3(const double* a) {
    M3 m;
    for (int i = 0; i < 9; ++i) m.a[i] = a[i];
    return m;
}

void state_from(const lk_state* s, State& x) {
    x.rot = to_m3(s->rot);
    const double* src[9] = {s->pos, s->vel, s->ba, s->bw, s->grav, s->imu_a, s->imu_w, s->bv, s->contact};
    V3* dst[9] = {&x.pos, &x.vel, &x.ba, &x.bw, &x.grav, &x.imu_a, &x.imu_w, &x.bv, &x.contact};
    for (int b = 0; b < 9; ++b)
        for (int k = 0; k < 3; ++k) (*dst[b])[k] = src[b][k];
}

void state_to(const State& x, lk_state* s) {
    for (int i = 0; i < 9; ++i) s->rot[i] = x.rot.a[i];
    double* dst[9] = {s->pos, s->vel, s->ba, s->bw, s->grav, s->imu_a, s->imu_w, s->bv, s->contact};
    const V3* src[9] = {&x.pos, &x.vel, &x.ba, &x.bw, &x.grav, &x.imu_a, &x.imu_w, &x.bv, &x.contact};
    for (int b = 0; b < 9; ++b)
        for (int k = 0; k < 3; ++k) dst[b][k] = (*src[b])[k];
}

KinImuMeas kin_from(const lk_kinimu_meas& k) {
    KinImuMeas m;
    m.time_stamp_ = k.stamp;
    std::memcpy(m.foot_pos_, k.foot_pos, sizeof(m.foot_pos_));
    std::memcpy(m.foot_vel_, k.foot_vel, sizeof(m.foot_vel_));
    for (int i = 0; i < 4; ++i) m.contact_[i] = k.contact[i] != 0;
    std::memcpy(m.acc_, k.acc, sizeof(m.acc_));
    std::memcpy(m.gyr_, k.gyr, sizeof(m.gyr_));
    return m;
}

// ---- map export in the lk_map blob format -------------------------------------------------
struct Exporter {
    std::vector<lk_map_root> roots;
    std::vector<lk_map_node> nodes;
    std::vector<lk_map_aux> aux;
    std::vector<lk_map_point> points;

    void fill(int idx, const VoxelOctoTree* t, int parent) {
        lk_map_node n;
        std::memset(&n, 0, sizeof(n));
        lk_map_aux a;
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
        uint32_t flags = 0;
        if (p.is_plane_) flags |= LK_NODE_IS_PLANE;
        if (t->init_octo_) flags |= LK_NODE_INIT_OCTO;
        if (t->update_enable_) flags |= LK_NODE_UPDATE_ENABLE;
        flags |= (uint32_t)t->layer_ << LK_NODE_LAYER_SHIFT;
        uint32_t mask = 0;
        for (int i = 0; i < 8; ++i)
            if (t->leaves_[i]) mask |= 1u << i;
        flags |= mask << LK_NODE_CHILDMASK_SHIFT;
        n.flags = flags;
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
            int base = (int)nodes.size();
            n.child_base = base;
            nodes.resize(nodes.size() + 8);
            aux.resize(aux.size() + 8);
            for (int i = 0; i < 8; ++i) {
                std::memset(&nodes[base + i], 0, sizeof(lk_map_node));
                std::memset(&aux[base + i], 0, sizeof(lk_map_aux));
                nodes[base + i].child_base = -1;
                aux[base + i].parent = idx;
            }
            nodes[idx] = n;
            aux[idx] = a;
            for (int i = 0; i < 8; ++i)
                if (t->leaves_[i]) fill(base + i, t->leaves_[i], idx);
        } else {
            nodes[idx] = n;
            aux[idx] = a;
        }
    }
};

}  // namespace

extern "C" {

void* lko_create(const lk_eskf_cfg* ec, const lk_map_cfg* mc, const double extR[9], const double extT[3]) {
    Ctx* c = new Ctx;
    c->ec = to_ec(ec);
    c->mc = to_mc(mc);
    c->extR = to_m3(extR);
    c->extT = vec3(extT[0], extT[1], extT[2]);
    c->map = std::make_shared<VoxelMapManager>(c->mc);
    c->kilo.reset(new KILO(c->ec, c->map, c->extR, c->extT));
    return c;
}

void lko_destroy(void* h) { delete (Ctx*)h; }

void lko_set_filter(void* h, const lk_state* x, const double* P, const double* Q, const lk_stream_clock* clk) {
    Ctx* c = (Ctx*)h;
    if (x) state_from(x, c->kilo->eskf_.state);
    if (P) std::memcpy(c->kilo->eskf_.cov.a, P, sizeof(double) * 900);
    if (Q) std::memcpy(c->kilo->eskf_.Q.a, Q, sizeof(double) * 900);
    if (clk) {
        c->kilo->last_state_predict_time_ = clk->last_predict_time;
        c->kilo->last_state_update_time_ = clk->last_update_time;
    }
}

void lko_get_filter(void* h, lk_state* x, double* P, double* Q, lk_stream_clock* clk) {
    Ctx* c = (Ctx*)h;
    if (x) state_to(c->kilo->eskf_.state, x);
    if (P) std::memcpy(P, c->kilo->eskf_.cov.a, sizeof(double) * 900);
    if (Q) std::memcpy(Q, c->kilo->eskf_.Q.a, sizeof(double) * 900);
    if (clk) {
        clk->last_predict_time = c->kilo->last_state_predict_time_;
        clk->last_update_time = c->kilo->last_state_update_time_;
    }
}

void lko_set_options(void* h, int gain_mode, int iters, int update_map, int imu_mode_only, double gravity,
                     double acc_norm) {
    Ctx* c = (Ctx*)h;
    c->kilo->gain_mode_ = gain_mode ? GAIN_INFORMATION : GAIN_LITERAL;
    c->kilo->iters_ = iters;
    c->kilo->update_map_ = update_map != 0;
    c->kilo->imu_mode_only_ = imu_mode_only != 0;
    c->kilo->gravity_ = gravity;
    c->kilo->acc_norm_ = acc_norm;
}

void lko_predict(void* h, double dt, int prop_state, int prop_cov) {
    ((Ctx*)h)->kilo->eskf_.predict(dt, prop_state != 0, prop_cov != 0);
}

void lko_build_voxel_map(void* h, const float* xyz_world, const float* xyz_body, size_t n, const double R[9],
                         const double rot_cov[9], const double pos_cov[9]) {
    Ctx* c = (Ctx*)h;
    c->map->BuildVoxelMap(xyz_world, xyz_body, n, to_m3(R), to_m3(rot_cov), to_m3(pos_cov));
}

// One bucket: KILO::predictUpdatePoint. pts4 = float4 (x,y,z,curvature). Debug arrays nullable.
int lko_predict_update_point(void* h, double t, const float* pts4, uint32_t n, float* world4_out,
                             uint32_t* n_eff_inout, uint8_t* ok_out, double* h_out, double* z_out, double* R_out,
                             int32_t* key_out) {
    Ctx* c = (Ctx*)h;
    std::vector<PointXYZT> body(n);
    std::memcpy(body.data(), pts4, sizeof(float) * 4 * n);
    std::vector<PointXYZI> world(n);
    size_t succ = n_eff_inout ? *n_eff_inout : 0;
    BucketDebug dbg;
    bool upd = c->kilo->predictUpdatePoint(t, 0, n, body, world, succ, ok_out ? &dbg : nullptr);
    if (world4_out) std::memcpy(world4_out, world.data(), sizeof(float) * 4 * n);
    if (n_eff_inout) *n_eff_inout = (uint32_t)succ;
    if (ok_out) {
        std::memcpy(ok_out, dbg.ok.data(), n);
        if (h_out) std::memcpy(h_out, dbg.h.data(), sizeof(double) * 6 * n);
        if (z_out) std::memcpy(z_out, dbg.z.data(), sizeof(double) * n);
        if (R_out) std::memcpy(R_out, dbg.R.data(), sizeof(double) * n);
        if (key_out)
            for (size_t i = 0; i < (size_t)n * 3; ++i) key_out[i] = dbg.key[i];
    }
    return upd ? 1 : 0;
}

// The bucket loop of KILO::process on already ordered points (stable by curvature).
int lko_process_scan(void* h, double begin_time, const float* pts4, uint32_t n, const lk_imu_meas* imu,
                     uint32_t n_imu, const lk_kinimu_meas* kin, uint32_t n_kin, float* world4_out,
                     uint32_t* n_eff_out, uint32_t* n_consumed_out) {
    Ctx* c = (Ctx*)h;
    std::vector<PointXYZT> body(n);
    std::memcpy(body.data(), pts4, sizeof(float) * 4 * n);
    std::vector<PointXYZI> world;
    std::deque<ImuMeas> imus;
    std::deque<KinImuMeas> kins;
    for (uint32_t i = 0; i < n_imu; ++i) {
        ImuMeas m;
        m.stamp = imu[i].stamp;
        std::memcpy(m.acc, imu[i].acc, sizeof(m.acc));
        std::memcpy(m.gyr, imu[i].gyr, sizeof(m.gyr));
        imus.push_back(m);
    }
    for (uint32_t i = 0; i < n_kin; ++i) kins.push_back(kin_from(kin[i]));
    size_t succ = 0;
    c->kilo->processSorted(begin_time, body, world, imus, kins, succ);
    if (world4_out) std::memcpy(world4_out, world.data(), sizeof(float) * 4 * n);
    if (n_eff_out) *n_eff_out = (uint32_t)succ;
    if (n_consumed_out)
        *n_consumed_out = c->kilo->imu_mode_only_ ? (uint32_t)(n_imu - imus.size()) : (uint32_t)(n_kin - kins.size());
    return 0;
}

void lko_update_by_points(void* h, uint32_t n, const double* ph, const double* pz, const double* pR, int gain_mode) {
    Ctx* c = (Ctx*)h;
    ObsPoints obs;
    obs.h.assign(ph, ph + (size_t)n * 6);
    obs.z.assign(pz, pz + n);
    obs.R.assign(pR, pR + n);
    c->kilo->eskf_.updateByPoints(obs, gain_mode ? GAIN_INFORMATION : GAIN_LITERAL);
}

void lko_obs_imu(void* h, const lk_imu_meas* imu, uint32_t n) {
    Ctx* c = (Ctx*)h;
    for (uint32_t i = 0; i < n; ++i) {
        ImuMeas m;
        m.stamp = imu[i].stamp;
        std::memcpy(m.acc, imu[i].acc, sizeof(m.acc));
        std::memcpy(m.gyr, imu[i].gyr, sizeof(m.gyr));
        c->kilo->predictUpdateImu(m);
    }
}

void lko_obs_kinimu(void* h, const lk_kinimu_meas* kin, uint32_t n) {
    Ctx* c = (Ctx*)h;
    for (uint32_t i = 0; i < n; ++i) c->kilo->predictUpdateKinImu(kin_from(kin[i]));
}

void lko_calc_body_cov(double pb[3], float range_inc, float degree_inc, double cov[9]) {
    V3 p = vec3(pb[0], pb[1], pb[2]);
    M3 cv;
    calcBodyCov(p, range_inc, degree_inc, cv);
    for (int k = 0; k < 3; ++k) pb[k] = p[k];
    for (int i = 0; i < 9; ++i) cov[i] = cv.a[i];
}

// init_plane on explicit points: pw[n*3], var[n*9]. out: center3 normal3 plane_var36 d radius
// is_plane eig(min,mid,max).
void lko_init_plane(uint32_t n, const double* pw, const double* var, float planer_threshold, double* center,
                    double* normal, double* plane_var36, float* d, float* radius, int* is_plane, float* eig3) {
    VoxelOctoTree t(2, 0, 5, 50, planer_threshold);
    std::vector<pointWithVar> pts(n);
    for (uint32_t i = 0; i < n; ++i) {
        pts[i].point_w = vec3(pw[3 * i], pw[3 * i + 1], pw[3 * i + 2]);
        for (int k = 0; k < 9; ++k) pts[i].var.a[k] = var[9 * i + k];
    }
    t.init_plane(pts, t.plane_ptr_);
    const VoxelPlane& p = *t.plane_ptr_;
    for (int k = 0; k < 3; ++k) {
        center[k] = p.center_[k];
        normal[k] = p.normal_[k];
    }
    for (int k = 0; k < 36; ++k) plane_var36[k] = p.plane_var_.a[k];
    *d = p.d_;
    *radius = p.radius_;
    *is_plane = p.is_plane_ ? 1 : 0;
    eig3[0] = p.min_eigen_value_;
    eig3[1] = p.mid_eigen_value_;
    eig3[2] = p.max_eigen_value_;
}

void lko_boxminus(const lk_state* a, const lk_state* b, double out30[30]) {
    State x, y;
    state_from(a, x);
    state_from(b, y);
    StateVec d = x.boxminus(y);
    for (int i = 0; i < 30; ++i) out30[i] = d[i];
}

void lko_boxplus(lk_state* a, const double delta30[30]) {
    State x;
    state_from(a, x);
    StateVec d;
    for (int i = 0; i < 30; ++i) d[i] = delta30[i];
    x.boxplus(d);
    state_to(x, a);
}

void lko_exp3(double v1, double v2, double v3, double out9[9]) {
    M3 m = Exp3(v1, v2, v3);
    for (int i = 0; i < 9; ++i) out9[i] = m.a[i];
}

void lko_log(const double R9[9], double out3[3]) {
    V3 v = Log(to_m3(R9));
    for (int k = 0; k < 3; ++k) out3[k] = v[k];
}

// Export the map as an lk_map blob. buf == NULL -> size query.
int lko_map_export(void* h, void* buf, size_t cap, size_t* bytes_out) {
    Ctx* c = (Ctx*)h;
    Exporter ex;
    size_t nroots = c->map->voxel_map_.size();
    ex.roots.reserve(nroots);
    ex.nodes.resize(nroots);
    ex.aux.resize(nroots);
    // deterministic order: sort keys
    std::vector<std::pair<std::array<int, 3>, const VoxelOctoTree*>> items;
    for (auto& kv : c->map->voxel_map_) items.push_back({kv.first, kv.second});
    std::sort(items.begin(), items.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    int idx = 0;
    for (auto& it : items) {
        lk_map_root r;
        for (int k = 0; k < 3; ++k) r.key[k] = it.first[k];
        r.node = idx;
        ex.roots.push_back(r);
        ex.fill(idx, it.second, -1);
        for (int k = 0; k < 3; ++k) ex.aux[idx].key[k] = it.first[k];
        ++idx;
    }
    lk_map_blob_header hd;
    std::memset(&hd, 0, sizeof(hd));
    hd.magic = LK_MAP_MAGIC;
    hd.version = 1;
    hd.n_roots = (uint32_t)ex.roots.size();
    hd.n_nodes = (uint32_t)ex.nodes.size();
    hd.n_points = ex.points.size();
    size_t bytes = sizeof(hd) + ex.roots.size() * sizeof(lk_map_root) + ex.nodes.size() * sizeof(lk_map_node) +
                   ex.aux.size() * sizeof(lk_map_aux) + ex.points.size() * sizeof(lk_map_point);
    if (bytes_out) *bytes_out = bytes;
    if (!buf) return 0;
    if (cap < bytes) return -1;
    char* p = (char*)buf;
    std::memcpy(p, &hd, sizeof(hd)); p += sizeof(hd);
    std::memcpy(p, ex.roots.data(), ex.roots.size() * sizeof(lk_map_root)); p += ex.roots.size() * sizeof(lk_map_root);
    std::memcpy(p, ex.nodes.data(), ex.nodes.size() * sizeof(lk_map_node)); p += ex.nodes.size() * sizeof(lk_map_node);
    std::memcpy(p, ex.aux.data(), ex.aux.size() * sizeof(lk_map_aux)); p += ex.aux.size() * sizeof(lk_map_aux);
    std::memcpy(p, ex.points.data(), ex.points.size() * sizeof(lk_map_point));
    return 0;
}

// Import a blob into the oracle's map (so GPU-built or analytic maps can be fed to the CPU path).
int lko_map_import(void* h, const void* buf, size_t bytes) {
    Ctx* c = (Ctx*)h;
    const char* p = (const char*)buf;
    lk_map_blob_header hd;
    if (bytes < sizeof(hd)) return -1;
    std::memcpy(&hd, p, sizeof(hd));
    if (hd.magic != LK_MAP_MAGIC) return -1;
    const lk_map_root* roots = (const lk_map_root*)(p + sizeof(hd));
    const lk_map_node* nodes = (const lk_map_node*)(roots + hd.n_roots);
    const lk_map_aux* aux = (const lk_map_aux*)(nodes + hd.n_nodes);
    const lk_map_point* pts = (const lk_map_point*)(aux + hd.n_nodes);
    for (auto& kv : c->map->voxel_map_) delete kv.second;
    c->map->voxel_map_.clear();
    const VoxelMapConfig& mc = c->map->config_setting_;
    struct Rec {
        static VoxelOctoTree* make(const VoxelMapConfig& mc, const lk_map_node* nodes, const lk_map_aux* aux,
                                   const lk_map_point* pts, int idx) {
            const lk_map_node& n = nodes[idx];
            const lk_map_aux& a = aux[idx];
            int layer = (n.flags >> LK_NODE_LAYER_SHIFT) & 0xff;
            VoxelOctoTree* t = new VoxelOctoTree(mc.max_layer_, layer, mc.layer_init_num_[layer], mc.max_points_num_,
                                                 (float)mc.planner_threshold_);
            t->layer_init_num_ = mc.layer_init_num_;
            for (int k = 0; k < 3; ++k) {
                t->voxel_center_[k] = a.voxel_center[k];
                t->plane_ptr_->center_[k] = n.center[k];
                t->plane_ptr_->normal_[k] = n.normal[k];
            }
            t->quater_length_ = a.quater_length;
            int q = 0;
            for (int i = 0; i < 6; ++i)
                for (int j = i; j < 6; ++j) {
                    t->plane_ptr_->plane_var_(i, j) = n.plane_var[q];
                    t->plane_ptr_->plane_var_(j, i) = n.plane_var[q];
                    ++q;
                }
            t->plane_ptr_->d_ = n.d;
            t->plane_ptr_->radius_ = n.radius;
            t->plane_ptr_->is_plane_ = (n.flags & LK_NODE_IS_PLANE) != 0;
            t->init_octo_ = (n.flags & LK_NODE_INIT_OCTO) != 0;
            t->update_enable_ = (n.flags & LK_NODE_UPDATE_ENABLE) != 0;
            t->new_points_ = a.new_points;
            t->temp_points_.resize(a.pts_count);
            for (int i = 0; i < a.pts_count; ++i) {
                const lk_map_point& mp = pts[a.pts_base + i];
                pointWithVar& pv = t->temp_points_[i];
                for (int k = 0; k < 3; ++k) pv.point_w[k] = mp.pw[k];
                pv.var(0, 0) = mp.var[0]; pv.var(0, 1) = mp.var[1]; pv.var(0, 2) = mp.var[2];
                pv.var(1, 0) = mp.var[1]; pv.var(1, 1) = mp.var[3]; pv.var(1, 2) = mp.var[4];
                pv.var(2, 0) = mp.var[2]; pv.var(2, 1) = mp.var[4]; pv.var(2, 2) = mp.var[5];
            }
            uint32_t mask = (n.flags >> LK_NODE_CHILDMASK_SHIFT) & 0xff;
            t->octo_state_ = (t->init_octo_ && !t->plane_ptr_->is_plane_ && layer < mc.max_layer_) ? 1 : 0;
            if (n.child_base >= 0)
                for (int i = 0; i < 8; ++i)
                    if (mask & (1u << i)) t->leaves_[i] = make(mc, nodes, aux, pts, n.child_base + i);
            return t;
        }
    };
    for (uint32_t r = 0; r < hd.n_roots; ++r) {
        std::array<int, 3> key = {roots[r].key[0], roots[r].key[1], roots[r].key[2]};
        c->map->voxel_map_[key] = Rec::make(mc, nodes, aux, pts, roots[r].node);
    }
    return 0;
}

uint64_t lko_map_num_roots(void* h) { return ((Ctx*)h)->map->voxel_map_.size(); }

// CPU baseline: `batch` independent scans (one bucket each, static map) on `nthreads` threads,
// each through KILO::predictUpdatePoint with update_map = false. Returns elapsed seconds of the
// update loop only (the region of the reference's "State predict/update & Map update" timer).
double lko_batch_run(void* h, int batch, const lk_state* x_in, const double* P_in, const lk_stream_clock* clk_in,
                     const float* pts4, const uint32_t* scan_offsets, const double* bucket_times, int iters,
                     int gain_mode, int nthreads, lk_state* x_out, double* P_out, uint32_t* n_eff_out) {
    Ctx* c = (Ctx*)h;
    if (nthreads < 1) nthreads = 1;
    {
        // The loop allocates tens of MB per scan (pointWithVar is 384 B, KILO.cc:120). With glibc's default thresholds
        // every such vector is an mmap / munmap pair, and 100+ threads then fight over the process's mm lock (a 5x
        // box-to-box swing of the all-core figure in round 1). Keep large blocks in the per-thread arenas instead.
        static bool tuned = false;
        if (!tuned) {
            mallopt(M_MMAP_THRESHOLD, 1 << 30);
            mallopt(M_TRIM_THRESHOLD, 1 << 30);
            mallopt(M_ARENA_MAX, 256);
            tuned = true;
        }
    }
    const long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    auto worker = [&](int tid) {
        if (nthreads > 1 && ncpu > 0) {  // one scan loop per hardware thread, pinned
            cpu_set_t set;
            CPU_ZERO(&set);
            CPU_SET((int)(tid % ncpu), &set);
            pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
        }
        KILO k(c->ec, c->map, c->extR, c->extT);
        k.gain_mode_ = gain_mode ? GAIN_INFORMATION : GAIN_LITERAL;
        k.iters_ = iters;
        k.update_map_ = false;
        std::memcpy(k.eskf_.Q.a, c->kilo->eskf_.Q.a, sizeof(double) * 900);
        for (int s = tid; s < batch; s += nthreads) {
            state_from(&x_in[s], k.eskf_.state);
            std::memcpy(k.eskf_.cov.a, P_in + (size_t)s * 900, sizeof(double) * 900);
            k.last_state_predict_time_ = clk_in[s].last_predict_time;
            k.last_state_update_time_ = clk_in[s].last_update_time;
            uint32_t n = scan_offsets[s + 1] - scan_offsets[s];
            std::vector<PointXYZT> body(n);
            std::memcpy(body.data(), pts4 + (size_t)scan_offsets[s] * 4, sizeof(float) * 4 * n);
            std::vector<PointXYZI> world(n);
            size_t succ = 0;
            k.predictUpdatePoint(bucket_times[s], 0, n, body, world, succ);
            if (x_out) state_to(k.eskf_.state, &x_out[s]);
            if (P_out) std::memcpy(P_out + (size_t)s * 900, k.eskf_.cov.a, sizeof(double) * 900);
            if (n_eff_out) n_eff_out[s] = (uint32_t)succ;
        }
    };
    auto t0 = std::chrono::steady_clock::now();
    if (nthreads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"

// ============================ what feeds the path ==========================
#include <climits>
extern "C" {

// LidarProcessing::{velodyne,ouster,hesai}Handler — legkilo/src/preprocess/lidar_processing.cc:25-108.
// Returns the number of points kept. pts_out = float4 (x, y, z, curvature).
uint32_t lko_decode_pointcloud2(const uint8_t* data, uint32_t n, const lk_pc2_layout* L, float blind, int filter_num,
                                double time_scale, float* pts_out, float* intensity_out, double* first_time,
                                double* last_time) {
    if (!n) return 0;
    auto f32 = [&](uint32_t i, uint32_t off) { float v; std::memcpy(&v, data + (size_t)i * L->point_step + off, 4); return v; };
    uint32_t m = 0;
    if (L->lidar_type == LK_LIDAR_HESAI) {
        auto ts = [&](uint32_t i) { double v; std::memcpy(&v, data + (size_t)i * L->point_step + L->off_time, 8); return v; };
        double first_point_time = time_scale * ts(0);
        double last_point_time = time_scale * ts(n - 1);
        if (first_time) *first_time = first_point_time;
        if (last_time) *last_time = last_point_time;
        for (uint32_t i = 0; i < n; ++i) {
            float x = f32(i, L->off_x), y = f32(i, L->off_y), z = f32(i, L->off_z);
            if ((i % filter_num) || (blind * blind > x * x + y * y + z * z)) continue;
            double cur_point_time = time_scale * ts(i);
            float curvature = std::round((cur_point_time - first_point_time) * 500.0f) / 500.0f;
            pts_out[4 * m] = x; pts_out[4 * m + 1] = y; pts_out[4 * m + 2] = z; pts_out[4 * m + 3] = curvature;
            if (intensity_out) intensity_out[m] = f32(i, L->off_intensity);
            ++m;
        }
        return m;
    }
    auto ts = [&](uint32_t i) -> double {
        if (L->lidar_type == LK_LIDAR_VELODYNE) return (double)f32(i, L->off_time);
        uint32_t v; std::memcpy(&v, data + (size_t)i * L->point_step + L->off_time, 4); return (double)v;
    };
    float first_point_time = time_scale * ts(0);
    float last_point_time = time_scale * ts(n - 1);
    if (first_time) *first_time = first_point_time;
    if (last_time) *last_time = last_point_time;
    for (uint32_t i = 0; i < n; ++i) {
        float x = f32(i, L->off_x), y = f32(i, L->off_y), z = f32(i, L->off_z);
        if ((i % filter_num) || (blind * blind > x * x + y * y + z * z)) continue;
        float cur_point_time = time_scale * ts(i);
        float curvature = std::round((cur_point_time - first_point_time) * 500.0f) / 500.0f;
        pts_out[4 * m] = x; pts_out[4 * m + 1] = y; pts_out[4 * m + 2] = z; pts_out[4 * m + 3] = curvature;
        if (intensity_out) intensity_out[m] = f32(i, L->off_intensity);
        ++m;
    }
    return m;
}

// pcl::VoxelGrid<PointXYZINormal>::applyFilter (PCL 1.8 voxel_grid.hpp; the library is absent, its
// published algorithm is restated: leaf index = floor(p * inverse_leaf) - min_b, dot divb_mul; centroid
// of x, y, z and curvature as float sums / float(n); leaves in ascending index) as KILO::process calls
// it (KILO.cc:356-360), then the sort by curvature (stable: canonical order) and the equal-curvature runs
// (KILO.cc:370-378). Returns -1 when the leaf size is too small for the extent (PCL only warns).
int lko_preprocess_scan(const float* in, uint32_t n, float leaf, float* out, uint32_t* n_out, uint32_t* bucket_offsets,
                        float* bucket_curv, uint32_t* n_buckets) {
    *n_out = 0; *n_buckets = 0; bucket_offsets[0] = 0;
    if (!n) return 0;
    float inv = 1.0f / leaf;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t i = 0; i < n; ++i) {
        const float* p = in + 4 * i;
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        for (int k = 0; k < 3; ++k) { mn[k] = std::min(mn[k], p[k]); mx[k] = std::max(mx[k], p[k]); }
    }
    int min_b[3], max_b[3], div_b[3];
    for (int k = 0; k < 3; ++k) {
        min_b[k] = (int)std::floor(mn[k] * inv);
        max_b[k] = (int)std::floor(mx[k] * inv);
        div_b[k] = max_b[k] - min_b[k] + 1;
    }
    if ((long long)div_b[0] * div_b[1] * div_b[2] > (long long)INT_MAX) return -1;
    const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
    std::vector<std::pair<int, uint32_t>> idx;
    idx.reserve(n);
    for (uint32_t i = 0; i < n; ++i) {
        const float* p = in + 4 * i;
        if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue;
        int i0 = (int)std::floor(p[0] * inv) - min_b[0], i1 = (int)std::floor(p[1] * inv) - min_b[1],
            i2 = (int)std::floor(p[2] * inv) - min_b[2];
        idx.push_back({i0 * mul[0] + i1 * mul[1] + i2 * mul[2], i});
    }
    std::stable_sort(idx.begin(), idx.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    std::vector<std::array<float, 4>> cent;
    for (size_t a = 0; a < idx.size();) {
        size_t b = a;
        float s[4] = {0, 0, 0, 0};
        while (b < idx.size() && idx[b].first == idx[a].first) {
            for (int k = 0; k < 4; ++k) s[k] += in[4 * idx[b].second + k];
            ++b;
        }
        float fn = (float)(b - a);
        cent.push_back({s[0] / fn, s[1] / fn, s[2] / fn, s[3] / fn});
        a = b;
    }
    std::stable_sort(cent.begin(), cent.end(), [](const auto& a, const auto& b) { return a[3] < b[3]; });
    uint32_t nb = 0;
    for (size_t i = 0; i < cent.size(); ++i) {
        std::memcpy(out + 4 * i, cent[i].data(), 16);
        if (i == 0 || cent[i][3] != cent[i - 1][3]) { bucket_offsets[nb] = (uint32_t)i; bucket_curv[nb] = cent[i][3]; ++nb; }
    }
    bucket_offsets[nb] = (uint32_t)cent.size();
    *n_out = (uint32_t)cent.size();
    *n_buckets = nb;
    return 0;
}

}  // extern "C"
