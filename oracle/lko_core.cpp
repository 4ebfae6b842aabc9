// oracle/lko_core.cpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md and lko_core.hpp).
// CPU restatement of the reference hot path. Citations are to /root/reference/legkilo/src/...
#include "lko_core.hpp"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>

namespace lko {

// ============================ common/math_utils.hpp ============================================

M3 skew(const V3& v) {  // math_utils.hpp:13-17
    M3 m;
    m(0, 0) = 0.0;   m(0, 1) = -v[2]; m(0, 2) = v[1];
    m(1, 0) = v[2];  m(1, 1) = 0.0;   m(1, 2) = -v[0];
    m(2, 0) = -v[1]; m(2, 1) = v[0];  m(2, 2) = 0.0;
    return m;
}

static M3 rodrigues(const V3& axis, double ang) {
    M3 K = skew(axis);
    // "Eye3 + sin*K + (1-cos)*K*K", evaluated left to right as in the reference expressions
    return M3::Identity() + std::sin(ang) * K + ((1.0 - std::cos(ang)) * K) * K;
}

M3 Exp_vec(const V3& ang) {  // math_utils.hpp:20-32
    double ang_norm = ang.norm();
    if (ang_norm > 0.0000001) {
        V3 r_axis = ang / ang_norm;
        return rodrigues(r_axis, ang_norm);
    }
    return M3::Identity();
}

M3 Exp_vel_dt(const V3& ang_vel, double dt) {  // math_utils.hpp:35-52
    double ang_vel_norm = ang_vel.norm();
    if (ang_vel_norm > 0.0000001) {
        V3 r_axis = ang_vel / ang_vel_norm;
        double r_ang = ang_vel_norm * dt;
        return rodrigues(r_axis, r_ang);
    }
    return M3::Identity();
}

M3 Exp3(double v1, double v2, double v3) {  // math_utils.hpp:55-68
    double norm = std::sqrt(v1 * v1 + v2 * v2 + v3 * v3);
    if (norm > 0.00001) {
        V3 r_ang = vec3(v1 / norm, v2 / norm, v3 / norm);
        return rodrigues(r_ang, norm);
    }
    return M3::Identity();
}

V3 Log(const M3& R) {  // math_utils.hpp:72-76
    double theta = (R.trace() > 3.0 - 1e-6) ? 0.0 : std::acos(0.5 * (R.trace() - 1));
    V3 K = vec3(R(2, 1) - R(1, 2), R(0, 2) - R(2, 0), R(1, 0) - R(0, 1));
    return (std::abs(theta) < 0.001) ? (0.5 * K) : (0.5 * theta / std::sin(theta) * K);
}

// ============================ core/slam/eskf.cc ================================================

State::State() {  // eskf.cc:5-16
    rot = M3::Identity();
    grav = vec3(0.0, 0.0, -9.81);
}

void State::boxplus(const StateVec& delta) {  // eskf.cc:18-29
    rot = rot * Exp3(delta[0], delta[1], delta[2]);
    V3* blocks[9] = {&pos, &vel, &ba, &bw, &grav, &imu_a, &imu_w, &bv, &contact};
    for (int b = 0; b < 9; ++b)
        for (int k = 0; k < 3; ++k) (*blocks[b])[k] += delta[3 + 3 * b + k];
}

StateVec State::boxminus(const State& other) const {  // eskf.cc:31-45
    StateVec delta;
    M3 rot_delta = other.rot.transpose() * rot;
    V3 l = Log(rot_delta);
    const V3* mine[9] = {&pos, &vel, &ba, &bw, &grav, &imu_a, &imu_w, &bv, &contact};
    const V3* theirs[9] = {&other.pos, &other.vel, &other.ba, &other.bw, &other.grav,
                           &other.imu_a, &other.imu_w, &other.bv, &other.contact};
    for (int k = 0; k < 3; ++k) delta[k] = l[k];
    for (int b = 0; b < 9; ++b)
        for (int k = 0; k < 3; ++k) delta[3 + 3 * b + k] = (*mine[b])[k] - (*theirs[b])[k];
    return delta;
}

void ESKF::initProcessCovQ() {  // eskf.cc:47-62
    Q = StateCov::Zero();
    const double d[7] = {config.vel_process_cov,      config.acc_bias_process_cov,
                         config.gyr_bias_process_cov, config.imu_acc_process_cov,
                         config.imu_gyr_process_cov,  config.kin_bias_process_cov,
                         config.contact_process_cov};
    const int at[7] = {6, 9, 12, 18, 21, 24, 27};
    for (int b = 0; b < 7; ++b)
        for (int k = 0; k < 3; ++k) Q(at[b] + k, at[b] + k) = d[b];
}

StateVec ESKF::getFunctionf(double dt) const {  // eskf.cc:64-70
    StateVec vec;
    V3 w = dt * state.imu_w;
    V3 v = dt * state.vel;
    V3 a = dt * (state.rot * state.imu_a + state.grav);
    for (int k = 0; k < 3; ++k) {
        vec[k] = w[k];
        vec[3 + k] = v[k];
        vec[6 + k] = a[k];
    }
    return vec;
}

StateCov ESKF::getFx(double dt) const {  // eskf.cc:72-81
    StateCov Fx = StateCov::Identity();
    Fx.setBlock<3, 3>(0, 0, Exp_vec((-dt) * state.imu_w));
    Fx.setBlock<3, 3>(0, 21, dt * M3::Identity());
    Fx.setBlock<3, 3>(3, 6, dt * M3::Identity());
    Fx.setBlock<3, 3>(6, 0, ((-dt) * state.rot) * skew(state.imu_a));
    Fx.setBlock<3, 3>(6, 15, dt * M3::Identity());
    Fx.setBlock<3, 3>(6, 18, dt * state.rot);
    return Fx;
}

void ESKF::predict(double dt, bool prop_state, bool prop_cov) {  // eskf.cc:83-89
    if (prop_state) state.boxplus(getFunctionf(dt));
    if (prop_cov) {
        StateCov Fx = getFx(dt);
        cov = (Fx * cov) * Fx.transpose() + (dt * dt) * Q;
    }
}

// eskf.cc:91-113. GAIN_LITERAL mirrors the measurement-space form (N x N system, partial-pivot
// LU as MatrixXd::inverse() does; K = PHT * S^-1 obtained by solving S^T K^T = PHT^T rather than
// forming S^-1 explicitly). GAIN_INFORMATION is the algebraically identical 6x6 form
//  that the device path uses.
PointGain ESKF::pointGain(const ObsPoints& obs, GainMode mode) const {
    const int N = obs.n();
    PointGain g;
    Mat<DIM, 6> P6 = cov.block<DIM, 6>(0, 0);
    if (mode == GAIN_LITERAL) {
        if (N == 1) {  // eskf.cc:98-104
            Mat<6, 1> ht;
            for (int j = 0; j < 6; ++j) ht[j] = obs.h[j];
            Mat<DIM, 1> PHT = P6 * ht;
            double hPh = 0;
            {
                double s = obs.h[0] * PHT[0];
                for (int j = 1; j < 6; ++j) s += obs.h[j] * PHT[j];
                hPh = s;
            }
            double HPHT_R_inv = 1 / (0.0001 + hPh + obs.R[0]);
            Mat<DIM, 1> K = HPHT_R_inv * PHT;
            g.delta = K * obs.z[0];
            for (int i = 0; i < DIM; ++i)
                for (int j = 0; j < 6; ++j) g.KH(i, j) = K[i] * obs.h[j];
            return g;
        }
        // PHT = P[:,0:6] * h^T   (30 x N)
        DMat PHT(DIM, N);
        for (int i = 0; i < DIM; ++i)
            for (int k = 0; k < N; ++k) {
                double s = P6(i, 0) * obs.h[(size_t)k * 6 + 0];
                for (int j = 1; j < 6; ++j) s += P6(i, j) * obs.h[(size_t)k * 6 + j];
                PHT(i, k) = s;
            }
        // HPHT_R = h * PHT.topRows(6) + diag(r)   (N x N); we factor its transpose
        DMat St(N, N);
        for (int a = 0; a < N; ++a)
            for (int b = 0; b < N; ++b) {
                double s = obs.h[(size_t)a * 6 + 0] * PHT(0, b);
                for (int j = 1; j < 6; ++j) s += obs.h[(size_t)a * 6 + j] * PHT(j, b);
                if (a == b) s += obs.R[a];
                St(b, a) = s;
            }
        std::vector<int> perm;
        lu_factor(St, perm);
        DMat Kt(N, DIM);
        for (int k = 0; k < N; ++k)
            for (int i = 0; i < DIM; ++i) Kt(k, i) = PHT(i, k);
        lu_solve(St, perm, Kt);  // Kt = (S^T)^-1 PHT^T  =>  K = PHT S^-1
        for (int i = 0; i < DIM; ++i) {
            double s = 0;
            for (int k = 0; k < N; ++k) s += Kt(k, i) * obs.z[k];
            g.delta[i] = s;
            for (int j = 0; j < 6; ++j) {
                double t = 0;
                for (int k = 0; k < N; ++k) t += Kt(k, i) * obs.h[(size_t)k * 6 + j];
                g.KH(i, j) = t;
            }
        }
        return g;
    }
    // information form: A = sum h^T h / R, b = sum h^T z / R, M = I + A P66,
    // delta = P6 M^-1 b, KH = P6 M^-1 A. For N == 1 the reference adds 1e-4 to S.
    M6 A;
    V6 b;
    for (int k = 0; k < N; ++k) {
        double r = obs.R[k];
        if (N == 1) r += 0.0001;
        double w = 1.0 / r;
        for (int i = 0; i < 6; ++i) {
            double hw = obs.h[(size_t)k * 6 + i] * w;
            b[i] += hw * obs.z[k];
            for (int j = 0; j < 6; ++j) A(i, j) += hw * obs.h[(size_t)k * 6 + j];
        }
    }
    M6 P66 = cov.block<6, 6>(0, 0);
    M6 Mx = M6::Identity() + A * P66;
    DMat Mlu(6, 6), rhs(6, 7);
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) {
            Mlu(i, j) = Mx(i, j);
            rhs(i, 1 + j) = A(i, j);
        }
        rhs(i, 0) = b[i];
    }
    std::vector<int> perm;
    lu_factor(Mlu, perm);
    lu_solve(Mlu, perm, rhs);
    V6 y;
    M6 W;
    for (int i = 0; i < 6; ++i) {
        y[i] = rhs(i, 0);
        for (int j = 0; j < 6; ++j) W(i, j) = rhs(i, 1 + j);
    }
    g.delta = P6 * y;
    g.KH = P6 * W;
    return g;
}

void ESKF::applyPointGain(const PointGain& g, bool update_state, bool update_cov) {
    if (update_state) state.boxplus(g.delta);
    if (update_cov) {
        Mat<6, DIM> P6r = cov.block<6, DIM>(0, 0);
        cov = cov - g.KH * P6r;  // eskf.cc:103 / :112 (no symmetrisation)
    }
}

static void solve_gain(int n_state, int m, const DMat& PHT, const DMat& S, DMat& K) {
    // K = PHT * S^-1 via S^T K^T = PHT^T (partial-pivot LU == Eigen's inverse() for m > 4)
    DMat St(m, m);
    for (int a = 0; a < m; ++a)
        for (int b = 0; b < m; ++b) St(b, a) = S(a, b);
    std::vector<int> perm;
    lu_factor(St, perm);
    DMat Kt(m, n_state);
    for (int k = 0; k < m; ++k)
        for (int i = 0; i < n_state; ++i) Kt(k, i) = PHT(i, k);
    lu_solve(St, perm, Kt);
    K = DMat(n_state, m);
    for (int i = 0; i < n_state; ++i)
        for (int k = 0; k < m; ++k) K(i, k) = Kt(k, i);
}

void ESKF::updateByImu(const double z[6], const double R[6]) {  // eskf.cc:125-135
    DMat PHT(DIM, 6), HP(6, DIM), HPHT(6, 6), K;
    for (int i = 0; i < DIM; ++i)
        for (int j = 0; j < 6; ++j) {
            PHT(i, j) = cov(i, 9 + j) + cov(i, 18 + j);
            HP(j, i) = cov(9 + j, i) + cov(18 + j, i);
        }
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) HPHT(i, j) = PHT(9 + i, j) + PHT(18 + i, j);
    for (int i = 0; i < 6; ++i) HPHT(i, i) += R[i];
    solve_gain(DIM, 6, PHT, HPHT, K);
    StateVec delta_x;
    for (int i = 0; i < DIM; ++i) {
        double s = K(i, 0) * z[0];
        for (int j = 1; j < 6; ++j) s += K(i, j) * z[j];
        delta_x[i] = s;
    }
    state.boxplus(delta_x);
    StateCov KHP;
    for (int i = 0; i < DIM; ++i)
        for (int j = 0; j < DIM; ++j) {
            double s = K(i, 0) * HP(0, j);
            for (int k = 1; k < 6; ++k) s += K(i, k) * HP(k, j);
            KHP(i, j) = s;
        }
    cov = cov - KHP;
}

void ESKF::updateByKinImu(int m, const std::vector<double>& H, const std::vector<double>& z,
                          const std::vector<double>& R) {  // eskf.cc:137-145
    DMat PHT(DIM, m), HPHT(m, m), K;
    for (int i = 0; i < DIM; ++i)
        for (int a = 0; a < m; ++a) {
            double s = 0;
            for (int k = 0; k < DIM; ++k) s += cov(i, k) * H[(size_t)a * DIM + k];
            PHT(i, a) = s;
        }
    for (int a = 0; a < m; ++a)
        for (int b = 0; b < m; ++b) {
            double s = 0;
            for (int k = 0; k < DIM; ++k) s += H[(size_t)a * DIM + k] * PHT(k, b);
            HPHT(a, b) = s;
        }
    for (int a = 0; a < m; ++a) HPHT(a, a) += R[a];
    solve_gain(DIM, m, PHT, HPHT, K);
    StateVec delta_x;
    for (int i = 0; i < DIM; ++i) {
        double s = 0;
        for (int a = 0; a < m; ++a) s += K(i, a) * z[a];
        delta_x[i] = s;
    }
    state.boxplus(delta_x);
    // cov = cov - K * H * cov   ((K*H)*cov)
    StateCov KH, out;
    for (int i = 0; i < DIM; ++i)
        for (int j = 0; j < DIM; ++j) {
            double s = 0;
            for (int a = 0; a < m; ++a) s += K(i, a) * H[(size_t)a * DIM + j];
            KH(i, j) = s;
        }
    out = cov - KH * cov;
    cov = out;
}

// ============================ core/slam/voxel_map.cc ===========================================

// PCL's DEG2RAD macro (pcl/pcl_macros.h), which voxel_map.cc:27 picks up: ((x)*0.017453293).
static inline double DEG2RAD(double x) { return x * 0.017453293; }

void calcBodyCov(V3& pb, const float range_inc, const float degree_inc, M3& cov) {  // :22-40
    if (pb[2] == 0) pb[2] = 0.0001;
    float range = std::sqrt(pb[0] * pb[0] + pb[1] * pb[1] + pb[2] * pb[2]);
    float range_var = range_inc * range_inc;
    double dv = std::pow(std::sin(DEG2RAD(degree_inc)), 2);
    Mat<2, 2> direction_var;
    direction_var(0, 0) = dv;
    direction_var(1, 1) = dv;
    V3 direction = normalized(pb);
    M3 direction_hat = skew(direction);
    V3 base_vector1 = vec3(1, 1, -(direction[0] + direction[1]) / direction[2]);
    base_vector1 = normalized(base_vector1);
    V3 base_vector2 = cross(base_vector1, direction);
    base_vector2 = normalized(base_vector2);
    Mat<3, 2> N;
    for (int k = 0; k < 3; ++k) {
        N(k, 0) = base_vector1[k];
        N(k, 1) = base_vector2[k];
    }
    Mat<3, 2> A = ((double)range * direction_hat) * N;
    cov = (direction * (double)range_var) * direction.transpose() + (A * direction_var) * A.transpose();
}

VoxelOctoTree::VoxelOctoTree(int max_layer, int layer, int points_size_threshold, int max_points_num,
                             float planer_threshold)  // voxel_map.h:149-163
    : layer_(layer),
      planer_threshold_(planer_threshold),
      points_size_threshold_(points_size_threshold),
      max_points_num_(max_points_num),
      max_layer_(max_layer) {
    octo_state_ = 0;
    new_points_ = 0;
    update_size_threshold_ = 5;
    init_octo_ = false;
    update_enable_ = true;
    for (int i = 0; i < 8; i++) leaves_[i] = nullptr;
    plane_ptr_ = new VoxelPlane;
    voxel_center_[0] = voxel_center_[1] = voxel_center_[2] = 0;
    quater_length_ = 0;
}

VoxelOctoTree::~VoxelOctoTree() {
    for (int i = 0; i < 8; i++) delete leaves_[i];
    delete plane_ptr_;
}

void VoxelOctoTree::init_plane(const std::vector<pointWithVar>& points, VoxelPlane* plane) {  // :42-117
    plane->plane_var_ = M6::Zero();
    plane->covariance_ = M3::Zero();
    plane->center_ = V3::Zero();
    plane->normal_ = V3::Zero();
    plane->points_size_ = (int)points.size();
    plane->radius_ = 0;
    for (const auto& pv : points) {
        plane->covariance_ += pv.point_w * pv.point_w.transpose();
        plane->center_ += pv.point_w;
    }
    plane->center_ = plane->center_ / plane->points_size_;
    plane->covariance_ = plane->covariance_ / plane->points_size_ - plane->center_ * plane->center_.transpose();
    double evalsReal[3];
    M3 evecs;
    eig_sym3(plane->covariance_, evalsReal, evecs);
    int evalsMin = 0, evalsMax = 0;  // minCoeff / maxCoeff return the FIRST extremum
    for (int i = 1; i < 3; ++i) {
        if (evalsReal[i] < evalsReal[evalsMin]) evalsMin = i;
        if (evalsReal[i] > evalsReal[evalsMax]) evalsMax = i;
    }
    int evalsMid = 3 - evalsMin - evalsMax;
    if (evalsMid > 2) evalsMid = 2;  // all-equal eigenvalues: UB in the reference (:63); pick a slot
    auto col = [&](int c) { return vec3(evecs(0, c), evecs(1, c), evecs(2, c)); };
    M3 J_Q;
    for (int i = 0; i < 3; ++i) J_Q(i, i) = 1.0 / plane->points_size_;
    if (evalsReal[evalsMin] < planer_threshold_) {
        for (size_t i = 0; i < points.size(); i++) {
            Mat<6, 3> J;
            M3 F;
            for (int m = 0; m < 3; m++) {
                if (m != evalsMin) {
                    Mat<1, 3> lhs = (points[i].point_w - plane->center_).transpose() /
                                    ((plane->points_size_) * (evalsReal[evalsMin] - evalsReal[m]));
                    M3 sym = col(m) * col(evalsMin).transpose() + col(evalsMin) * col(m).transpose();
                    Mat<1, 3> F_m = lhs * sym;
                    for (int k = 0; k < 3; ++k) F(m, k) = F_m(0, k);
                } else {
                    for (int k = 0; k < 3; ++k) F(m, k) = 0;
                }
            }
            J.setBlock<3, 3>(0, 0, evecs * F);
            J.setBlock<3, 3>(3, 0, J_Q);
            plane->plane_var_ += (J * points[i].var) * J.transpose();
        }
        plane->normal_ = col(evalsMin);
        plane->y_normal_ = col(evalsMid);
        plane->x_normal_ = col(evalsMax);
        plane->min_eigen_value_ = evalsReal[evalsMin];
        plane->mid_eigen_value_ = evalsReal[evalsMid];
        plane->max_eigen_value_ = evalsReal[evalsMax];
        plane->radius_ = std::sqrt(evalsReal[evalsMax]);
        plane->d_ = -(plane->normal_[0] * plane->center_[0] + plane->normal_[1] * plane->center_[1] +
                      plane->normal_[2] * plane->center_[2]);
        plane->is_plane_ = true;
        plane->is_update_ = true;
        if (!plane->is_init_) plane->is_init_ = true;
    } else {
        plane->is_update_ = true;
        plane->is_plane_ = false;
    }
}

void VoxelOctoTree::init_octo_tree() {  // :119-137
    if (temp_points_.size() > (size_t)points_size_threshold_) {
        init_plane(temp_points_, plane_ptr_);
        if (plane_ptr_->is_plane_ == true) {
            octo_state_ = 0;
            if (temp_points_.size() > (size_t)max_points_num_) {
                update_enable_ = false;
                std::vector<pointWithVar>().swap(temp_points_);
                new_points_ = 0;
            }
        } else {
            octo_state_ = 1;
            cut_octo_tree();
        }
        init_octo_ = true;
        new_points_ = 0;
    }
}

static VoxelOctoTree* make_child(VoxelOctoTree* parent, const int xyz[3]) {  // :151-157, :220-226
    VoxelOctoTree* c = new VoxelOctoTree(parent->max_layer_, parent->layer_ + 1,
                                         parent->layer_init_num_[parent->layer_ + 1],
                                         parent->max_points_num_, parent->planer_threshold_);
    c->layer_init_num_ = parent->layer_init_num_;
    for (int k = 0; k < 3; ++k)
        c->voxel_center_[k] = parent->voxel_center_[k] + (2 * xyz[k] - 1) * parent->quater_length_;
    c->quater_length_ = parent->quater_length_ / 2;
    return c;
}

void VoxelOctoTree::cut_octo_tree() {  // :139-183
    if (layer_ >= max_layer_) {
        octo_state_ = 0;
        return;
    }
    for (size_t i = 0; i < temp_points_.size(); i++) {
        int xyz[3] = {0, 0, 0};
        for (int k = 0; k < 3; ++k)
            if (temp_points_[i].point_w[k] > voxel_center_[k]) xyz[k] = 1;
        int leafnum = 4 * xyz[0] + 2 * xyz[1] + xyz[2];
        if (leaves_[leafnum] == nullptr) leaves_[leafnum] = make_child(this, xyz);
        leaves_[leafnum]->temp_points_.push_back(temp_points_[i]);
        leaves_[leafnum]->new_points_++;
    }
    for (unsigned i = 0; i < 8; i++) {
        if (leaves_[i] != nullptr) {
            if (leaves_[i]->temp_points_.size() > (size_t)leaves_[i]->points_size_threshold_) {
                init_plane(leaves_[i]->temp_points_, leaves_[i]->plane_ptr_);
                if (leaves_[i]->plane_ptr_->is_plane_) {
                    leaves_[i]->octo_state_ = 0;
                    if (leaves_[i]->temp_points_.size() > (size_t)leaves_[i]->max_points_num_) {
                        leaves_[i]->update_enable_ = false;
                        std::vector<pointWithVar>().swap(leaves_[i]->temp_points_);
                        new_points_ = 0;  // resets the PARENT's counter (:172) — harmless quirk
                    }
                } else {
                    leaves_[i]->octo_state_ = 1;
                    leaves_[i]->cut_octo_tree();
                }
                leaves_[i]->init_octo_ = true;
                leaves_[i]->new_points_ = 0;
            }
        }
    }
}

void VoxelOctoTree::UpdateOctoTree(const pointWithVar& pv) {  // :185-241
    if (!init_octo_) {
        new_points_++;
        temp_points_.push_back(pv);
        if (temp_points_.size() > (size_t)points_size_threshold_) init_octo_tree();
    } else {
        if (plane_ptr_->is_plane_) {
            if (update_enable_) {
                new_points_++;
                temp_points_.push_back(pv);
                if (new_points_ > update_size_threshold_) {
                    init_plane(temp_points_, plane_ptr_);
                    new_points_ = 0;
                }
                if (temp_points_.size() >= (size_t)max_points_num_) {
                    update_enable_ = false;
                    std::vector<pointWithVar>().swap(temp_points_);
                    new_points_ = 0;
                }
            }
        } else {
            if (layer_ < max_layer_) {
                int xyz[3] = {0, 0, 0};
                for (int k = 0; k < 3; ++k)
                    if (pv.point_w[k] > voxel_center_[k]) xyz[k] = 1;
                int leafnum = 4 * xyz[0] + 2 * xyz[1] + xyz[2];
                if (leaves_[leafnum] == nullptr) leaves_[leafnum] = make_child(this, xyz);
                leaves_[leafnum]->UpdateOctoTree(pv);
            } else {
                if (update_enable_) {
                    new_points_++;
                    temp_points_.push_back(pv);
                    if (new_points_ > update_size_threshold_) {
                        init_plane(temp_points_, plane_ptr_);
                        new_points_ = 0;
                    }
                    if (temp_points_.size() > (size_t)max_points_num_) {
                        update_enable_ = false;
                        std::vector<pointWithVar>().swap(temp_points_);
                        new_points_ = 0;
                    }
                }
            }
        }
    }
}

size_t KeyHash::operator()(const std::array<int, 3>& v) const {
    // The reference hash (eigen_types.hpp:80-82) only places keys in buckets; results never
    // depend on it (exact-match map). Same Teschner primes, evaluated in unsigned arithmetic.
    return size_t(((uint32_t)v[0] * 73856093u) ^ ((uint32_t)v[1] * 471943u) ^ ((uint32_t)v[2] * 83492791u));
}

VoxelMapManager::~VoxelMapManager() {
    for (auto& kv : voxel_map_) delete kv.second;
}

static std::array<int, 3> voxelKeyFloor(const V3& pt, double voxel_size) {  // eigen_types.hpp:89-95
    return {static_cast<int>(std::floor(pt[0] / voxel_size)), static_cast<int>(std::floor(pt[1] / voxel_size)),
            static_cast<int>(std::floor(pt[2] / voxel_size))};
}

static VoxelOctoTree* make_root(const VoxelMapConfig& c, const std::array<int, 3>& position, float voxel_size) {
    float planer_threshold = c.planner_threshold_;
    VoxelOctoTree* octo_tree =
        new VoxelOctoTree(c.max_layer_, 0, c.layer_init_num_[0], c.max_points_num_, planer_threshold);
    octo_tree->quater_length_ = voxel_size / 4;
    for (int k = 0; k < 3; ++k) octo_tree->voxel_center_[k] = (0.5 + position[k]) * voxel_size;
    octo_tree->layer_init_num_ = c.layer_init_num_;
    return octo_tree;
}

void VoxelMapManager::BuildVoxelMap(const float* xyz_world, const float* xyz_body, size_t n, const M3& rot,
                                    const M3& rot_cov, const M3& pos_cov) {  // :287-334
    float voxel_size = config_setting_.max_voxel_size_;
    std::vector<pointWithVar> input_points;
    input_points.reserve(n);
    for (size_t i = 0; i < n; i++) {
        pointWithVar pv;
        pv.point_w = vec3(xyz_world[3 * i], xyz_world[3 * i + 1], xyz_world[3 * i + 2]);
        V3 point_this = vec3(xyz_body[3 * i], xyz_body[3 * i + 1], xyz_body[3 * i + 2]);
        M3 var;
        calcBodyCov(point_this, config_setting_.dept_err_, config_setting_.beam_err_, var);
        M3 point_crossmat = skew(point_this);
        M3 rE = rot * extR_;
        var = (rE * var) * rE.transpose() + ((-point_crossmat) * rot_cov) * (-point_crossmat).transpose() + pos_cov;
        pv.var = var;
        input_points.push_back(pv);
    }
    for (size_t i = 0; i < n; i++) {
        const pointWithVar& p_v = input_points[i];
        std::array<int, 3> position = voxelKeyFloor(p_v.point_w, voxel_size);
        auto iter = voxel_map_.find(position);
        if (iter != voxel_map_.end()) {
            iter->second->temp_points_.push_back(p_v);
            iter->second->new_points_++;
        } else {
            VoxelOctoTree* octo_tree = make_root(config_setting_, position, voxel_size);
            voxel_map_[position] = octo_tree;
            octo_tree->temp_points_.push_back(p_v);
            octo_tree->new_points_++;
        }
    }
    for (auto iter = voxel_map_.begin(); iter != voxel_map_.end(); ++iter) iter->second->init_octo_tree();
}

void VoxelMapManager::UpdateVoxelMap(const std::vector<pointWithVar>& input_points) {  // :336-361
    float voxel_size = config_setting_.max_voxel_size_;
    for (size_t i = 0; i < input_points.size(); i++) {
        const pointWithVar& p_v = input_points[i];
        std::array<int, 3> position = voxelKeyFloor(p_v.point_w, voxel_size);
        auto iter = voxel_map_.find(position);
        if (iter != voxel_map_.end()) {
            iter->second->UpdateOctoTree(p_v);
        } else {
            VoxelOctoTree* octo_tree = make_root(config_setting_, position, voxel_size);
            voxel_map_[position] = octo_tree;
            octo_tree->UpdateOctoTree(p_v);
        }
    }
}

void VoxelMapManager::build_single_residual(pointWithVar& pv, const VoxelOctoTree* current_octo,
                                            const int current_layer, bool& is_success, double& prob,
                                            PointToPlane& single_ptpl) const {  // :363-427
    int max_layer = config_setting_.max_layer_;
    double sigma_num = config_setting_.sigma_num_;
    double radius_k = 3;
    V3 p_w = pv.point_w;
    if (current_octo->plane_ptr_->is_plane_) {
        VoxelPlane& plane = *current_octo->plane_ptr_;
        float dis_to_plane = std::fabs(plane.normal_[0] * p_w[0] + plane.normal_[1] * p_w[1] +
                                       plane.normal_[2] * p_w[2] + plane.d_);
        float dis_to_center = (plane.center_[0] - p_w[0]) * (plane.center_[0] - p_w[0]) +
                              (plane.center_[1] - p_w[1]) * (plane.center_[1] - p_w[1]) +
                              (plane.center_[2] - p_w[2]) * (plane.center_[2] - p_w[2]);
        float range_dis = std::sqrt(dis_to_center - dis_to_plane * dis_to_plane);  // float arithmetic
        if (range_dis <= radius_k * plane.radius_) {
            Mat<1, 6> J_nq;
            for (int k = 0; k < 3; ++k) {
                J_nq(0, k) = p_w[k] - plane.center_[k];
                J_nq(0, 3 + k) = -plane.normal_[k];
            }
            double sigma_l = ((J_nq * plane.plane_var_) * J_nq.transpose())(0, 0);
            sigma_l += ((plane.normal_.transpose() * pv.var) * plane.normal_)(0, 0);
            if (dis_to_plane < sigma_num * std::sqrt(sigma_l)) {
                is_success = true;
                double this_prob = 1.0 / (std::sqrt(sigma_l)) * std::exp(-0.5 * dis_to_plane * dis_to_plane / sigma_l);
                if (this_prob > prob) {
                    prob = this_prob;
                    pv.normal = plane.normal_;
                    single_ptpl.body_cov_ = pv.body_var;
                    single_ptpl.point_b_ = pv.point_b;
                    single_ptpl.point_w_ = pv.point_w;
                    single_ptpl.plane_var_ = plane.plane_var_;
                    single_ptpl.normal_ = plane.normal_;
                    single_ptpl.center_ = plane.center_;
                    single_ptpl.d_ = plane.d_;
                    single_ptpl.layer_ = current_layer;
                    single_ptpl.dis_to_plane_ = plane.normal_[0] * p_w[0] + plane.normal_[1] * p_w[1] +
                                                plane.normal_[2] * p_w[2] + plane.d_;
                    single_ptpl.point_crossmat_ = pv.point_crossmat;
                }
                return;
            }
            return;
        }
        return;
    } else {
        if (current_layer < max_layer) {
            for (size_t leafnum = 0; leafnum < 8; leafnum++) {
                if (current_octo->leaves_[leafnum] != nullptr) {
                    build_single_residual(pv, current_octo->leaves_[leafnum], current_layer + 1, is_success, prob,
                                          single_ptpl);
                }
            }
        }
        return;
    }
}

// ============================ core/slam/KILO.cc ================================================

KILO::KILO(const EskfConfig& ec, std::shared_ptr<VoxelMapManager> map, const M3& ext_rot, const V3& ext_t)
    : eskf_(ec), map_manager_(std::move(map)), ext_rot_(ext_rot), ext_t_(ext_t) {
    map_manager_->extR_ = ext_rot;  // KILO.cc:78-79
    map_manager_->extT_ = ext_t;
    eskf_.initProcessCovQ();
}

bool KILO::predictUpdatePoint(double current_time, size_t idx_i, size_t idx_j,
                              const std::vector<PointXYZT>& cloud_down_body,
                              std::vector<PointXYZI>& cloud_down_world, size_t& success_pts_size_out,
                              BucketDebug* dbg) {  // KILO.cc:108-233
    // 1) Predict state (:110-115)
    double dt_cov = current_time - last_state_update_time_;
    eskf_.predict(dt_cov, false, true);
    double dt = current_time - last_state_predict_time_;
    eskf_.predict(dt, true, false);
    last_state_predict_time_ = current_time;

    size_t points_size = idx_j - idx_i;
    std::vector<pointWithVar> pv_list(points_size);
    std::vector<PointToPlane> ptpl_list;
    bool eskf_update = false;
    const double voxel = map_manager_->config_setting_.max_voxel_size_;
    if (dbg) {
        dbg->ok.assign(points_size, 0);
        dbg->h.assign(points_size * 6, 0.0);
        dbg->z.assign(points_size, 0.0);
        dbg->R.assign(points_size, 0.0);
        dbg->key.assign(points_size * 3, 0);
    }

    for (int it = 0; it < iters_; ++it) {
        // 2) Residuals (:117-184). it > 0 re-linearises at the current estimate with the
        //    predicted covariance held; it == 0, iters_ == 1 is the reference.
        ptpl_list.clear();
        ptpl_list.reserve(points_size);
        std::vector<size_t> ptpl_src;
        size_t success_this_iter = 0;
        for (size_t i = 0; i < points_size; ++i) {
            const PointXYZT& cur_pt = cloud_down_body[i + idx_i];
            pointWithVar& cur_pt_var = pv_list[i];
            if (it == 0) {
                cur_pt_var.point_b = vec3(cur_pt.x, cur_pt.y, cur_pt.z);
                cur_pt_var.point_i = ext_rot_ * cur_pt_var.point_b + ext_t_;
            }
            cur_pt_var.point_w = eskf_.getRot() * cur_pt_var.point_i + eskf_.getPos();
            if (it == 0) {
                cloud_down_world[idx_i + i].x = cur_pt_var.point_w[0];
                cloud_down_world[idx_i + i].y = cur_pt_var.point_w[1];
                cloud_down_world[idx_i + i].z = cur_pt_var.point_w[2];
                cloud_down_world[idx_i + i].intensity = 0;
                calcBodyCov(cur_pt_var.point_b, map_manager_->config_setting_.dept_err_,
                            map_manager_->config_setting_.beam_err_, cur_pt_var.body_var);
                cur_pt_var.point_crossmat = skew(cur_pt_var.point_i);
            }
            M3 rot_extR = eskf_.getRot() * ext_rot_;
            M3 rot_crossmat = eskf_.getRot() * cur_pt_var.point_crossmat;
            cur_pt_var.var = (rot_extR * cur_pt_var.body_var) * rot_extR.transpose() +
                             (rot_crossmat * eskf_.getRotCov()) * rot_crossmat.transpose() + eskf_.getPosCov();

            // 2.2 residual (:142-183)
            float loc_xyz[3];
            for (int j = 0; j < 3; j++) {
                loc_xyz[j] = cur_pt_var.point_w[j] / voxel;
                if (loc_xyz[j] < 0) loc_xyz[j] -= 1.0;
            }
            std::array<int, 3> position = {(int)loc_xyz[0], (int)loc_xyz[1], (int)loc_xyz[2]};
            if (dbg && it == iters_ - 1)
                for (int j = 0; j < 3; ++j) dbg->key[i * 3 + j] = position[j];
            auto iter = map_manager_->voxel_map_.find(position);
            if (iter != map_manager_->voxel_map_.end()) {
                VoxelOctoTree* current_octo = iter->second;
                PointToPlane single_ptpl;
                bool is_success = false;
                double prob = 0;
                map_manager_->build_single_residual(cur_pt_var, current_octo, 0, is_success, prob, single_ptpl);
                if (!is_success) {
                    std::array<int, 3> near_position = position;
                    for (int j = 0; j < 3; ++j) {  // voxel-unit loc vs metre centre: reference quirk
                        if (loc_xyz[j] > (current_octo->voxel_center_[j] + current_octo->quater_length_)) {
                            near_position[j] = near_position[j] + 1;
                        } else if (loc_xyz[j] < (current_octo->voxel_center_[j] - current_octo->quater_length_)) {
                            near_position[j] = near_position[j] - 1;
                        }
                    }
                    auto iter_near = map_manager_->voxel_map_.find(near_position);
                    if (iter_near != map_manager_->voxel_map_.end()) {
                        map_manager_->build_single_residual(cur_pt_var, iter_near->second, 0, is_success, prob,
                                                           single_ptpl);
                    }
                }
                if (is_success) {
                    ++success_this_iter;
                    ptpl_list.push_back(single_ptpl);
                    ptpl_src.push_back(i);
                }
            }
        }

        // 3) KF update with points (:186-213)
        size_t effect_num = ptpl_list.size();
        if (effect_num > 0) {
            ObsPoints obs;
            obs.h.resize(effect_num * 6);
            obs.R.resize(effect_num);
            obs.z.resize(effect_num);
            for (size_t k = 0; k < effect_num; ++k) {
                V3 crossmat_rotT_u = (ptpl_list[k].point_crossmat_ * eskf_.getRot().transpose()) * ptpl_list[k].normal_;
                for (int j = 0; j < 3; ++j) {
                    obs.h[k * 6 + j] = crossmat_rotT_u[j];
                    obs.h[k * 6 + 3 + j] = ptpl_list[k].normal_[j];
                }
                obs.z[k] = -ptpl_list[k].dis_to_plane_;
                Mat<1, 6> J_nq;
                for (int j = 0; j < 3; ++j) {
                    J_nq(0, j) = ptpl_list[k].point_w_[j] - ptpl_list[k].center_[j];
                    J_nq(0, 3 + j) = -ptpl_list[k].normal_[j];
                }
                M3 var = (((eskf_.getRot() * ext_rot_) * ptpl_list[k].body_cov_) * ext_rot_.transpose()) *
                         eskf_.getRot().transpose();
                double single_l = ((J_nq * ptpl_list[k].plane_var_) * J_nq.transpose())(0, 0);
                obs.R[k] = eskf_.config.lidar_point_meas_ratio *
                           (single_l + ((ptpl_list[k].normal_.transpose() * var) * ptpl_list[k].normal_)(0, 0));
            }
            if (dbg && it == iters_ - 1) {
                for (size_t k = 0; k < effect_num; ++k) {
                    size_t i = ptpl_src[k];
                    dbg->ok[i] = 1;
                    for (int j = 0; j < 6; ++j) dbg->h[i * 6 + j] = obs.h[k * 6 + j];
                    dbg->z[i] = obs.z[k];
                    dbg->R[i] = obs.R[k];
                }
            }
            PointGain g = eskf_.pointGain(obs, gain_mode_);
            eskf_.applyPointGain(g, true, it == iters_ - 1);
            last_state_update_time_ = current_time;
            eskf_update = true;
        }
        if (it == iters_ - 1) success_pts_size_out += success_this_iter;
    }

    // 4) voxel map update (:215-231)
    if (eskf_update) {
        for (size_t i = 0; i < points_size; ++i) {
            pv_list[i].point_w = eskf_.getRot() * pv_list[i].point_i + eskf_.getPos();
            cloud_down_world[idx_i + i].x = pv_list[i].point_w[0];
            cloud_down_world[idx_i + i].y = pv_list[i].point_w[1];
            cloud_down_world[idx_i + i].z = pv_list[i].point_w[2];
            cloud_down_world[idx_i + i].intensity = 255;
            M3 rot_extR = eskf_.getRot() * ext_rot_;
            M3 rot_crossmat = eskf_.getRot() * pv_list[i].point_crossmat;
            pv_list[i].var = (rot_extR * pv_list[i].body_var) * rot_extR.transpose() +
                             (rot_crossmat * eskf_.getRotCov()) * rot_crossmat.transpose() + eskf_.getPosCov();
        }
    }
    if (update_map_) map_manager_->UpdateVoxelMap(pv_list);
    return eskf_update;
}

bool KILO::predictUpdateImu(const ImuMeas& imu) {  // KILO.cc:235-258
    double current_time = imu.stamp;
    double dt_cov = current_time - last_state_update_time_;
    eskf_.predict(dt_cov, false, true);
    double dt = current_time - last_state_predict_time_;
    eskf_.predict(dt, true, false);
    last_state_predict_time_ = current_time;
    V3 imu_acc = vec3(imu.acc[0], imu.acc[1], imu.acc[2]);
    V3 imu_gyr = vec3(imu.gyr[0], imu.gyr[1], imu.gyr[2]);
    V3 za = (gravity_ / acc_norm_) * imu_acc - eskf_.state.imu_a - eskf_.state.ba;
    V3 zw = imu_gyr - eskf_.state.imu_w - eskf_.state.bw;
    double z[6] = {za[0], za[1], za[2], zw[0], zw[1], zw[2]};
    const EskfConfig& c = eskf_.config;
    double R[6] = {c.imu_acc_meas_noise, c.imu_acc_meas_noise, c.imu_acc_z_meas_noise,
                   c.imu_gyr_meas_noise, c.imu_gyr_meas_noise, c.imu_gyr_meas_noise};
    eskf_.updateByImu(z, R);
    last_state_update_time_ = current_time;
    return true;
}

bool KILO::predictUpdateKinImu(const KinImuMeas& kin_imu) {  // KILO.cc:260-314
    double current_time = kin_imu.time_stamp_;
    double dt_cov = current_time - last_state_update_time_;
    eskf_.predict(dt_cov, false, true);
    double dt = current_time - last_state_predict_time_;
    eskf_.predict(dt, true, false);
    last_state_predict_time_ = current_time;

    int contact_nums = 0;
    for (int i = 0; i < 4; ++i)
        if (kin_imu.contact_[i]) contact_nums++;
    const int m = 6 + 3 * contact_nums;
    std::vector<double> H((size_t)m * DIM, 0.0), z(m, 0.0), R(m, 0.0);
    for (int i = 0; i < 6; ++i) {
        H[(size_t)i * DIM + 9 + i] = 1.0;
        H[(size_t)i * DIM + 18 + i] = 1.0;
    }
    V3 imu_acc = vec3(kin_imu.acc_[0], kin_imu.acc_[1], kin_imu.acc_[2]);
    V3 imu_gyr = vec3(kin_imu.gyr_[0], kin_imu.gyr_[1], kin_imu.gyr_[2]);
    V3 za = (gravity_ / acc_norm_) * imu_acc - eskf_.state.imu_a - eskf_.state.ba;
    V3 zw = imu_gyr - eskf_.state.imu_w - eskf_.state.bw;
    for (int k = 0; k < 3; ++k) {
        z[k] = za[k];
        z[3 + k] = zw[k];
    }
    const EskfConfig& c = eskf_.config;
    R[0] = c.imu_acc_meas_noise; R[1] = c.imu_acc_meas_noise; R[2] = c.imu_acc_z_meas_noise;
    R[3] = c.imu_gyr_meas_noise; R[4] = c.imu_gyr_meas_noise; R[5] = c.imu_gyr_meas_noise;
    int idx = 0;
    M3 w_skew = skew(eskf_.state.imu_w);
    for (int i = 0; i < 4; ++i) {
        if (kin_imu.contact_[i]) {
            V3 foot_pos = vec3(kin_imu.foot_pos_[i][0], kin_imu.foot_pos_[i][1], kin_imu.foot_pos_[i][2]);
            V3 foot_vel = vec3(kin_imu.foot_vel_[i][0], kin_imu.foot_vel_[i][1], kin_imu.foot_vel_[i][2]);
            V3 w_skew_pos_vel = w_skew * foot_pos + foot_vel;
            M3 Hth = (-eskf_.getRot()) * skew(w_skew_pos_vel);
            M3 Hw = (-eskf_.getRot()) * skew(foot_pos);
            V3 zk = -eskf_.getVel() - eskf_.getRot() * w_skew_pos_vel;
            for (int r = 0; r < 3; ++r) {
                size_t row = (size_t)(6 + 3 * idx + r) * DIM;
                for (int cidx = 0; cidx < 3; ++cidx) {
                    H[row + 0 + cidx] = Hth(r, cidx);
                    H[row + 6 + cidx] = (r == cidx) ? 1.0 : 0.0;
                    H[row + 21 + cidx] = Hw(r, cidx);
                }
                z[6 + 3 * idx + r] = zk[r];
                R[6 + 3 * idx + r] = c.kin_meas_noise;
            }
            idx++;
        }
    }
    eskf_.updateByKinImu(m, H, z, R);
    last_state_update_time_ = current_time;
    return true;
}

void KILO::processSorted(double begin_time, const std::vector<PointXYZT>& pts, std::vector<PointXYZI>& world,
                         std::deque<ImuMeas>& imus, std::deque<KinImuMeas>& kin_imus,
                         size_t& success_pts_size_out) {  // KILO.cc:367-396
    const size_t pts_size = pts.size();
    world.resize(pts_size);
    size_t idx_i = 0;
    while (idx_i < pts_size) {
        double cur_point_time = begin_time + pts[idx_i].curvature;
        size_t idx_j = idx_i + 1;
        while (idx_j < pts_size && pts[idx_i].curvature == pts[idx_j].curvature) idx_j++;
        if (imu_mode_only_) {
            while (!imus.empty() && imus.front().stamp < cur_point_time) {
                predictUpdateImu(imus.front());
                imus.pop_front();
            }
        } else {
            while (!kin_imus.empty() && kin_imus.front().time_stamp_ < cur_point_time) {
                predictUpdateKinImu(kin_imus.front());
                kin_imus.pop_front();
            }
        }
        predictUpdatePoint(cur_point_time, idx_i, idx_j, pts, world, success_pts_size_out);
        idx_i = idx_j;
    }
}

}  // namespace lko
