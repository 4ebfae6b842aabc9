// lk_obs.cuh — the filter steps between buckets, on a shared-memory-resident filter
// (BlockFilter): KILO::predictUpdateImu (KILO.cc:235-258) -> ESKF::updateByImu (eskf.cc:125-135) and
// KILO::predictUpdateKinImu (KILO.cc:260-314) -> ESKF::updateByKinImu (eskf.cc:137-145), each preceded
// by the two predicts of KILO.cc:237-241 / :262-266. Block-level (FB threads); the m x m innovation
// system (m = 6 or 6 + 3 x contacts <= 18) is solved by a warp-parallel Gauss-Jordan with partial
// pivoting on the augmented matrix [S | z | H P] held in shared memory.
#pragma once
#include "lk_predict.cuh"
#include "lk_solve.cuh"

namespace lk {

constexpr int OBS_MAX_M = 18;
constexpr int OBS_AUG_COLS = OBS_MAX_M + 1 + 30;  // [S | z | HP]

struct ObsScratch {            // aliases the predict scratch / point-pass area of the caller
    double H[OBS_MAX_M * 30];
    double PHT[30 * OBS_MAX_M];
    double Aug[OBS_MAX_M * OBS_AUG_COLS];
    double z[OBS_MAX_M];
    double r[OBS_MAX_M];
    double delta[30];
    int ok;
};

// Both predicts of predictUpdate{Imu,KinImu,Point}: covariance with dt since the last UPDATE (F from
// the pre-propagation state), then the state with dt since the last PREDICT. clk = {last_predict,
// last_update}. F, T, Ps: 900-double scratch arrays. All threads.
__device__ inline void block_predict_to(BlockFilter* f, double* clk, double t, double* F, double* T, double* Ps,
                                        const double* Q) {
    const double dtc = t - clk[1];
    const double dt = t - clk[0];
    __syncthreads();
    if (dtc != 0.0) {  // dt == 0 is an exact no-op (F = I, dt^2 Q = 0)
        build_F(F, f->x, dtc);
        cov_predict(f->P, F, T, Ps, Q, dtc);
    }
    if (dt != 0.0) {
        if (threadIdx.x == 0) state_predict(f->x, dt);
    }
    if (threadIdx.x == 0) clk[0] = t;
    __syncthreads();
}

// Gauss-Jordan with partial pivoting on Aug (m rows, ncols columns, row stride OBS_AUG_COLS): the
// first m columns are reduced to I, the rest become S^-1 [z | HP]. One warp; returns false on a
// zero pivot.
__device__ inline bool warp_gauss_jordan(double* Aug, int m, int ncols, int lane) {
    bool ok = true;
    for (int k = 0; k < m; ++k) {
        int piv = k;
        double best = fabs(Aug[k * OBS_AUG_COLS + k]);
        for (int i = k + 1; i < m; ++i) {
            double a = fabs(Aug[i * OBS_AUG_COLS + k]);
            if (a > best) { best = a; piv = i; }
        }
        if (best == 0.0) ok = false;
        __syncwarp();
        if (piv != k)
            for (int c = lane; c < ncols; c += 32) {
                double t = Aug[k * OBS_AUG_COLS + c];
                Aug[k * OBS_AUG_COLS + c] = Aug[piv * OBS_AUG_COLS + c];
                Aug[piv * OBS_AUG_COLS + c] = t;
            }
        __syncwarp();
        const double inv = 1.0 / Aug[k * OBS_AUG_COLS + k];
        double mult[OBS_MAX_M];
        for (int i = 0; i < m; ++i) mult[i] = Aug[i * OBS_AUG_COLS + k];
        __syncwarp();
        for (int c = lane; c < ncols; c += 32) {
            const double pk = Aug[k * OBS_AUG_COLS + c] * inv;
            Aug[k * OBS_AUG_COLS + c] = pk;
            for (int i = 0; i < m; ++i)
                if (i != k) Aug[i * OBS_AUG_COLS + c] -= mult[i] * pk;
        }
        __syncwarp();
    }
    return ok;
}

// K = PHT S^-1, delta = K z, x (+)= delta, P -= K (H P) with a dense m x 30 H in s->H, innovation
// s->z and noise s->r (eskf.cc:137-145; also used for the IMU-only update, whose closed form
// eskf.cc:127-134 is this with H = [0 I6 0 I6 0]). All threads.
// imu_rows: the first 6 rows of H are the inertial rows [0 I6 0 I6 0] (ones at columns 9 + a and 18 + a): their products
// are two-term sums. Dropping the structural zeros keeps every sum bit-identical to the dense loops (same order, exact zeros).
template <int NTHREADS>
__device__ inline void block_update_dense(BlockFilter* f, ObsScratch* s, int m, bool imu_rows = false) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // PHT = P H^T (30 x m) and HP = H P (m x 30) into the augmented matrix
    for (int e = tid; e < 30 * m; e += NTHREADS) {
        const int i = e / m, a = e % m;
        double v = 0.0;
        if (imu_rows && a < 6) {
            v += f->P[i * 30 + 9 + a];
            v += f->P[i * 30 + 18 + a];
        } else {
            for (int k = 0; k < 30; ++k) v += f->P[i * 30 + k] * s->H[a * 30 + k];
        }
        s->PHT[i * OBS_MAX_M + a] = v;
    }
    for (int e = tid; e < m * 30; e += NTHREADS) {
        const int a = e / 30, j = e % 30;
        double v = 0.0;
        if (imu_rows && a < 6) {
            v += f->P[(9 + a) * 30 + j];
            v += f->P[(18 + a) * 30 + j];
        } else {
            for (int k = 0; k < 30; ++k) v += s->H[a * 30 + k] * f->P[k * 30 + j];
        }
        s->Aug[a * OBS_AUG_COLS + m + 1 + j] = v;
    }
    __syncthreads();
    // S = H PHT + diag(r), z
    for (int e = tid; e < m * m; e += NTHREADS) {
        const int a = e / m, b = e % m;
        double v = (a == b) ? s->r[a] : 0.0;
        if (imu_rows && a < 6) {
            v += s->PHT[(9 + a) * OBS_MAX_M + b];
            v += s->PHT[(18 + a) * OBS_MAX_M + b];
        } else {
            for (int k = 0; k < 30; ++k) v += s->H[a * 30 + k] * s->PHT[k * OBS_MAX_M + b];
        }
        s->Aug[a * OBS_AUG_COLS + b] = v;
    }
    if (tid < m) s->Aug[tid * OBS_AUG_COLS + m] = s->z[tid];
    __syncthreads();
    if (warp == 0) {
        const bool ok = warp_gauss_jordan(s->Aug, m, m + 1 + 30, lane);
        if (lane == 0) s->ok = ok ? 1 : 0;
    }
    __syncthreads();
    if (s->ok) {
        if (tid < 30) {
            double v = 0.0;
            for (int a = 0; a < m; ++a) v += s->PHT[tid * OBS_MAX_M + a] * s->Aug[a * OBS_AUG_COLS + m];
            s->delta[tid] = v;
        }
        __syncthreads();
        // P -= PHT (S^-1 HP)
        double upd[4];
        int cnt = 0;
        for (int e = tid; e < 900; e += NTHREADS, ++cnt) {
            const int i = e / 30, j = e % 30;
            double v = 0.0;
            for (int a = 0; a < m; ++a) v += s->PHT[i * OBS_MAX_M + a] * s->Aug[a * OBS_AUG_COLS + m + 1 + j];
            upd[cnt] = v;
        }
        __syncthreads();
        cnt = 0;
        for (int e = tid; e < 900; e += NTHREADS, ++cnt) f->P[e] -= upd[cnt];
        if (tid == 0) state_boxplus(f->x, s->delta);
    }
    __syncthreads();
}

// KILO::predictUpdateImu's observation (KILO.cc:243-254). Thread 0 fills z / r, all build H.
template <int NTHREADS>
__device__ inline void block_obs_imu(BlockFilter* f, ObsScratch* s, const lk_imu_meas* meas, const lk_eskf_cfg* cfg,
                                     double gravity, double acc_norm) {
    const int tid = threadIdx.x;
    for (int e = tid; e < 6 * 30; e += NTHREADS) {
        const int a = e / 30, k = e % 30;
        s->H[e] = (k == 9 + a || k == 18 + a) ? 1.0 : 0.0;
    }
    if (tid < 6) {
        const double* x = f->x;
        const double sc = gravity / acc_norm;
        // z = [(g/|a|) acc - imu_a - ba ; gyr - imu_w - bw]
        s->z[tid] = (tid < 3) ? (sc * meas->acc[tid] - x[24 + tid] - x[15 + tid]) : (meas->gyr[tid - 3] - x[27 + tid - 3] - x[18 + tid - 3]);
        s->r[tid] = (tid < 2) ? cfg->imu_acc_meas_noise : (tid == 2 ? cfg->imu_acc_z_meas_noise : cfg->imu_gyr_meas_noise);
    }
    __syncthreads();
    block_update_dense<NTHREADS>(f, s, 6, true);
}

// KILO::predictUpdateKinImu's observation (KILO.cc:268-310).
template <int NTHREADS>
__device__ inline void block_obs_kinimu(BlockFilter* f, ObsScratch* s, const lk_kinimu_meas* meas, const lk_eskf_cfg* cfg,
                                        double gravity, double acc_norm) {
    const int tid = threadIdx.x;
    int legs[4], nc = 0;
    for (int i = 0; i < 4; ++i)
        if (meas->contact[i]) legs[nc++] = i;
    const int m = 6 + 3 * nc;
    for (int e = tid; e < m * 30; e += NTHREADS) {
        const int a = e / 30, k = e % 30;
        s->H[e] = (a < 6 && (k == 9 + a || k == 18 + a)) ? 1.0 : 0.0;
    }
    __syncthreads();
    const double* x = f->x;
    if (tid < 6) {
        const double sc = gravity / acc_norm;
        s->z[tid] = (tid < 3) ? (sc * meas->acc[tid] - x[24 + tid] - x[15 + tid]) : (meas->gyr[tid - 3] - x[27 + tid - 3] - x[18 + tid - 3]);
        s->r[tid] = (tid < 2) ? cfg->imu_acc_meas_noise : (tid == 2 ? cfg->imu_acc_z_meas_noise : cfg->imu_gyr_meas_noise);
    }
    if (tid >= 32 && tid < 32 + nc) {  // one thread per contact foot
        const int c = tid - 32, leg = legs[c];
        const double* R = x;
        const double w[3] = {x[27], x[28], x[29]};
        const double fp[3] = {meas->foot_pos[leg][0], meas->foot_pos[leg][1], meas->foot_pos[leg][2]};
        const double fv[3] = {meas->foot_vel[leg][0], meas->foot_vel[leg][1], meas->foot_vel[leg][2]};
        // w x p + v
        const double u[3] = {w[1] * fp[2] - w[2] * fp[1] + fv[0], w[2] * fp[0] - w[0] * fp[2] + fv[1], w[0] * fp[1] - w[1] * fp[0] + fv[2]};
        const double Ku[9] = {0, -u[2], u[1], u[2], 0, -u[0], -u[1], u[0], 0};
        const double Kp[9] = {0, -fp[2], fp[1], fp[2], 0, -fp[0], -fp[1], fp[0], 0};
        double mR[9], Hth[9], Hw[9];
        for (int i = 0; i < 9; ++i) mR[i] = -R[i];
        mat3_mul(mR, Ku, Hth);  // -R [w x p + v]x
        mat3_mul(mR, Kp, Hw);   // -R [p]x
        for (int r = 0; r < 3; ++r) {
            double* row = s->H + (6 + 3 * c + r) * 30;
            for (int q = 0; q < 3; ++q) {
                row[q] = Hth[r * 3 + q];
                row[6 + q] = (r == q) ? 1.0 : 0.0;
                row[21 + q] = Hw[r * 3 + q];
            }
            // z = -vel - R (w x p + v)
            s->z[6 + 3 * c + r] = -x[12 + r] - (R[r * 3] * u[0] + R[r * 3 + 1] * u[1] + R[r * 3 + 2] * u[2]);
            s->r[6 + 3 * c + r] = cfg->kin_meas_noise;
        }
    }
    __syncthreads();
    block_update_dense<NTHREADS>(f, s, m, true);
}

}  // namespace lk
