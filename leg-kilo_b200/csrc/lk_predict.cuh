// lk_predict.cuh — ESKF::predict pieces (eskf.cc:64-89) as block-level device functions.
#pragma once
#include "lk_device.cuh"

namespace lk {

constexpr int FB = 256;  // threads per block of every kernel using these helpers

// getFx (eskf.cc:72-81) into a dense 30x30 in shared memory. x = 36-double lk_state.
__device__ inline void build_F(double* F, const double* x, double dt) {
    const int tid = threadIdx.x;
    for (int e = tid; e < 900; e += FB) F[e] = (e / 30 == e % 30) ? 1.0 : 0.0;
    __syncthreads();
    if (tid == 0) {
        const double* R = x;
        const double* a = x + 24;  // imu_a
        const double* w = x + 27;  // imu_w
        double E[9];
        so3_exp_vec(-dt * w[0], -dt * w[1], -dt * w[2], E);
        double Ka[9] = {0, -a[2], a[1], a[2], 0, -a[0], -a[1], a[0], 0};
        double mR[9], RK[9];
        for (int i = 0; i < 9; ++i) mR[i] = (-dt) * R[i];
        mat3_mul(mR, Ka, RK);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                F[i * 30 + j] = E[i * 3 + j];
                F[i * 30 + 21 + j] = (i == j) ? dt : 0.0;
                F[(3 + i) * 30 + 6 + j] = (i == j) ? dt : 0.0;
                F[(6 + i) * 30 + j] = RK[i * 3 + j];
                F[(6 + i) * 30 + 15 + j] = (i == j) ? dt : 0.0;
                F[(6 + i) * 30 + 18 + j] = dt * R[i * 3 + j];
            }
    }
    __syncthreads();
}

// P <- F P F^T + dt^2 Q  (eskf.cc:86-87). F, T in shared; P in global (L2-resident, 7.2 KB).
__device__ inline void cov_predict(double* Pg, const double* F, double* T, double* Ps, const double* Q, double dt) {
    const int tid = threadIdx.x;
    for (int e = tid; e < 900; e += FB) Ps[e] = Pg[e];
    __syncthreads();
    for (int e = tid; e < 900; e += FB) {
        int i = e / 30, j = e % 30;
        double s = 0.0;
        for (int k = 0; k < 30; ++k) s += F[i * 30 + k] * Ps[k * 30 + j];
        T[e] = s;
    }
    __syncthreads();
    const double dt2 = dt * dt;
    for (int e = tid; e < 900; e += FB) {
        int i = e / 30, j = e % 30;
        double s = 0.0;
        for (int k = 0; k < 30; ++k) s += T[i * 30 + k] * F[j * 30 + k];
        Pg[e] = s + dt2 * Q[e];
    }
    __syncthreads();
}

// getFunctionf + State::operator+= (eskf.cc:64-70, :18-29), single thread.
__device__ inline void state_predict(double* x, double dt) {
    double d[30];
    for (int i = 0; i < 30; ++i) d[i] = 0.0;
    const double* R = x;
    const double* vel = x + 12;
    const double* grav = x + 21;
    const double* a = x + 24;
    const double* w = x + 27;
    for (int k = 0; k < 3; ++k) {
        d[k] = dt * w[k];
        d[3 + k] = dt * vel[k];
        d[6 + k] = dt * (R[k * 3] * a[0] + R[k * 3 + 1] * a[1] + R[k * 3 + 2] * a[2] + grav[k]);
    }
    state_boxplus(x, d);
}

}  // namespace lk
