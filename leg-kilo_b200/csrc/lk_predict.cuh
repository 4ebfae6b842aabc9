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

// The structural non-zeros of getFx's F (eskf.cc:72-81), row by row, columns ascending: rows 0-2 (theta) hold Exp(-w dt) and
// dt on imu_w; rows 3-5 (pos) 1 and dt on vel; rows 6-8 (vel) -dt R [a]x, 1, dt on grav, dt R on imu_a; every other row is a
// row of the identity. Skipping the structural zeros leaves every sum bit-identical to the dense product (the terms that
// remain are added in the same ascending order; the dropped ones are exact zeros).
__device__ __forceinline__ int f_row_pattern(int r, int (&col)[8]) {
    if (r < 3) { col[0] = 0; col[1] = 1; col[2] = 2; col[3] = 21 + r; return 4; }
    if (r < 6) { col[0] = r; col[1] = r + 3; return 2; }
    if (r < 9) { col[0] = 0; col[1] = 1; col[2] = 2; col[3] = r; col[4] = r + 9; col[5] = 18; col[6] = 19; col[7] = 20; return 8; }
    col[0] = r;
    return 1;
}

// P <- F P F^T + dt^2 Q  (eskf.cc:86-87). F, T, Ps in shared memory; P wherever the caller keeps it.
__device__ inline void cov_predict(double* Pg, const double* F, double* T, double* Ps, const double* Q, double dt) {
    const int tid = threadIdx.x;
    for (int e = tid; e < 900; e += FB) Ps[e] = Pg[e];
    __syncthreads();
    // T = F P: only rows 0..8 of F differ from the identity
    for (int e = tid; e < 900; e += FB) {
        const int i = e / 30, j = e % 30;
        double s;
        if (i >= 9) {
            s = Ps[e];
        } else {
            int col[8];
            const int n = f_row_pattern(i, col);
            s = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < n) s += F[i * 30 + col[q]] * Ps[col[q] * 30 + j];
        }
        T[e] = s;
    }
    __syncthreads();
    const double dt2 = dt * dt;
    // P = T F^T + dt^2 Q: column j of the product is a sparse combination for j < 9, a copy otherwise
    for (int e = tid; e < 900; e += FB) {
        const int i = e / 30, j = e % 30;
        double s;
        if (j >= 9) {
            s = T[e];
        } else {
            int col[8];
            const int n = f_row_pattern(j, col);
            s = 0.0;
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < n) s += T[i * 30 + col[q]] * F[j * 30 + col[q]];
        }
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
