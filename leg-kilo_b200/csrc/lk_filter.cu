// lk_filter.cu — ESKF steps that are not per-point: covariance / state prediction
// (eskf.cc:64-89, called as in KILO.cc:110-115) and the per-bucket preparation of the constants
// the residual blocks read.
#include "lk_kernels.h"

namespace lk {

namespace {

constexpr int FB = 256;

// getFx (eskf.cc:72-81) into a dense 30x30 in shared memory. x = 36-double lk_state.
__device__ void build_F(double* F, const double* x, double dt) {
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
__device__ void cov_predict(double* Pg, const double* F, double* T, double* Ps, const double* Q, double dt) {
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
__device__ void state_predict(double* x, double dt) {
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

__device__ void write_scan_const(ScanConst* sc, const double* x, const double* Pg) {
    for (int i = 0; i < 9; ++i) sc->R[i] = x[i];
    for (int i = 0; i < 3; ++i) sc->p[i] = x[9 + i];
    const int ut[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
    for (int q = 0; q < 6; ++q) {
        int i = ut[q][0], j = ut[q][1];
        sc->Pth[q] = 0.5 * (Pg[i * 30 + j] + Pg[j * 30 + i]);
        sc->Ppp[q] = 0.5 * (Pg[(3 + i) * 30 + 3 + j] + Pg[(3 + j) * 30 + 3 + i]);
    }
}

// Step 1 of KILO::predictUpdatePoint (KILO.cc:110-115) for every scan of the batch, then the
// per-scan constants of this bucket.
__global__ void __launch_bounds__(FB) k_predict_prepare(const PredictArgs a) {
    __shared__ double sF[900], sT[900], sP[900];
    __shared__ double sx[36];
    const int scan = blockIdx.x, tid = threadIdx.x;
    const StepInit in = a.init[scan];
    if (tid == 0) {
        ScanStep st;
        st.chunk_begin = in.chunk_begin; st.chunk_end = in.chunk_end;
        st.pt_begin = in.pt_begin; st.pt_end = in.pt_end;
        st.t_bucket = in.t_bucket; st.active = in.active; st.updated = 0; st.n_eff_last = 0; st.pad = 0;
        a.step[scan] = st;
        a.ticket[scan] = 0;
    }
    if (!in.active) return;
    double* xg = a.x + (size_t)scan * 36;
    double* Pg = a.P + (size_t)scan * 900;
    if (tid < 36) sx[tid] = xg[tid];
    __syncthreads();
    const double dtc = in.t_bucket - a.clk[scan].last_update_time;
    const double dt = in.t_bucket - a.clk[scan].last_predict_time;
    if (dtc != 0.0) {  // dt == 0 is an exact no-op (F = I, dt^2 Q = 0)
        build_F(sF, sx, dtc);
        cov_predict(Pg, sF, sT, sP, a.Q, dtc);
    }
    if (tid == 0) {
        if (dt != 0.0) {
            state_predict(sx, dt);
            for (int i = 0; i < 36; ++i) xg[i] = sx[i];
        }
        a.clk[scan].last_predict_time = in.t_bucket;
        write_scan_const(a.sc + scan, sx, Pg);
    }
}

__global__ void __launch_bounds__(FB) k_predict_dt(double* x, double* P, const double* Q, const double* dtv,
                                                   int prop_state, int prop_cov) {
    __shared__ double sF[900], sT[900], sP[900];
    __shared__ double sx[36];
    const int f = blockIdx.x, tid = threadIdx.x;
    double* xg = x + (size_t)f * 36;
    double* Pg = P + (size_t)f * 900;
    const double dt = dtv[f];
    if (tid < 36) sx[tid] = xg[tid];
    __syncthreads();
    // ESKF::predict applies the state first, then builds Fx from the UPDATED state (eskf.cc:84-87)
    if (prop_state) {
        if (tid == 0) {
            state_predict(sx, dt);
            for (int i = 0; i < 36; ++i) xg[i] = sx[i];
        }
        __syncthreads();
    }
    if (prop_cov) {
        build_F(sF, sx, dt);
        cov_predict(Pg, sF, sT, sP, Q, dt);
    }
}

}  // namespace

void launch_predict_prepare(const PredictArgs& a, cudaStream_t s) {
    if (a.batch <= 0) return;
    k_predict_prepare<<<a.batch, FB, 0, s>>>(a);
}

void launch_predict_dt(double* x, double* P, const double* Q, const double* dt, int batch, int prop_state,
                       int prop_cov, cudaStream_t s) {
    if (batch <= 0) return;
    k_predict_dt<<<batch, FB, 0, s>>>(x, P, Q, dt, prop_state, prop_cov);
}

}  // namespace lk
