// lk_filter.cu — ESKF steps that are not per-point: covariance / state prediction
// (eskf.cc:64-89, called as in KILO.cc:110-115) and the per-bucket preparation of the constants
// the residual blocks read.
#include "lk_kernels.h"
#include "lk_obs.cuh"
#include "lk_predict.cuh"

namespace lk {

namespace {

// Step 1 of KILO::predictUpdatePoint (KILO.cc:110-115) for every scan of the range, then the
// per-scan constants of this bucket. When `reset` is set the filter is first re-loaded from the
// staged inputs (x_in / P_in / clk_in), which makes a run idempotent without extra copies.
__global__ void __launch_bounds__(FB) k_predict_prepare(const PredictArgs a) {
    __shared__ double sF[900], sT[900], sP[900];
    __shared__ double sx[36];
    __shared__ double sclk[2];
    const int scan = a.scan_first + blockIdx.x, tid = threadIdx.x;
    const StepInit in = a.init[scan];
    double* xg = a.x + (size_t)scan * 36;
    double* Pg = a.P + (size_t)scan * 900;
    if (a.reset) {
        const double* xi = a.x_in + (size_t)scan * 36;
        const double* Pi = a.P_in + (size_t)scan * 900;
        for (int e = tid; e < 900; e += FB) Pg[e] = Pi[e];
        if (tid < 36) { double v = xi[tid]; sx[tid] = v; xg[tid] = v; }
        if (tid < 2) sclk[tid] = reinterpret_cast<const double*>(a.clk_in + scan)[tid];
        if (tid == 0) a.n_eff[scan] = 0;
    } else {
        if (tid < 36) sx[tid] = xg[tid];
        if (tid < 2) sclk[tid] = reinterpret_cast<const double*>(a.clk + scan)[tid];
    }
    if (tid == 0) {
        ScanStep st;
        st.chunk_begin = in.chunk_begin; st.chunk_end = in.chunk_end;
        st.pt_begin = in.pt_begin; st.pt_end = in.pt_end;
        st.t_bucket = in.t_bucket; st.active = in.active; st.updated = 0; st.n_eff_last = 0; st.pad = 0;
        a.step[scan] = st;
        a.ticket[scan] = 0;
    }
    __syncthreads();
    if (!in.active) {
        if (a.reset && tid < 2) reinterpret_cast<double*>(a.clk + scan)[tid] = sclk[tid];
        return;
    }
    const double dtc = in.t_bucket - sclk[1];  // since the last UPDATE  (KILO.cc:111)
    const double dt = in.t_bucket - sclk[0];   // since the last PREDICT (KILO.cc:113)
    if (dtc != 0.0) {  // dt == 0 is an exact no-op (F = I, dt^2 Q = 0)
        build_F(sF, sx, dtc);
        cov_predict(Pg, sF, sT, sP, a.Q, dtc);
    }
    if (dt != 0.0) {
        if (tid == 0) state_predict(sx, dt);
        __syncthreads();
        if (tid < 36) xg[tid] = sx[tid];
    }
    if (tid == 0) {
        a.clk[scan].last_predict_time = in.t_bucket;
        a.clk[scan].last_update_time = sclk[1];
    }
    // constants for the residual blocks
    ScanConst* sc = a.sc + scan;
    if (tid < 9) sc->R[tid] = sx[tid];
    else if (tid < 12) sc->p[tid - 9] = sx[tid];
    else if (tid < 24) {
        const int ut[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
        const int q = (tid - 12) % 6, o = (tid < 18) ? 0 : 3;
        const int i = ut[q][0] + o, j = ut[q][1] + o;
        const double v = 0.5 * (Pg[i * 30 + j] + Pg[j * 30 + i]);
        if (tid < 18) sc->Pth[q] = v; else sc->Ppp[q] = v;
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

struct ObsSmem {
    BlockFilter f;
    double clk[2];
    double F[900], T[900], Ps[900];
    ObsScratch obs;
    double slice[(FB / 32) * 32];
};

// A queue of inertial (imu) or kinematic+inertial (kin) samples applied to ONE host-visible filter:
// predictUpdateImu / predictUpdateKinImu per sample, in order (KILO.cc:235-314).
__global__ void __launch_bounds__(FB) k_filter_obs(double* x, double* P, const double* Q, lk_stream_clock* clk,
                                                   const lk_imu_meas* imu, const lk_kinimu_meas* kin, uint32_t n,
                                                   lk_eskf_cfg cfg, double gravity, double acc_norm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ObsSmem* sm = reinterpret_cast<ObsSmem*>(smem_raw);
    const int tid = threadIdx.x;
    for (int e = tid; e < 900; e += FB) sm->f.P[e] = P[e];
    if (tid < 36) sm->f.x[tid] = x[tid];
    if (tid < 2) sm->clk[tid] = reinterpret_cast<const double*>(clk)[tid];
    __syncthreads();
    for (uint32_t i = 0; i < n; ++i) {
        const double t = imu ? imu[i].stamp : kin[i].stamp;
        block_predict_to(&sm->f, sm->clk, t, sm->F, sm->T, sm->Ps, Q);
        if (imu) block_obs_imu<FB>(&sm->f, &sm->obs, imu + i, &cfg, gravity, acc_norm);
        else block_obs_kinimu<FB>(&sm->f, &sm->obs, kin + i, &cfg, gravity, acc_norm);
        if (tid == 0) sm->clk[1] = t;  // last_state_update_time_ = current_time
        __syncthreads();
    }
    for (int e = tid; e < 900; e += FB) P[e] = sm->f.P[e];
    if (tid < 36) x[tid] = sm->f.x[tid];
    if (tid < 2) reinterpret_cast<double*>(clk)[tid] = sm->clk[tid];
}

// Streaming (one scan, KILO.cc:375-395): the samples that precede a bucket AND the bucket's own predict + constants in ONE launch —
// k_filter_obs followed by k_predict_prepare(reset = 0) with the filter kept in shared memory in between. Same device functions
// in the same order on the same values => bit-identical to the two launches.
__global__ void __launch_bounds__(FB) k_obs_predict_prepare(const PredictArgs a, const lk_imu_meas* imu, const lk_kinimu_meas* kin,
                                                            uint32_t n, lk_eskf_cfg cfg, double gravity, double acc_norm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ObsSmem* sm = reinterpret_cast<ObsSmem*>(smem_raw);
    const int scan = a.scan_first, tid = threadIdx.x;
    const StepInit in = a.init[scan];
    double* xg = a.x + (size_t)scan * 36;
    double* Pg = a.P + (size_t)scan * 900;
    for (int e = tid; e < 900; e += FB) sm->f.P[e] = Pg[e];
    if (tid < 36) sm->f.x[tid] = xg[tid];
    if (tid < 2) sm->clk[tid] = reinterpret_cast<const double*>(a.clk + scan)[tid];
    if (tid == 0) {
        ScanStep st;
        st.chunk_begin = in.chunk_begin; st.chunk_end = in.chunk_end;
        st.pt_begin = in.pt_begin; st.pt_end = in.pt_end;
        st.t_bucket = in.t_bucket; st.active = in.active; st.updated = 0; st.n_eff_last = 0; st.pad = 0;
        a.step[scan] = st;
        a.ticket[scan] = 0;
    }
    __syncthreads();
    if (!in.active) return;
    for (uint32_t i = 0; i < n; ++i) {  // predictUpdateImu / predictUpdateKinImu per sample (KILO.cc:235-314)
        const double t = imu ? imu[i].stamp : kin[i].stamp;
        block_predict_to(&sm->f, sm->clk, t, sm->F, sm->T, sm->Ps, a.Q);
        if (imu) block_obs_imu<FB>(&sm->f, &sm->obs, imu + i, &cfg, gravity, acc_norm);
        else block_obs_kinimu<FB>(&sm->f, &sm->obs, kin + i, &cfg, gravity, acc_norm);
        if (tid == 0) sm->clk[1] = t;
        __syncthreads();
    }
    block_predict_to(&sm->f, sm->clk, in.t_bucket, sm->F, sm->T, sm->Ps, a.Q);  // KILO.cc:110-115
    for (int e = tid; e < 900; e += FB) Pg[e] = sm->f.P[e];
    if (tid < 36) xg[tid] = sm->f.x[tid];
    if (tid < 2) reinterpret_cast<double*>(a.clk + scan)[tid] = sm->clk[tid];
    ScanConst* sc = a.sc + scan;
    if (tid < 9) sc->R[tid] = sm->f.x[tid];
    else if (tid < 12) sc->p[tid - 9] = sm->f.x[tid];
    else if (tid < 24) {
        const int ut[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
        const int q = (tid - 12) % 6, o = (tid < 18) ? 0 : 3;
        const int i = ut[q][0] + o, j = ut[q][1] + o;
        const double v = 0.5 * (sm->f.P[i * 30 + j] + sm->f.P[j * 30 + i]);
        if (tid < 18) sc->Pth[q] = v; else sc->Ppp[q] = v;
    }
}

// ESKF::updateByPoints (eskf.cc:91-113) from explicit rows: information-form sums, then the block solve.
__global__ void __launch_bounds__(FB) k_update_by_points(double* x, double* P, uint32_t n, const double* h,
                                                         const double* z, const double* r) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ObsSmem* sm = reinterpret_cast<ObsSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int e = tid; e < 900; e += FB) sm->f.P[e] = P[e];
    if (tid < 36) sm->f.x[tid] = x[tid];
    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.0;
    for (uint32_t k = tid; k < n; k += FB) {
        Row row;
        for (int j = 0; j < 6; ++j) row.h[j] = h[(size_t)k * 6 + j];
        row.z = z[k];
        row.R = r[k];
        const double w = 1.0 / row.R;
        int q = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            const double hw = row.h[a] * w;
#pragma unroll
            for (int c = a; c < 6; ++c) acc[q++] += hw * row.h[c];
            acc[ACC_B + a] += hw * row.z;
        }
        acc[ACC_SUMR] += row.R;
        acc[ACC_CNT] += 1.0;
    }
    const double tot = warp_transpose_sum(acc, lane);
    sm->slice[warp * 32 + lane] = tot;
    __syncthreads();
    if (tid < 32) {
        double v = 0.0;
        for (int w = 0; w < FB / 32; ++w) v += sm->slice[w * 32 + tid];
        sm->f.acc[tid] = v;
    }
    __syncthreads();
    block_solve_update<FB>(&sm->f, true);
    __syncthreads();
    for (int e = tid; e < 900; e += FB) P[e] = sm->f.P[e];
    if (tid < 36) x[tid] = sm->f.x[tid];
}

}  // namespace

static void obs_attrs() {  // function attributes are per device
    static PerDeviceOnce once;
    if (once.first()) {
        cudaFuncSetAttribute(k_filter_obs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ObsSmem));
        cudaFuncSetAttribute(k_obs_predict_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ObsSmem));
        cudaFuncSetAttribute(k_update_by_points, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ObsSmem));
    }
}

void launch_filter_obs(double* x, double* P, const double* Q, lk_stream_clock* clk, const lk_imu_meas* imu,
                       const lk_kinimu_meas* kin, uint32_t n, const lk_eskf_cfg& cfg, double gravity, double acc_norm,
                       cudaStream_t s) {
    obs_attrs();
    k_filter_obs<<<1, FB, sizeof(ObsSmem), s>>>(x, P, Q, clk, imu, kin, n, cfg, gravity, acc_norm);
}

void launch_update_by_points(double* x, double* P, uint32_t n, const double* h, const double* z, const double* r,
                             cudaStream_t s) {
    obs_attrs();
    k_update_by_points<<<1, FB, sizeof(ObsSmem), s>>>(x, P, n, h, z, r);
}

void launch_obs_predict_prepare(const PredictArgs& a, const lk_imu_meas* imu, const lk_kinimu_meas* kin, uint32_t n,
                                const lk_eskf_cfg& cfg, double gravity, double acc_norm, cudaStream_t s) {
    obs_attrs();
    k_obs_predict_prepare<<<1, FB, sizeof(ObsSmem), s>>>(a, imu, kin, n, cfg, gravity, acc_norm);
}

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
