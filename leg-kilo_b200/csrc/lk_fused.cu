// lk_fused.cu — the whole bucket loop of ONE scan (KILO.cc:367-396 -> predictUpdatePoint
// KILO.cc:108-233) as a single persistent cooperative kernel: every block keeps its own copy of
// the filter (state 36 + covariance 900 doubles) in shared memory and repeats the tiny serial parts
// (predict, 6x6 solve, state / covariance update) redundantly, so the only grid-wide
// synchronisation is ONE barrier per iteration — between writing the per-chunk partial sums of
// H^T R^-1 H / H^T R^-1 z and reading all of them. No relaunch, no host round trip between buckets
// or iterations. Used for batch = 1 (latency / streaming mode); large batches use lk_residual.cu.
#include <cooperative_groups.h>

#include "lk_kernels.h"
#include "lk_obs.cuh"
#include "lk_pass.cuh"
#include "lk_predict.cuh"
#include "lk_solve.cuh"

namespace lk {

namespace {

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define FT(slot) do { if (a.trace && threadIdx.x == 0 && (slot) < 32) a.trace[(size_t)blockIdx.x * 32 + (slot)] = gtime(); } while (0)

constexpr int BLOCK = FB;
constexpr int WARPS = BLOCK / 32;

struct PredictScratch {
    double F[900];
    double T[900];
    double Ps[900];
    ObsScratch obs;  // the IMU / Kin+IMU update follows its predict, so both live here
};

struct FusedSmem {
    BlockFilter f;
    ScanConst sc;
    double slice[WARPS * 32];
    double clk[2];
    union {  // predict and the point passes never overlap in time
        PredictScratch pr;
        PassSmem<BLOCK> pass;
    } u;
};

static_assert(sizeof(PredictScratch) <= sizeof(((PassSmem<BLOCK>*)0)->tile), "predict scratch must not reach the mbarriers");

__device__ __forceinline__ void grid_barrier(uint32_t* bar, uint32_t target) {
    // the block's partial rows were stored by other threads: bar.sync makes them observed by thread 0, whose
    // gpu-scope RELEASE reduction (no return value, so no round trip) publishes them cumulatively; the acquire
    // load that sees the last arrival, followed by bar.sync, orders every thread's reads of the other blocks' rows
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        uint32_t v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while (v < target);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(BLOCK, 1) k_scan_fused(const __grid_constant__ FusedArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem* sm = reinterpret_cast<FusedSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t scan = a.scan;
    uint32_t* bar = a.bar + a.parity;
    if (blockIdx.x == 0 && tid == 0) a.bar[1 - a.parity] = 0;  // the other launch parity's counter
    uint32_t bar_target = 0;
    FT(0);

    // the filter: reload from the staged inputs (idempotent runs)
    {
        const double* Pin = a.inline_in ? a.inl.P : a.P_in + (size_t)scan * 900;
        const double* xin = a.inline_in ? a.inl.x : a.x_in + (size_t)scan * 36;
        const double* cin = a.inline_in ? a.inl.clk : reinterpret_cast<const double*>(a.clk_in + scan);
        for (int e = tid; e < 900; e += BLOCK) sm->f.P[e] = Pin[e];
        if (tid < 36) sm->f.x[tid] = xin[tid];
        if (tid < 2) sm->clk[tid] = cin[tid];
    }
    pass_init<BLOCK>(&sm->u.pass);
    FT(1);
    uint32_t n_eff_total = 0;
    uint32_t phase = 0;
    DebugRows dbg;
    dbg.ok = nullptr; dbg.h = nullptr; dbg.z = nullptr; dbg.R = nullptr; dbg.key = nullptr;
    int it_global = 0;
    uint32_t mi = 0;  // next inertial / kinematic sample

    for (uint32_t k = 0; k < a.n_steps; ++k) {
        const StepInit in = a.inline_in ? a.inl.steps[k] : a.inits[(size_t)k * a.batch + scan];
        if (!in.active) continue;
        // with one chunk per block a lane sees the same point in every iteration of the bucket: issue its
        // load now, ahead of the predict (the point may sit in page-locked host memory)
        const bool one_chunk = a.lane_cache && (in.chunk_end - in.chunk_begin) <= gridDim.x;
        float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);
        if (one_chunk) {
            const uint32_t q = in.pt_begin + blockIdx.x * (uint32_t)BLOCK + (uint32_t)tid;
            if (blockIdx.x < in.chunk_end - in.chunk_begin && q < in.pt_end) pre = __ldg(a.pts + q);
        }
        // 0) every queued inertial / kinematic sample older than this bucket (KILO.cc:379-390)
        bool drained = false;
        while (mi < a.n_meas) {
            const double ts = a.imu ? a.imu[mi].stamp : a.kin[mi].stamp;
            if (!(ts < in.t_bucket)) break;
            block_predict_to(&sm->f, sm->clk, ts, sm->u.pr.F, sm->u.pr.T, sm->u.pr.Ps, a.Q);
            if (a.imu) block_obs_imu<BLOCK>(&sm->f, &sm->u.pr.obs, a.imu + mi, &a.ecfg, a.gravity, a.acc_norm);
            else block_obs_kinimu<BLOCK>(&sm->f, &sm->u.pr.obs, a.kin + mi, &a.ecfg, a.gravity, a.acc_norm);
            if (tid == 0) sm->clk[1] = ts;
            __syncthreads();
            ++mi;
            drained = true;
        }
        if (drained) {  // the scratch aliased the record tile
            fence_proxy_async();
            __syncthreads();
        }
        // 1) predict (KILO.cc:110-115): covariance with dt since the last UPDATE, state with dt since
        //    the last PREDICT; F is built from the pre-propagation state.
        const double dtc = in.t_bucket - sm->clk[1];
        const double dt = in.t_bucket - sm->clk[0];
        if (dtc != 0.0) {
            build_F(sm->u.pr.F, sm->f.x, dtc);
            cov_predict(sm->f.P, sm->u.pr.F, sm->u.pr.T, sm->u.pr.Ps, a.Q, dtc);
            // the scratch aliased the record tile (not the mbarriers, which sit behind it and keep
            // their phases): order these generic-proxy writes before the next bulk copies
            fence_proxy_async();
            __syncthreads();
        }
        if (dt != 0.0) {
            if (tid == 0) state_predict(sm->f.x, dt);
            __syncthreads();
        }
        if (tid == 0) sm->clk[0] = in.t_bucket;
        bool updated = false;
        uint32_t n_last = 0;
        LaneCache lc;
        lc.have = 0;
        lc.fail = 0;
        const uint32_t n_chunks = in.chunk_end - in.chunk_begin;
        for (int it = 0; it < a.iters; ++it, ++it_global) {
            scan_const_from(&sm->f, &sm->sc);
            __syncthreads();
            double* partial = a.partial + (size_t)(it_global & 1) * a.partial_stride;
            // 2) residual rows of my chunks -> one partial row per chunk
            for (uint32_t c = in.chunk_begin + blockIdx.x; c < in.chunk_end; c += gridDim.x) {
                ChunkDesc cd;  // chunks of a bucket are BLOCK points each (the host stages them the same way)
                cd.start = in.pt_begin + (c - in.chunk_begin) * (uint32_t)BLOCK;
                cd.count = min((uint32_t)BLOCK, in.pt_end - cd.start);
                double acc[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[i] = 0.0;
                unsigned long long* ptr = (a.trace && it_global == 1) ? a.trace + (size_t)(gridDim.x + blockIdx.x) * 64 : nullptr;
                if (one_chunk)
                    block_points_pass<BLOCK, false, true>(&sm->u.pass, phase, a.pts + cd.start, cd.count, (size_t)cd.start,
                                                          sm->sc, a.mv, a.g, acc, dbg, lc, ptr, pre);
                else
                    block_points_pass<BLOCK, false, false>(&sm->u.pass, phase, a.pts + cd.start, cd.count, (size_t)cd.start,
                                                           sm->sc, a.mv, a.g, acc, dbg, lc, ptr);
                double tot = warp_transpose_sum(acc, lane);
                __syncthreads();
                sm->slice[warp * 32 + lane] = tot;
                __syncthreads();
                if (tid < 32) {
                    double v = 0.0;
#pragma unroll
                    for (int w = 0; w < WARPS; ++w) v += sm->slice[w * 32 + tid];
                    partial[(size_t)c * PARTIAL_STRIDE + tid] = v;
                }
            }
            FT(2 + it_global * 4);
            // 3) the one grid-wide barrier of the iteration
            bar_target += gridDim.x;
            grid_barrier(bar, bar_target);
            FT(3 + it_global * 4);
            // 4) every block reduces all partial rows in the same fixed order and solves (eskf.cc:91-113)
            block_sum_partials<WARPS>(partial, in.chunk_begin, in.chunk_end, sm->slice, sm->f.acc);
            FT(4 + it_global * 4);
            const uint32_t n = block_solve_update<BLOCK>(&sm->f, it == a.iters - 1,
                                                         (a.trace && it_global >= 1 && it_global <= 2) ? a.trace + (size_t)(2 * gridDim.x + blockIdx.x) * 64 + (it_global - 1) * 8 : nullptr);
            FT(5 + it_global * 4);
            if (n > 0) {
                updated = true;
                if (tid == 0) sm->clk[1] = in.t_bucket;  // KILO.cc:212
            }
            n_last = n;
        }
        n_eff_total += n_last;
        (void)n_chunks;
        // 5) re-projection with the updated state (KILO.cc:216-224)
        __syncthreads();
        const float inten = updated ? 255.0f : 0.0f;
        const Globals& g = a.g;
        for (uint32_t c = in.chunk_begin + blockIdx.x; c < in.chunk_end; c += gridDim.x) {
            ChunkDesc cd;
            cd.start = in.pt_begin + (c - in.chunk_begin) * (uint32_t)BLOCK;
            cd.count = min((uint32_t)BLOCK, in.pt_end - cd.start);
            if ((uint32_t)tid < cd.count) {
                double pix, piy, piz;
                if (one_chunk && lc.have) {  // the lane's own point, already in the IMU frame
                    pix = lc.pix; piy = lc.piy; piz = lc.piz;
                } else {
                    float4 pt = __ldg(a.pts + cd.start + tid);
                    double bx = pt.x, by = pt.y, bz = pt.z;
                    pix = g.Re[0] * bx + g.Re[1] * by + g.Re[2] * bz + g.te[0];
                    piy = g.Re[3] * bx + g.Re[4] * by + g.Re[5] * bz + g.te[1];
                    piz = g.Re[6] * bx + g.Re[7] * by + g.Re[8] * bz + g.te[2];
                }
                const double* X = sm->f.x;
                float4 o;
                o.x = (float)(X[0] * pix + X[1] * piy + X[2] * piz + X[9]);
                o.y = (float)(X[3] * pix + X[4] * piy + X[5] * piz + X[10]);
                o.z = (float)(X[6] * pix + X[7] * piy + X[8] * piz + X[11]);
                o.w = inten;
                a.world[cd.start + tid] = o;
            }
        }
    }
    FT(30);
    if (blockIdx.x == 0) {
        __syncthreads();
        for (int e = tid; e < 900; e += BLOCK) a.P[(size_t)scan * 900 + e] = sm->f.P[e];
        if (tid < 36) a.x[(size_t)scan * 36 + tid] = sm->f.x[tid];
        if (tid < 2) reinterpret_cast<double*>(a.clk + scan)[tid] = sm->clk[tid];
        if (tid == 0) a.n_eff[scan] = n_eff_total;
    }
    FT(31);
}

}  // namespace

size_t fused_smem_bytes() { return sizeof(FusedSmem); }

int fused_max_blocks(int device) {
    static int cached[64];
    static bool have[64];
    if (device >= 0 && device < 64 && have[device]) return cached[device];
    cudaFuncSetAttribute(k_scan_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem));
    int per_sm = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scan_fused, BLOCK, sizeof(FusedSmem));
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int n = per_sm * sms;
    if (device >= 0 && device < 64) { cached[device] = n; have[device] = true; }
    return n;
}

// `cooperative` = go through cudaLaunchCooperativeKernel (the driver checks co-residency and serialises
// cooperative grids: ~10 us between back-to-back launches). The plain launch relies on the same fact the
// host already enforces — grid <= fused_max_blocks(), i.e. every block fits on the device at once — so the
// blocks spinning on the grid barrier can only ever wait for blocks that are resident or that become
// resident as soon as unrelated work drains; nothing they wait for depends on them.
cudaError_t launch_scan_fused(const FusedArgs& a, uint32_t grid, cudaStream_t s, bool cooperative) {
    if (cooperative) {
        void* params[] = {(void*)&a};
        return cudaLaunchCooperativeKernel((const void*)k_scan_fused, dim3(grid), dim3(BLOCK), params, sizeof(FusedSmem), s);
    }
    k_scan_fused<<<dim3(grid), dim3(BLOCK), sizeof(FusedSmem), s>>>(a);
    return cudaGetLastError();
}

}  // namespace lk
