// lk_residual.cu — the hot kernel: per point transform -> voxel key -> hash probe -> plane gates ->
// residual / Jacobian row -> block-reduced H^T R^-1 H, H^T R^-1 z; the last block of every scan
// does the 6x6 information-form Kalman solve and the state / covariance update.
//
// Follows, row by row: a3 KILO.cc:127-140 + voxel_map.cc:22-40, a4 KILO.cc:143-149,
// a5 voxel_map.cc:363-427, a6 KILO.cc:156-178, a7 KILO.cc:187-210, a8 eskf.cc:91-113,
// a9 eskf.cc:18-29. Algebra is restructured (never the results' meaning):
//   * calcBodyCov's A*A^T is range^2 (I - u u^T) because {b1, b2, u} is orthonormal, so
//     n^T M Sigma_b M^T n = rv (u.w)^2 + range^2 dv (|w|^2 - (u.w)^2), w = M^T n;
//   * n^T (R[pi]x) P_tt (R[pi]x)^T n = h_t^T P_tt h_t with h_t = pi x (R^T n), the Jacobian row itself;
//   * K = P H^T (H P H^T + R)^-1 is evaluated as P[:,0:6] (I + A P66)^-1 with A = sum h^T h / R.
#include "lk_kernels.h"
#include "lk_pass.cuh"
#include "lk_solve.cuh"

namespace lk {

namespace {

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define LK_TRACE(slot) do { if (a.trace && threadIdx.x == 0) a.trace[(size_t)blockIdx.x * 8 + (slot)] = gtime(); } while (0)
#define LK_TRACE_TAIL(slot) do { if (a.trace && threadIdx.x == 0) a.trace[(size_t)gridDim.x * 8 + (slot)] = gtime(); } while (0)

constexpr int BLOCK = 256;  // 8 warps; one point per thread per pass
constexpr int WARPS = BLOCK / 32;

struct TailSmem {
    BlockFilter f;
    double slice[WARPS * 32];
};

// eskf.cc:91-113 + State::operator+= for one scan; run by the last block to finish that scan.
__device__ void scan_solve(const ResidualArgs& a, uint32_t scan, TailSmem* ts) {
    const int tid = threadIdx.x;
    const ScanStep st = a.step[scan];
    double* xg = a.x + (size_t)scan * 36;
    double* Pg = a.P + (size_t)scan * 900;
    // filter into shared memory (these loads overlap the partial-row loads below)
    for (int e = tid; e < 900; e += BLOCK) ts->f.P[e] = Pg[e];
    if (tid < 36) ts->f.x[tid] = xg[tid];
    block_sum_partials<WARPS>(a.partial, st.chunk_begin, st.chunk_end, ts->slice, ts->f.acc);
    LK_TRACE_TAIL(1);
    const uint32_t n = block_solve_update<BLOCK>(&ts->f, a.last_iter != 0);
    LK_TRACE_TAIL(2);
    if (n > 0) {
        if (tid < 36) xg[tid] = ts->f.x[tid];
        if (a.last_iter)
            for (int e = tid; e < 900; e += BLOCK) Pg[e] = ts->f.P[e];
        if (tid == 0) a.clk[scan].last_update_time = st.t_bucket;  // KILO.cc:212
    }
    if (n > 0 || a.last_iter) scan_const_from(&ts->f, a.sc + scan);
    if (tid == 0) {
        ScanStep* sp = a.step + scan;
        sp->n_eff_last = n;
        if (n > 0) sp->updated = 1;
        if (a.last_iter) a.n_eff[scan] += n;
        a.ticket[scan] = 0;
    }
}

union ResidualSmem {  // the tail runs after the passes: same storage
    PassSmem<BLOCK> pass;
    TailSmem tail;
};

// SINGLE: every chunk has at most BLOCK points (one pass, latency mode) so no accumulator is live
// while a point is being evaluated; otherwise the block makes several passes over its chunk.
template <bool DEBUG, bool SINGLE>
__global__ void __launch_bounds__(BLOCK, (DEBUG || SINGLE) ? 1 : 2) k_residual(const __grid_constant__ ResidualArgs a) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ ScanConst s_sc;
    __shared__ uint32_t s_last;
    ResidualSmem* rs = reinterpret_cast<ResidualSmem*>(s_raw);
    TailSmem* ts = &rs->tail;
    const int tid = threadIdx.x;
    LK_TRACE(0);
    const ChunkDesc cd = a.chunks[a.chunk_first + blockIdx.x];
    if (tid < (int)(sizeof(ScanConst) / sizeof(double)))
        reinterpret_cast<double*>(&s_sc)[tid] = reinterpret_cast<const double*>(a.sc + cd.scan)[tid];
    pass_init<BLOCK>(&rs->pass);
    LK_TRACE(1);
    MapView mv;
    mv.slots = a.slots; mv.hash_mask = a.hash_mask; mv.nodes = a.nodes;
    DebugRows dbg;
    dbg.ok = a.dbg_ok; dbg.h = a.dbg_h; dbg.z = a.dbg_z; dbg.R = a.dbg_R; dbg.key = a.dbg_key;

    double acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.0;
    uint32_t phase = 0;
    for (uint32_t off = 0; off < cd.count; off += BLOCK) {
        const uint32_t n = min((uint32_t)BLOCK, cd.count - off);
        block_points_pass<BLOCK, DEBUG>(&rs->pass, phase, a.pts + cd.start + off, n, (size_t)cd.start + off, s_sc, mv, a.g,
                                        acc, dbg);
        if (SINGLE) break;
    }
    if (DEBUG) return;
    LK_TRACE(2);

    // block reduction: transposing shuffle tree in the warp, one row per warp in smem, fixed-order sum
    const int lane = tid & 31, warp = tid >> 5;
    double* s_red = ts->slice;
    double tot = warp_transpose_sum(acc, lane);
    s_red[warp * 32 + lane] = tot;
    __syncthreads();
    if (tid < 32) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) v += s_red[w * 32 + tid];
        a.partial[(size_t)(a.chunk_first + blockIdx.x) * PARTIAL_STRIDE + tid] = v;
        __threadfence();
    }
    __syncthreads();
    if (tid == 0) {
        const ScanStep* sp = a.step + cd.scan;
        uint32_t n_chunks = sp->chunk_end - sp->chunk_begin;
        uint32_t t = atomicAdd(a.ticket + cd.scan, 1u);
        s_last = (t == n_chunks - 1) ? 1u : 0u;
    }
    __syncthreads();
    LK_TRACE(3);
    if (s_last) {
        __threadfence();
        LK_TRACE_TAIL(0);
        scan_solve(a, cd.scan, ts);
        LK_TRACE_TAIL(7);
    }
}

// The per-scan solve on its own (after lk_stream2.cu's residual pass): one block per scan.
__global__ void __launch_bounds__(BLOCK) k_scan_tail(const __grid_constant__ ResidualArgs a, const uint32_t scan_first) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    TailSmem* ts = reinterpret_cast<TailSmem*>(s_raw);
    const uint32_t scan = scan_first + blockIdx.x;
    const ScanStep st = a.step[scan];
    if (!st.active || st.chunk_end == st.chunk_begin) return;
    scan_solve(a, scan, ts);
}

}  // namespace

void launch_scan_tail(const ResidualArgs& a, uint32_t scan_first, uint32_t n_scans, cudaStream_t s) {
    if (n_scans == 0) return;
    static PerDeviceOnce once;
    if (once.first()) cudaFuncSetAttribute(k_scan_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TailSmem));
    k_scan_tail<<<n_scans, BLOCK, sizeof(TailSmem), s>>>(a, scan_first);
}

void launch_residual(const ResidualArgs& a, uint32_t n_chunks, bool debug, bool single, cudaStream_t s) {
    if (n_chunks == 0) return;
    static PerDeviceOnce once;
    const size_t smem = sizeof(ResidualSmem);
    if (once.first()) {
        cudaFuncSetAttribute(k_residual<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_residual<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_residual<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_residual<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    }
    if (debug)
        k_residual<true, false><<<n_chunks, BLOCK, smem, s>>>(a);
    else if (single)
        k_residual<false, true><<<n_chunks, BLOCK, smem, s>>>(a);
    else
        k_residual<false, false><<<n_chunks, BLOCK, smem, s>>>(a);
}

// ---- re-projection with the updated state (KILO.cc:216-224) ---------------------------------
namespace {
__global__ void __launch_bounds__(BLOCK) k_reproject(const __grid_constant__ ReprojectArgs a) {
    __shared__ ScanConst s_sc;
    __shared__ uint32_t s_upd;
    const int tid = threadIdx.x;
    const ChunkDesc cd = a.chunks[a.chunk_first + blockIdx.x];
    if (tid < (int)(sizeof(ScanConst) / sizeof(double)))
        reinterpret_cast<double*>(&s_sc)[tid] = reinterpret_cast<const double*>(a.sc + cd.scan)[tid];
    if (tid == 0) s_upd = a.step[cd.scan].updated;
    __syncthreads();
    const Globals& g = a.g;
    const float inten = s_upd ? 255.0f : 0.0f;
    for (uint32_t i = tid; i < cd.count; i += BLOCK) {
        float4 pt = __ldg(a.pts + cd.start + i);
        double bx = pt.x, by = pt.y, bz = pt.z;
        double pix = g.Re[0] * bx + g.Re[1] * by + g.Re[2] * bz + g.te[0];
        double piy = g.Re[3] * bx + g.Re[4] * by + g.Re[5] * bz + g.te[1];
        double piz = g.Re[6] * bx + g.Re[7] * by + g.Re[8] * bz + g.te[2];
        float4 o;
        o.x = (float)(s_sc.R[0] * pix + s_sc.R[1] * piy + s_sc.R[2] * piz + s_sc.p[0]);
        o.y = (float)(s_sc.R[3] * pix + s_sc.R[4] * piy + s_sc.R[5] * piz + s_sc.p[1]);
        o.z = (float)(s_sc.R[6] * pix + s_sc.R[7] * piy + s_sc.R[8] * piz + s_sc.p[2]);
        o.w = inten;
        a.world[cd.start + i] = o;
    }
}
}  // namespace

void launch_reproject(const ReprojectArgs& a, uint32_t n_chunks, cudaStream_t s) {
    if (n_chunks == 0) return;
    k_reproject<<<n_chunks, BLOCK, 0, s>>>(a);
}

}  // namespace lk
