// lk_fused.cu — the whole bucket loop of ONE scan (KILO.cc:367-396 -> predictUpdatePoint
// KILO.cc:108-233) as a single persistent kernel, one 256-point chunk per block: every block keeps its
// own copy of the filter (state 36 + covariance 900 doubles) in shared memory and repeats the tiny
// serial parts (predict, 6x6 solve, state / covariance update) redundantly, so the only grid-wide
// exchange per iteration is ONE all-reduce of the 29 sums H^T R^-1 H | H^T R^-1 z | sum R | count.
//
// That all-reduce has no barrier: rows travel in the flagged format of lk_llsync.cuh (data and "ready"
// tag in the same 8-byte words), two levels deep — the block of a group's first chunk adds the group's
// LK_GROUP rows and publishes the group row, every block adds the (<= 19) group rows — so a round costs
// two L2 round trips after the slowest block, moves ~10 KB per block instead of every row to every block,
// and the sum has ONE fixed order that the multi-kernel path reproduces (block_sum_partials).
//
// Per-lane cache (lk_pass.cuh: cached_points_pass): a lane keeps its point, voxel keys, lookups and BOTH
// candidate plane records (home + the reference's one fallback neighbour, TMA-staged together) across the
// iterations of a bucket, so iterations 2..n touch no global memory unless a key moved.
//
// Two instantiations keep the instruction footprint of the common case small: OBS = false (no inertial /
// kinematic queue: the scan-at-once and streaming-without-queue shapes) and OBS = true (queue drained
// before every bucket, KILO.cc:379-390). The predict and the queue drain are out of line in both.
// The kernel is PDL-aware (griddepcontrol): launched with programmatic stream serialisation, the next
// scan's blocks become resident and run their prologue (filter load, point prefetch) while this scan's
// last blocks drain; everything that could collide with the previous launch (flagged rows, outputs) sits
// behind griddepcontrol.wait.
#include "lk_insert.cuh"
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
// per-bucket stamps of the streaming variants (block 0, first 64 buckets, 8 stamps each, behind the per-block area)
#define FTS(slot) do { if (a.trace && blockIdx.x == 0 && threadIdx.x == 0 && k < 64) a.trace[(size_t)gridDim.x * 32 + (size_t)k * 8 + (slot)] = gtime(); } while (0)

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
        CachedPassSmem<BLOCK> pass;
    } u;
};

// with the map insert inside: the plane-fit staging tiles of the warps, with their own mbarriers (initialised once)
struct FusedSmemIns {
    FusedSmem base;
    WarpTile wt[WARPS];
    MapDev md;   // copies for the out-of-line insert phase
    Globals g;
};

// Phase 2 of the in-kernel UpdateVoxelMap, out of line: the octree / plane-fit code then gets a register allocation of
// its own (as in the stand-alone insert kernel) instead of spilling inside the persistent kernel's.
__device__ __noinline__ void fused_insert_phase2(FusedSmemIns* si, const uint32_t* touched, const int* iroot, const DevPoint* ipts,
                                                 int* pend, uint32_t n_touched, uint32_t n_bucket) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    MapDev md = si->md;
    WarpTile* wt = si->wt + warp;
    for (uint32_t t = blockIdx.x * (uint32_t)WARPS + (uint32_t)warp; t < n_touched; t += gridDim.x * (uint32_t)WARPS)
        warp_insert_root_scan(md, si->g, wt, __ldcg(&touched[t]), iroot, ipts, n_bucket, pend, lane);
}

static_assert(sizeof(PredictScratch) <= sizeof(((CachedPassSmem<BLOCK>*)0)->tile), "predict scratch must not reach the mbarriers");
static_assert(sizeof(FusedSmemIns) <= 227 * 1024, "one block per SM");

// KILO.cc:110-115: covariance with dt since the last UPDATE, state with dt since the last PREDICT; F is built
// from the pre-propagation state. Out of line: a scan-at-once call (bucket time == both clocks) never gets here.
__device__ __noinline__ void fused_predict(FusedSmem* sm, const double* Q, double dtc, double dt) {
    if (dtc != 0.0) {
        build_F(sm->u.pr.F, sm->f.x, dtc);
        cov_predict(sm->f.P, sm->u.pr.F, sm->u.pr.T, sm->u.pr.Ps, Q, dtc);
    }
    if (dt != 0.0) {
        if (threadIdx.x == 0) state_predict(sm->f.x, dt);
    }
    // the scratch aliased the record tiles (not the mbarriers, which sit behind them and keep their
    // phases): order these generic-proxy writes before the next bulk copies
    fence_proxy_async();
    __syncthreads();
}

// every queued inertial / kinematic sample older than this bucket (KILO.cc:379-390)
__device__ __noinline__ void fused_drain_queue(FusedSmem* sm, const FusedArgs& a, uint32_t& mi, double t_bucket) {
    bool drained = false;
    while (mi < a.n_meas) {
        const double ts = a.imu ? a.imu[mi].stamp : a.kin[mi].stamp;
        if (!(ts < t_bucket)) break;
        block_predict_to(&sm->f, sm->clk, ts, sm->u.pr.F, sm->u.pr.T, sm->u.pr.Ps, a.Q);
        if (a.imu) block_obs_imu<BLOCK>(&sm->f, &sm->u.pr.obs, a.imu + mi, &a.ecfg, a.gravity, a.acc_norm);
        else block_obs_kinimu<BLOCK>(&sm->f, &sm->u.pr.obs, a.kin + mi, &a.ecfg, a.gravity, a.acc_norm);
        if (threadIdx.x == 0) sm->clk[1] = ts;
        __syncthreads();
        ++mi;
        drained = true;
    }
    if (drained) {  // the scratch aliased the record tiles
        fence_proxy_async();
        __syncthreads();
    }
}

template <bool INL> struct InlineSel { typedef FusedInline type; };
template <> struct InlineSel<false> { typedef FusedNoInline type; };

// A grid-wide barrier out of the flagged-row all-reduce (every block contributes a zero row). Release side: every thread
// fences its earlier global writes. Acquire side: a gpu-scope fence AFTER the barrier — on this architecture it also
// invalidates the SM's L1 (CCTL.IVALL, see the SASS), so the plain (L1-cached) loads of the next phase cannot be served from
// a line cached before another SM rewrote it; the read-only path (ld.global.nc) is not used on the map in these kernels.
__device__ __forceinline__ void grid_sync(const FusedArgs& a, uint32_t& sync_idx) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x < 32) (void)ll_allreduce(a.ll, sync_idx & 1u, a.epoch + sync_idx, blockIdx.x, gridDim.x, 0.0, (int)threadIdx.x);
    ++sync_idx;
    __syncthreads();
    __threadfence();
}

// OBS: an inertial / kinematic queue is drained before every bucket. INL: the small inputs ride in the parameter block.
// INS: UpdateVoxelMap runs inside the kernel after every bucket (KILO.cc:231): the map is then read through L2.
template <bool OBS, bool INL, bool INS>
__global__ void __launch_bounds__(BLOCK, 1) k_scan_fused(const __grid_constant__ FusedArgs a,
                                                         const __grid_constant__ typename InlineSel<INL>::type inl) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    FusedSmem* sm = reinterpret_cast<FusedSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t scan = a.scan;
    // let the next launch of the stream (if it was launched with programmatic serialisation) start placing its
    // blocks as soon as this grid's blocks are all running; it blocks in griddepcontrol.wait until we are done
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    FT(0);
    // the filter: always reloaded from the staged inputs (idempotent runs); these are inputs, never written by a kernel
    {
        const double* Pin;
        const double* xin;
        const double* cin;
        if constexpr (INL) { Pin = inl.P; xin = inl.x; cin = inl.clk; }
        else { Pin = a.P_in + (size_t)scan * 900; xin = a.x_in + (size_t)scan * 36; cin = reinterpret_cast<const double*>(a.clk_in + scan); }
        if (a.slim_p && blockIdx.x != 0) {
            // a single-bucket scan without a queue never predicts; blocks other than 0 (which stores the covariance) then
            // only ever read P[:, 0:6]: the 6x6 corner for the solve and the scan constants, the 30x6 strip for delta
            if (tid < 180) sm->f.P[(tid / 6) * 30 + tid % 6] = Pin[(tid / 6) * 30 + tid % 6];
        } else {
            for (int e = tid; e < 900; e += BLOCK) sm->f.P[e] = Pin[e];
        }
        if (tid < 36) sm->f.x[tid] = xin[tid];
        if (tid < 2) sm->clk[tid] = cin[tid];
    }
    if constexpr (INS) {
        FusedSmemIns* si = reinterpret_cast<FusedSmemIns*>(smem_raw);
        WarpTile* wt = si->wt + warp;
        if (lane == 0) {
            mbar_init(&wt->bar, 1);
            wt->phase = 0;
        }
        if (tid == 0) si->md = a.md;
        if (tid < (int)(sizeof(Globals) / 4)) reinterpret_cast<uint32_t*>(&si->g)[tid] = reinterpret_cast<const uint32_t*>(&a.g)[tid];
    }
    cached_pass_init<BLOCK>(&sm->u.pass);  // mbarrier init fence + block barrier
    FT(1);
    uint32_t n_eff_total = 0;
    uint32_t phase = 0;
    uint32_t it_global = 0;  // index of the next grid-wide exchange (all-reduce or barrier): tag and buffer parity
    uint32_t cslot = 0;      // which of the two "touched roots" counters the current bucket uses (INS)
    uint32_t mi = 0;  // next inertial / kinematic sample
    bool dep_waited = false;

    for (uint32_t k = 0; k < a.n_steps; ++k) {
        StepInit in;
        if constexpr (INL) in = inl.steps[k];
        else in = a.inits[(size_t)k * a.batch + scan];
        if (!in.active) continue;
        const uint32_t n_chunks = in.chunk_end - in.chunk_begin;  // <= gridDim.x (the host checks)
        // a lane sees the same point in every iteration of the bucket: issue its load now, ahead of the predict
        // (the point may sit in page-locked host memory)
        const uint32_t my_start = in.pt_begin + blockIdx.x * (uint32_t)BLOCK;
        const uint32_t my_count = blockIdx.x < n_chunks ? min((uint32_t)BLOCK, in.pt_end - my_start) : 0u;
        float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((uint32_t)tid < my_count) pre = __ldg(a.pts + my_start + tid);
        FTS(0);
        if constexpr (OBS) fused_drain_queue(sm, a, mi, in.t_bucket);
        FTS(1);
        const double dtc = in.t_bucket - sm->clk[1];
        const double dt = in.t_bucket - sm->clk[0];
        if (dtc != 0.0 || dt != 0.0) fused_predict(sm, a.Q, dtc, dt);
        if (tid == 0) sm->clk[0] = in.t_bucket;
        FTS(2);
        bool updated = false, cov_pending = false;
        uint32_t n_last = 0;
        LaneCache lc;
        lc.have = 0;
        const bool more_steps = k + 1 < a.n_steps;
        for (int it = 0; it < a.iters; ++it, ++it_global) {
            scan_const_from(&sm->f, &sm->sc);
            __syncthreads();
            // 1) residual rows of my chunk
            double acc[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = 0.0;
            if (!a.lane_cache && lc.have == 2) lc.have = 1;
            cached_points_pass<BLOCK, INS>(&sm->u.pass, phase, my_count, sm->sc, a.mv, a.g, acc, lc, pre);
            const double tot = warp_transpose_sum(acc, lane);
            sm->slice[warp * 32 + lane] = tot;
            __syncthreads();
            FT(2 + it_global * 4);
            if (it == 0) FTS(3);
            // 2) all-reduce of the block rows (warp 0), no barrier
            if (warp == 0) {
                double v = 0.0;
#pragma unroll
                for (int w = 0; w < WARPS; ++w) v += sm->slice[w * 32 + lane];
                if (!dep_waited) asm volatile("griddepcontrol.wait;" ::: "memory");  // the rows / outputs of the previous launch
                sm->f.acc[lane] = ll_allreduce(a.ll, it_global & 1u, a.epoch + it_global, blockIdx.x, n_chunks, v, lane);
            }
            dep_waited = true;
            __syncthreads();
            FT(3 + it_global * 4);
            if (it == 0) FTS(4);
            // 3) every block solves redundantly (eskf.cc:91-113); the covariance update of the last iteration is
            //    deferred behind the re-projection, and skipped where nobody reads the result
            const bool last = it == a.iters - 1;
            const uint32_t n = block_solve_state(&sm->f);
            FT(4 + it_global * 4);
            if (n > 0) {
                updated = true;
                if (tid == 0) sm->clk[1] = in.t_bucket;  // KILO.cc:212
                if (last) cov_pending = INS || more_steps || blockIdx.x == 0;  // the insert needs the updated covariance in every block
            }
            n_last = n;
        }
        n_eff_total += n_last;
        FTS(5);
        if constexpr (!INS) {
            // 4) re-projection with the updated state (KILO.cc:216-224)
            if ((uint32_t)tid < my_count) {
                const double* X = sm->f.x;
                float4 o;
                o.x = (float)(X[0] * lc.pix + X[1] * lc.piy + X[2] * lc.piz + X[9]);
                o.y = (float)(X[3] * lc.pix + X[4] * lc.piy + X[5] * lc.piz + X[10]);
                o.z = (float)(X[6] * lc.pix + X[7] * lc.piy + X[8] * lc.piz + X[11]);
                o.w = updated ? 255.0f : 0.0f;
                a.world[my_start + tid] = o;
            }
            if (cov_pending) block_cov_update<BLOCK>(&sm->f);
        } else {
            // 4') re-projection AND map insert with the updated state and covariance (KILO.cc:216-231). Phase 1, the blocks
            //     that hold the bucket's points: pointWithVar, world cloud, find-or-create of the root voxel.
            if (cov_pending) block_cov_update<BLOCK>(&sm->f);
            __syncthreads();
            scan_const_from(&sm->f, &sm->sc);
            __syncthreads();
            MapDev md = a.md;
            const uint32_t n_bucket = in.pt_end - in.pt_begin;
            if ((uint32_t)tid < my_count) {
                // (without an update the reference inserts the point as the residual loop left it, KILO.cc:127-140, :215: the
                // same formulas at the unchanged state)
                DevPoint p;
                make_insert_point(lc.pix, lc.piy, lc.piz, lc.pbx, lc.pby, lc.pbz, sm->sc, a.g, p);
                const uint32_t li = my_start + tid - in.pt_begin;
                a.ipts[li] = p;
                float4 o;
                o.x = (float)p.pw[0]; o.y = (float)p.pw[1]; o.z = (float)p.pw[2];
                o.w = updated ? 255.0f : 0.0f;
                a.world[my_start + tid] = o;
                a.iroot[li] = insert_register_point(md, a.g, p, a.pend, a.touched, &a.ins_counters[cslot]);
            }
            grid_sync(a, it_global);
            FTS(6);
            // Phase 2, every warp of every block: one touched root at a time, its points in index order
            {
                const uint32_t n_touched = __ldcg(&a.ins_counters[cslot]);
                if (blockIdx.x == 0 && tid == 0) a.ins_counters[cslot ^ 1u] = 0;  // the next bucket's counter
                fused_insert_phase2(reinterpret_cast<FusedSmemIns*>(smem_raw), a.touched, a.iroot, a.ipts, a.pend, n_touched, n_bucket);
                cslot ^= 1u;
            }
            grid_sync(a, it_global);
            FTS(7);
        }
    }
    FT(30);
    if (blockIdx.x == 0) {
        if (!dep_waited) asm volatile("griddepcontrol.wait;" ::: "memory");
        __syncthreads();
        for (int e = tid; e < 900; e += BLOCK) a.P[(size_t)scan * 900 + e] = sm->f.P[e];
        if (tid < 36) a.x[(size_t)scan * 36 + tid] = sm->f.x[tid];
        if (tid < 2) reinterpret_cast<double*>(a.clk + scan)[tid] = sm->clk[tid];
        if (tid == 0) {
            a.n_eff[scan] = n_eff_total;
            // did any wait of this launch give up (lk_llsync.cuh / lk_async.cuh watchdogs)? The host looks at this word first
            // and only then pays for the detailed read-back
            *a.status = (*reinterpret_cast<volatile uint32_t*>(a.ll.stall) ? 1u : 0u) | (*reinterpret_cast<volatile uint32_t*>(&lk_stall_note[0]) ? 2u : 0u);
        }
    }
    FT(31);
}

template <bool OBS, bool INL, bool INS>
cudaError_t launch_one(const FusedArgs& a, const FusedInline* inl, uint32_t grid, cudaStream_t s, int mode) {
    auto kern = k_scan_fused<OBS, INL, INS>;
    constexpr size_t SMEM = INS ? sizeof(FusedSmemIns) : sizeof(FusedSmem);
    static bool attr[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e != cudaSuccess) return e;
        attr[dev] = true;
    }
    typename InlineSel<INL>::type local_inl;
    const typename InlineSel<INL>::type* ip;
    if constexpr (INL) ip = inl;
    else { local_inl.unused = 0; ip = &local_inl; }
    if (mode == FUSED_LAUNCH_COOPERATIVE) {
        void* params[] = {(void*)&a, (void*)ip};
        return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(BLOCK), params, SMEM, s);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(BLOCK);
    cfg.dynamicSmemBytes = SMEM;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = mode == FUSED_LAUNCH_PDL ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, a, *ip);
}

}  // namespace

size_t fused_smem_bytes() { return sizeof(FusedSmem); }

// read and clear this translation unit's watchdog note (lk_async.cuh)
int fused_read_stall(uint32_t out[8]) {
    if (cudaMemcpyFromSymbol(out, lk_stall_note, 32) != cudaSuccess) return -1;
    if (out[0]) {
        const uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        cudaMemcpyToSymbol(lk_stall_note, z, 32);
    }
    return 0;
}

int fused_max_blocks(int device) {
    static int cached[64];
    static bool have[64];
    if (device >= 0 && device < 64 && have[device]) return cached[device];
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int n = sms < LL_MAX_CHUNKS ? sms : LL_MAX_CHUNKS;  // one block per SM (launch bounds + shared memory)
    if (device >= 0 && device < 64) { cached[device] = n; have[device] = true; }
    return n;
}

// mode: plain <<<>>>, cooperative (the driver checks co-residency and serialises cooperative grids: ~10 us between
// back-to-back launches) or programmatic stream serialisation (PDL). The non-cooperative launches rely on the
// fact the host enforces — grid <= SM count at one block per SM, i.e. every block fits on the device at once —
// so blocks polling for rows only ever wait for blocks that are resident or become resident as soon as unrelated
// work drains; lk_api.cu serialises fused grids of different handles of one process (see INTEGRATION.md for the
// multi-process caveat and the cooperative knob).
cudaError_t launch_scan_fused(const FusedArgs& a, const FusedInline* inl, uint32_t grid, cudaStream_t s, int mode) {
    const bool obs = a.n_meas > 0;
    if (a.insert) {  // streaming with map insertion: never with inline inputs
        if (inl) return cudaErrorInvalidValue;
        return obs ? launch_one<true, false, true>(a, inl, grid, s, mode) : launch_one<false, false, true>(a, inl, grid, s, mode);
    }
    if (obs) return inl ? launch_one<true, true, false>(a, inl, grid, s, mode) : launch_one<true, false, false>(a, inl, grid, s, mode);
    return inl ? launch_one<false, true, false>(a, inl, grid, s, mode) : launch_one<false, false, false>(a, inl, grid, s, mode);
}

}  // namespace lk
