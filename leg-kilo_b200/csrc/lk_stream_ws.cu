// lk_stream_ws.cu — throughput family of the residual pass (>= 2 scans per call), warp-specialised and
// persistent: one 512-thread block per SM walks its share of the launch's chunks;
//   * 4 PRODUCER warps (one warpgroup, 56 registers) run ahead of the arithmetic: point load -> voxel key
//     (KILO.cc:143-148) -> pair probe of the root table -> cooperative 16-byte async copies of the 32
//     plane records of a group into a ring of shared-memory stages (mbarrier full/empty handshake);
//   * 12 CONSUMER warps (152 registers) never wait for global memory on the main path: they take the
//     stages in order, evaluate gates + Jacobian row + noise (voxel_map.cc:363-411, KILO.cc:187-210) from
//     shared memory and accumulate H^T R^-1 H / H^T R^-1 z in registers.
// Points that fail at their home plane (or whose root is an octree interior node) are listed per warp and
// finished by all consumers together with the full reference sequence (home descent, then the one
// neighbour voxel, KILO.cc:156-178). Everything is statically assigned (group q -> producer q % 4, stage
// q % S, consumer q % 12; list entries in warp-major order), so sums are bitwise reproducible.
// The per-scan solve runs as its own small kernel afterwards (lk_residual.cu: k_scan_tail).
#include "lk_kernels.h"
#include <algorithm>

#include "lk_pass.cuh"

namespace lk {

namespace {

constexpr int WS_THREADS = 512;
constexpr int WS_NPROD = 8;   // two warpgroups (setmaxnreg granularity is a warpgroup)
constexpr int WS_NCONS = 8;
constexpr int WS_PROD_REGS = 64, WS_CONS_REGS = 192;
static_assert((WS_NPROD * WS_PROD_REGS + WS_NCONS * WS_CONS_REGS) * 32 <= 65536, "register file");
static_assert(WS_NPROD % 4 == 0 && WS_NCONS % 4 == 0 && (WS_NPROD + WS_NCONS) * 32 == WS_THREADS, "warpgroups");
constexpr int WS_CONS_THREADS = WS_NCONS * 32;
// Every stage must always be filled by the same producer and drained by the same consumer (parity waits
// cannot tell phase n from phase n-2): a multiple of both warp counts.
constexpr int WS_STAGES = 24;
static_assert(WS_STAGES % WS_NPROD == 0 && WS_STAGES % WS_NCONS == 0, "stage ownership must be static");
// Records travel by per-lane TMA bulk copies (true) or by warp-cooperative 16-byte cp.async (false).
constexpr bool WS_TMA = true;
constexpr int WS_MAXPTS = 2048;                                               // largest chunk lk_api.cu hands out
constexpr int WS_FB_CAP = ((WS_MAXPTS / 32 + WS_NCONS - 1) / WS_NCONS) * 32;  // a warp's groups of one chunk

// One 272-byte slot per lane: bytes 0..239 of the 256-byte plane record (the fields end at 232), the
// lane's root index at 240, its point at 256.
constexpr int WS_STAGE_BYTES = 32 * TILE_STRIDE;
constexpr int WS_SLOT_ROOT = 240, WS_SLOT_PT = 256;
static_assert(TILE_STRIDE == 272, "slot layout");

struct WsSmem {
    __align__(16) unsigned char st[WS_STAGES][WS_STAGE_BYTES];
    uint64_t full[WS_STAGES];
    uint64_t empty[WS_STAGES];
    ScanConst sc[WS_NPROD + WS_NCONS];  // every warp keeps the constants of the chunk it is working on
    double slice[WS_NCONS * 32];
    uint32_t fb[WS_NCONS][WS_FB_CAP];
    uint32_t nfb[WS_NCONS];
};
static_assert(sizeof(WsSmem) <= 227 * 1024, "ring does not fit");

// Poll loop in C++ on purpose: a branch hidden inside one asm statement leaves the warp's lanes diverged in
// a way the compiler cannot see before the warp-synchronous code that follows. With a debug buffer
// (page-locked host memory) a warp stuck for ~2^20 polls leaves {code, q} in its own record.
template <bool PROF>
__device__ __forceinline__ void ws_wait(uint32_t bar, uint32_t parity, unsigned long long* dbg, uint32_t code, uint32_t q) {
    uint32_t spins = 0;
    while (!mbar_try_wait_s(bar, parity)) {
        if (PROF && dbg && ++spins == (1u << 20) && (threadIdx.x & 31) == 0) {
            unsigned long long* r = dbg + ((size_t)blockIdx.x * 16 + (threadIdx.x >> 5)) * 8;
            r[0] = code; r[1] = q;
            __threadfence_system();
        }
    }
}

__device__ __forceinline__ void load_scan_const(ScanConst* dst, const ScanConst* src, int lane) {
    constexpr int N = (int)(sizeof(ScanConst) / sizeof(double));
    __syncwarp();
    if (lane < N) reinterpret_cast<double*>(dst)[lane] = __ldcg(reinterpret_cast<const double*>(src) + lane);
    __syncwarp();
}

__device__ __forceinline__ ChunkDesc load_chunk(const ChunkDesc* p) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    ChunkDesc cd;
    cd.scan = v.x; cd.start = v.y; cd.count = v.z; cd.pad = v.w;
    return cd;
}

// One group of 32 points somewhere in this block's chunk sequence.
struct Item {
    uint32_t base;  // absolute index of the group's first point
    uint32_t n;     // valid points (1..32), 0 = no item
    uint32_t scan;
    uint32_t su;    // stage | use << 8
};

// The groups of this block's chunks (b, b + grid, ...) numbered q = 0, 1, ...; a producer visits every
// WS_NPROD-th of them and, WS_STAGES being a multiple of WS_NPROD, always the same stages.
struct Walker {
    const ChunkDesc* chunks;
    uint32_t n_chunks, ci, j, G, s, use;
    ChunkDesc cd;
    bool valid;
    __device__ __forceinline__ void begin(const ChunkDesc* c, uint32_t n, uint32_t first) {
        chunks = c; n_chunks = n; ci = blockIdx.x; j = first; G = 0; s = first; use = 0;
        valid = ci < n_chunks;
        if (valid) { cd = load_chunk(chunks + ci); G = (cd.count + 31u) >> 5; settle(); }
    }
    __device__ __forceinline__ void settle() {  // carry j over chunk ends
        while (j >= G) {
            j -= G;
            ci += gridDim.x;
            if (ci >= n_chunks) { valid = false; return; }
            cd = load_chunk(chunks + ci);
            G = (cd.count + 31u) >> 5;
        }
    }
    __device__ __forceinline__ void next() {
        j += WS_NPROD;
        s += WS_NPROD;
        if (s >= (uint32_t)WS_STAGES) { s -= WS_STAGES; ++use; }
        if (valid) settle();
    }
    __device__ __forceinline__ Item item() const {
        Item it;
        it.n = 0; it.base = 0; it.scan = 0; it.su = s | (use << 8);
        if (valid) { it.base = cd.start + j * 32u; it.n = min(32u, cd.count - j * 32u); it.scan = cd.scan; }
        return it;
    }
};

template <bool PROF>
__device__ __forceinline__ void producer_loop(WsSmem* sm, const ResidualArgs& a, uint32_t n_chunks, int p, int lane) {
    const MapView mv = {a.slots, a.hash_mask, a.nodes};
    const Globals& g = a.g;
    ScanConst* sc = &sm->sc[p];
    uint32_t cur_scan = 0xffffffffu;
    const int half = lane >> 4, sub = lane & 15;
    const uint32_t st_base = smem_u32(&sm->st[0][0]), full_base = smem_u32(&sm->full[0]), empty_base = smem_u32(&sm->empty[0]);
    const uint32_t slot_off = (uint32_t)lane * TILE_STRIDE;
    const uint32_t copy_off = (uint32_t)half * TILE_STRIDE + (uint32_t)sub * 16u;
    const unsigned char* nodes_sub = reinterpret_cast<const unsigned char*>(mv.nodes) + sub * 16;
    const bool copier = sub < 15;
    Walker w;
    w.begin(a.chunks + a.chunk_first, n_chunks, (uint32_t)p);
    // pipeline registers: item0 = point load in flight, item1 = probe in flight, item2 = ready to stage
    Item it1, it2;
    it1.n = it2.n = 0; it1.base = it2.base = 0; it1.scan = it2.scan = 0; it1.su = it2.su = 0;
    float4 pt1 = make_float4(0.f, 0.f, 0.f, 0.f), pt2 = pt1;
    SlotPair pair2;
    pair2.a = make_int4(0, 0, 0, -1); pair2.b = pair2.a;
    int kx2 = 0, ky2 = 0, kz2 = 0;
    uint32_t ih2 = 0;
    long long t_begin = PROF ? clock64() : 0, t_empty = 0, t_resolve = 0, t_issue = 0, t_b = 0, n_items = 0;
    for (;;) {
        // A: the next group's points
        const Item it0 = w.item();
        float4 pt0 = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((uint32_t)lane < it0.n) pt0 = __ldg(a.pts + it0.base + lane);
        w.next();
        // C (uses last round's probe): resolve, wait for the stage, hand the records to the copy engine
        if (it2.n) {
            int root = -1;
            long long c0 = PROF ? clock64() : 0;
            if ((uint32_t)lane < it2.n) root = resolve_pair(mv.slots, mv.hash_mask, ih2, pair2, kx2, ky2, kz2);
            __syncwarp();
            long long c1 = PROF ? clock64() : 0;
            const uint32_t s = it2.su & 0xffu, use = it2.su >> 8;
            if (use) ws_wait<PROF>(empty_base + s * 8u, (use - 1u) & 1u, a.wdbg, 1u, it2.su);
            long long c2 = PROF ? clock64() : 0;
            const uint32_t stage = st_base + s * (uint32_t)WS_STAGE_BYTES;
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stage + slot_off + WS_SLOT_PT), "f"(pt2.x), "f"(pt2.y),
                         "f"(pt2.z), "f"(pt2.w) : "memory");
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(stage + slot_off + WS_SLOT_ROOT), "r"(root) : "memory");
            if (WS_TMA) {
                const uint32_t valid = __ballot_sync(0xffffffffu, root >= 0);  // also orders the stores above
                if (lane == 0)
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_base + s * 8u),
                                 "r"(240u * (uint32_t)__popc(valid)) : "memory");
                __syncwarp();
                if (root >= 0)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 240, [%2];" ::"r"(
                                     stage + slot_off),
                                 "l"(mv.nodes + root), "r"(full_base + s * 8u) : "memory");
            } else {
                const uint32_t dst0 = stage + copy_off;
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    const int r = __shfl_sync(0xffffffffu, root, 2 * jj + half);
                    if (r >= 0 && copier) cp_async16_s(dst0 + (uint32_t)jj * 2u * TILE_STRIDE, nodes_sub + ((size_t)(uint32_t)r << 8));
                }
                cp_async_mbar_arrive_noinc_s(full_base + s * 8u);
                mbar_arrive_s(full_base + s * 8u);
            }
            if (PROF) { t_resolve += c1 - c0; t_empty += c2 - c1; ++n_items; t_issue += clock64() - c2; }
        }
        long long cb = PROF ? clock64() : 0;
        // B: key + probe for the group whose points arrived
        it2 = it1; pt2 = pt1;
        if (it1.n) {
            if (it1.scan != cur_scan) { load_scan_const(sc, a.sc + it1.scan, lane); cur_scan = it1.scan; }
            if ((uint32_t)lane < it1.n) {
                PointCtx pc;
                float lx, ly, lz;
                prepare_point(pt1, *sc, g, pc, lx, ly, lz);
                kx2 = (int)lx; ky2 = (int)ly; kz2 = (int)lz;
                ih2 = hash_key(kx2, ky2, kz2) & mv.hash_mask;
                pair2 = load_pair(mv.slots, ih2);
            }
        }
        if (PROF) t_b += clock64() - cb;
        it1 = it0; pt1 = pt0;
        if (!it1.n && !it2.n) break;
    }
    if (PROF && a.wdbg && lane == 0) {
        unsigned long long* r = a.wdbg + ((size_t)blockIdx.x * 16 + p) * 8;
        r[0] = 20; r[1] = 0; r[2] = clock64() - t_begin; r[3] = t_empty; r[4] = t_resolve; r[5] = t_issue; r[6] = t_b; r[7] = n_items;
    }
}

template <bool PROF>
__device__ __forceinline__ void consumer_loop(WsSmem* sm, const ResidualArgs& a, uint32_t n_chunks, int cw, int lane) {
    const MapView mv = {a.slots, a.hash_mask, a.nodes};
    const Globals& g = a.g;
    const int ctid = cw * 32 + lane;
    ScanConst* sc = &sm->sc[WS_NPROD + cw];
    const ChunkDesc* chunks = a.chunks + a.chunk_first;
    const uint32_t full_base = smem_u32(&sm->full[0]), empty_base = smem_u32(&sm->empty[0]);
    uint32_t q0 = 0;
    long long t_begin = PROF ? clock64() : 0, t_full = 0, t_eval = 0, t_barA = 0, t_fb = 0, t_barB = 0;
    for (uint32_t ci = blockIdx.x; ci < n_chunks; ci += gridDim.x) {
        const ChunkDesc cd = load_chunk(chunks + ci);
        const uint32_t G = (cd.count + 31u) >> 5;
        load_scan_const(sc, a.sc + cd.scan, lane);
        double acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.0;
        uint32_t nfbw = 0;
        // my groups of this chunk: sequence numbers congruent to cw
        uint32_t j = ((uint32_t)cw + (uint32_t)WS_NCONS - q0 % (uint32_t)WS_NCONS) % (uint32_t)WS_NCONS;
        uint32_t s = (q0 + j) % (uint32_t)WS_STAGES, use = (q0 + j) / (uint32_t)WS_STAGES;
        for (; j < G; j += WS_NCONS) {
            long long c0 = PROF ? clock64() : 0;
            ws_wait<PROF>(full_base + s * 8u, use & 1u, a.wdbg, 2u, q0 + j);
            long long c1 = PROF ? clock64() : 0;
            const unsigned char* slot = &sm->st[s][0] + (size_t)lane * TILE_STRIDE;
            const int root = *reinterpret_cast<const int*>(slot + WS_SLOT_ROOT);
            bool fail = false;
            if (root >= 0) {
                const float4 pt = *reinterpret_cast<const float4*>(slot + WS_SLOT_PT);
                PointCtx pc;
                float lx, ly, lz;
                prepare_point(pt, *sc, g, pc, lx, ly, lz);
                const uint32_t flags = *reinterpret_cast<const uint32_t*>(slot + 224);
                Row row;
                if ((flags & LK_NODE_IS_PLANE) && eval_plane_staged(slot, pc, *sc, g, row)) accumulate_row(row, acc);
                else fail = true;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, fail);
            if (fail) sm->fb[cw][nfbw + __popc(m & ((1u << lane) - 1u))] = cd.start + j * 32u + (uint32_t)lane;
            nfbw += __popc(m);
            __syncwarp();
            if (lane == 0) mbar_arrive_s(empty_base + s * 8u);
            if (PROF) { t_full += c1 - c0; t_eval += clock64() - c1; }
            s += WS_NCONS;
            if (s >= (uint32_t)WS_STAGES) { s -= WS_STAGES; ++use; }
        }
        q0 += G;
        long long d0 = PROF ? clock64() : 0;
        if (lane == 0) sm->nfb[cw] = nfbw;
        named_barrier_sync<1, WS_CONS_THREADS>();
        long long d1 = PROF ? clock64() : 0;
        // the listed points, all consumers together, in warp-major list order
        {
            uint32_t cnt[WS_NCONS], total = 0;
#pragma unroll
            for (int w2 = 0; w2 < WS_NCONS; ++w2) { cnt[w2] = sm->nfb[w2]; total += cnt[w2]; }
            for (uint32_t e = (uint32_t)ctid; e < total; e += WS_CONS_THREADS) {
                uint32_t k = e;
                int w2 = 0;
#pragma unroll
                for (int t = 0; t < WS_NCONS - 1; ++t)
                    if (w2 == t && k >= cnt[t]) { k -= cnt[t]; w2 = t + 1; }
                Row row;
                if (point_row(__ldg(a.pts + sm->fb[w2][k]), *sc, mv, g, row, nullptr)) accumulate_row(row, acc);
            }
        }
        long long d2 = PROF ? clock64() : 0;
        const double tot = warp_transpose_sum(acc, lane);
        sm->slice[cw * 32 + lane] = tot;
        named_barrier_sync<1, WS_CONS_THREADS>();
        if (cw == 0) {
            double v = 0.0;
#pragma unroll
            for (int w2 = 0; w2 < WS_NCONS; ++w2) v += sm->slice[w2 * 32 + lane];
            a.partial[(size_t)(a.chunk_first + ci) * PARTIAL_STRIDE + lane] = v;
        }
        if (PROF && a.wdbg && lane == 0) {
            unsigned long long* r = a.wdbg + ((size_t)blockIdx.x * 16 + WS_NPROD + cw) * 8;
            t_barA += d1 - d0; t_fb += d2 - d1; t_barB += clock64() - d2;
            r[0] = 12; r[1] = ci; r[2] = clock64() - t_begin; r[3] = t_full; r[4] = t_eval; r[5] = t_barA; r[6] = t_fb; r[7] = t_barB;
        }
    }
}

template <bool PROF>
__global__ void __launch_bounds__(WS_THREADS, 1) k_residual_ws(const __grid_constant__ ResidualArgs a, const uint32_t n_chunks) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    WsSmem* sm = reinterpret_cast<WsSmem*>(s_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < WS_STAGES; ++s) {
            mbar_init(&sm->full[s], WS_TMA ? 1 : 64);  // TMA: one arrive.expect_tx; else 32 copy-completion + 32 plain arrivals
            mbar_init(&sm->empty[s], 1);
        }
    }
    if (tid < WS_NCONS) sm->nfb[tid] = 0;
    mbar_init_fence();
    __syncthreads();
    if (warp < WS_NPROD) {
        warpgroup_reg_dec<WS_PROD_REGS>();
        producer_loop<PROF>(sm, a, n_chunks, warp, lane);
    } else {
        warpgroup_reg_inc<WS_CONS_REGS>();
        consumer_loop<PROF>(sm, a, n_chunks, warp - WS_NPROD, lane);
    }
}

}  // namespace

void launch_residual_ws(const ResidualArgs& a, uint32_t n_chunks, int n_sms, cudaStream_t s) {
    if (n_chunks == 0) return;
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(k_residual_ws<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WsSmem));
        cudaFuncSetAttribute(k_residual_ws<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WsSmem));
        attr = true;
    }
    const uint32_t grid = std::min<uint32_t>(n_chunks, (uint32_t)std::max(1, n_sms));
    if (a.wdbg) k_residual_ws<true><<<grid, WS_THREADS, sizeof(WsSmem), s>>>(a, n_chunks);
    else k_residual_ws<false><<<grid, WS_THREADS, sizeof(WsSmem), s>>>(a, n_chunks);
}

}  // namespace lk
