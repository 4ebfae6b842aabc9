// lk_pass.cuh — one block-wide pass over up to NTHREADS points (one point per thread):
//   1. transform, voxel key, home probe AND the speculative probe of the one neighbour voxel the
//      reference falls back to (KILO.cc:156-178) — both 16-byte table reads are in flight together;
//   2. every lane stages its 256-byte plane record into shared memory with one TMA bulk copy
//      (cp.async.bulk global -> shared, completion on the warp's mbarrier) instead of 15 scattered
//      128-bit loads per lane (32 cache lines per warp-instruction);
//   3. gates + residual row from the staged record (conflict-free 128-bit shared loads, 272-byte
//      slot stride);
//   4. points whose home voxel gave no residual are compacted into a block-wide list and their
//      neighbour voxel is evaluated by the first threads of the block — a few percent of the
//      points fail, but almost every warp holds one, so without compaction every warp would pay
//      the second round.
#pragma once
#include "lk_async.cuh"
#include "lk_point.cuh"

namespace lk {

constexpr int TILE_STRIDE = 272;  // 256-byte record + 16: 128-bit reads of 8 consecutive lanes hit 32 banks

template <int NTHREADS>
struct PassSmem {
    __align__(16) unsigned char tile[NTHREADS * TILE_STRIDE];
    struct __align__(8) Fallback {
        double pc[11];
        int near;
        uint32_t idx;
    } fb[NTHREADS];
    uint64_t bar[NTHREADS / 32];
    uint32_t wcnt[NTHREADS / 32];  // fallback entries of each warp (its region starts at fb[warp * 32])
};

struct DebugRows {  // lk_debug_residuals outputs (nullable)
    uint8_t* ok;
    double* h;
    double* z;
    double* R;
    int32_t* key;
};

__device__ __forceinline__ void plane_from_smem(const unsigned char* slot, PlaneRec& r) {
    const double2* q = reinterpret_cast<const double2*>(slot);
    double2 v[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) v[i] = q[i];
    r.c[0] = v[0].x; r.c[1] = v[0].y; r.c[2] = v[1].x;
    r.n[0] = v[1].y; r.n[1] = v[2].x; r.n[2] = v[2].y;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        r.pv[2 * i] = v[3 + i].x;
        r.pv[2 * i + 1] = v[3 + i].y;
    }
    r.pv[20] = v[13].x;
    long long dr = __double_as_longlong(v[13].y);
    r.d = __int_as_float((int)(dr & 0xffffffffll));
    r.radius = __int_as_float((int)(dr >> 32));
    long long fc = __double_as_longlong(v[14].x);
    r.flags = (uint32_t)(fc & 0xffffffffll);
    r.child_base = (int)(fc >> 32);
}

__device__ __forceinline__ void accumulate_row(const Row& row, double (&acc)[32]) {
    const double w = 1.0 / row.R;
    int q = 0;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
        const double hw = row.h[r] * w;
#pragma unroll
        for (int c = r; c < 6; ++c) acc[q++] += hw * row.h[c];
        acc[ACC_B + r] += hw * row.z;
    }
    acc[ACC_SUMR] += row.R;
    acc[ACC_CNT] += 1.0;
}

// Call once per block before the first pass (all threads).
template <int NTHREADS>
__device__ __forceinline__ void pass_init(PassSmem<NTHREADS>* ps) {
    const int tid = threadIdx.x;
    if ((tid & 31) == 0) mbar_init(&ps->bar[tid >> 5], 1);
    mbar_init_fence();
    __syncthreads();
}

// Linear probing, two slots per step: an even-aligned pair of 16-byte slots shares one 32-byte
// sector, so the second slot is free. `pre` holds the pair at the key's home position (already
// loaded by the caller so several lookups can be in flight together).
struct SlotPair {
    int4 a, b;
};
template <bool COH = false>
__device__ __forceinline__ SlotPair load_pair(const HashSlot* __restrict__ slots, uint32_t i) {
    SlotPair p;
    const int4* q = reinterpret_cast<const int4*>(slots + (i & ~1u));
    p.a = COH ? __ldcg(q) : __ldg(q);
    p.b = COH ? __ldcg(q + 1) : __ldg(q + 1);
    return p;
}
// A slot whose node is -2 / -3 is being published / was poisoned by an insert (lk_insert.cuh); both are negative, so
// the probe treats them as "end of the chain" — the persistent kernel only probes between inserts, never during one.
template <bool COH = false>
__device__ __forceinline__ int resolve_pair(const HashSlot* __restrict__ slots, uint32_t mask, uint32_t i, SlotPair p,
                                            int kx, int ky, int kz) {
    // first step may start on the odd slot of its pair
    if ((i & 1u) == 0) {
        if (p.a.w < 0) return -1;
        if (p.a.x == kx && p.a.y == ky && p.a.z == kz) return p.a.w;
    }
    if (p.b.w < 0) return -1;
    if (p.b.x == kx && p.b.y == ky && p.b.z == kz) return p.b.w;
    uint32_t j = ((i & ~1u) + 2) & mask;
    for (uint32_t probes = 0;; ++probes) {
        if (probes > mask) {  // walked the whole table without meeting an empty slot (see lk_stall_note)
            stall_note(4u, i);
            return -1;
        }
        p = load_pair<COH>(slots, j);
        if (p.a.w < 0) return -1;
        if (p.a.x == kx && p.a.y == ky && p.a.z == kz) return p.a.w;
        if (p.b.w < 0) return -1;
        if (p.b.x == kx && p.b.y == ky && p.b.z == kz) return p.b.w;
        j = (j + 2) & mask;
    }
}

// voxel key in float: quotient, -1 shift for negatives; the caller truncates (KILO.cc:143-148)
__device__ __forceinline__ void voxel_loc(const PointCtx& pc, const Globals& g, float& lx, float& ly, float& lz) {
    if (g.voxel_pow2) {
        lx = (float)(pc.pwx * g.inv_voxel); ly = (float)(pc.pwy * g.inv_voxel); lz = (float)(pc.pwz * g.inv_voxel);
    } else {
        lx = (float)(pc.pwx / g.voxel); ly = (float)(pc.pwy / g.voxel); lz = (float)(pc.pwz / g.voxel);
    }
    if (lx < 0) lx = (float)((double)lx - 1.0);
    if (ly < 0) ly = (float)((double)ly - 1.0);
    if (lz < 0) lz = (float)((double)lz - 1.0);
}

// the ONE neighbour the reference falls back to: loc in VOXEL units against a centre in METRES (the reference's
// own unit mismatch, KILO.cc:158-172)
__device__ __forceinline__ void neighbour_key(const Globals& g, float lx, float ly, float lz, int kx, int ky, int kz, int& nx,
                                              int& ny, int& nz) {
    const double q = (double)(g.voxel_f / 4.0f);
    const double cx = (0.5 + kx) * (double)g.voxel_f, cy = (0.5 + ky) * (double)g.voxel_f, cz = (0.5 + kz) * (double)g.voxel_f;
    nx = kx; ny = ky; nz = kz;
    if ((double)lx > cx + q) nx++; else if ((double)lx < cx - q) nx--;
    if ((double)ly > cy + q) ny++; else if ((double)ly < cy - q) ny--;
    if ((double)lz > cz + q) nz++; else if ((double)lz < cz - q) nz--;
}

__device__ __forceinline__ void prepare_point(float4 pt, const ScanConst& sc, const Globals& g, PointCtx& pc, float& lx,
                                              float& ly, float& lz) {
    const double bx = (double)pt.x, by = (double)pt.y, bz = (double)pt.z;
    pc.pix = g.Re[0] * bx + g.Re[1] * by + g.Re[2] * bz + g.te[0];
    pc.piy = g.Re[3] * bx + g.Re[4] * by + g.Re[5] * bz + g.te[1];
    pc.piz = g.Re[6] * bx + g.Re[7] * by + g.Re[8] * bz + g.te[2];
    pc.pbx = bx; pc.pby = by; pc.pbz = (bz == 0.0) ? 0.0001 : bz;  // calcBodyCov's mutation (voxel_map.cc:23)
    pc.r2 = pc.pbx * pc.pbx + pc.pby * pc.pby + pc.pbz * pc.pbz;
    const float range = (float)sqrt(pc.r2);
    pc.range2 = (double)range * (double)range;
    pc.pwx = sc.R[0] * pc.pix + sc.R[1] * pc.piy + sc.R[2] * pc.piz + sc.p[0];
    pc.pwy = sc.R[3] * pc.pix + sc.R[4] * pc.piy + sc.R[5] * pc.piz + sc.p[1];
    pc.pwz = sc.R[6] * pc.pix + sc.R[7] * pc.piy + sc.R[8] * pc.piz + sc.p[2];
    voxel_loc(pc, g, lx, ly, lz);
}

// Points that produced no residual at home are listed per warp in lane order (no atomics: the order, hence the
// sums, are reproducible); entry `tid` of the warp-major concatenation of the lists is handled by thread `tid`.
template <int NTHREADS, class PS>
__device__ __forceinline__ void fallback_list(PS* ps, bool want, const PointCtx& pc, int near, int lane, int warp) {
    const uint32_t m = __ballot_sync(0xffffffffu, want);
    if (lane == 0) ps->wcnt[warp] = (uint32_t)__popc(m);
    if (want) {
        const uint32_t slot = (uint32_t)warp * 32u + (uint32_t)__popc(m & ((1u << lane) - 1u));
        auto& f = ps->fb[slot];
        f.pc[0] = pc.pbx; f.pc[1] = pc.pby; f.pc[2] = pc.pbz; f.pc[3] = pc.pix; f.pc[4] = pc.piy; f.pc[5] = pc.piz;
        f.pc[6] = pc.pwx; f.pc[7] = pc.pwy; f.pc[8] = pc.pwz; f.pc[9] = pc.r2; f.pc[10] = pc.range2;
        f.near = near;
        f.idx = (uint32_t)threadIdx.x;
    }
}
template <int NTHREADS, class PS>
__device__ __forceinline__ bool fallback_pick(const PS* ps, uint32_t& fb_slot) {
    uint32_t n_fb = 0, k = threadIdx.x;
    bool found = false;
#pragma unroll
    for (int w = 0; w < NTHREADS / 32; ++w) {
        const uint32_t c = ps->wcnt[w];
        if (!found && k < c) { fb_slot = (uint32_t)w * 32u + k; found = true; }
        if (!found) k -= c;
        n_fb += c;
    }
    return threadIdx.x < n_fb;
}

template <bool COH = false>
__device__ __forceinline__ bool eval_record(const MapNode* __restrict__ nodes, const PlaneRec& r, const PointCtx& pc,
                                            const ScanConst& sc, const Globals& g, Row& row) {
    double prob = 0.0;
    if (r.flags & LK_NODE_IS_PLANE) return eval_plane(r, pc, sc, g, false, prob, row);
    const uint32_t cmask = (r.flags >> LK_NODE_CHILDMASK_SHIFT) & 0xffu;
    if (g.max_layer >= 1 && r.child_base >= 0 && cmask) return visit_subtree<COH>(nodes, r.child_base, cmask, &pc, &sc, &g, &prob, &row);
    return false;
}

// One pass, every point looked up afresh (multi-kernel path). `phase` is the warp's mbarrier parity (start at 0,
// carried between passes). base_idx = absolute index of pts[0] (debug output addressing).
template <int NTHREADS, bool DEBUG>
__device__ __forceinline__ void block_points_pass(PassSmem<NTHREADS>* ps, uint32_t& phase, const float4* __restrict__ pts,
                                                  uint32_t count, size_t base_idx, const ScanConst& sc, const MapView& mv,
                                                  const Globals& g, double (&acc)[32], const DebugRows& dbg) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool active = (uint32_t)tid < count;
    PointCtx pc;
    int root = -1, near = -1;
    int key[3] = {0, 0, 0};
    if (active) {
        float lx, ly, lz;
        prepare_point(__ldg(pts + tid), sc, g, pc, lx, ly, lz);
        const int kx = (int)lx, ky = (int)ly, kz = (int)lz;
        key[0] = kx; key[1] = ky; key[2] = kz;
        int nx, ny, nz;
        neighbour_key(g, lx, ly, lz, kx, ky, kz, nx, ny, nz);
        const bool differs = (nx != kx) || (ny != ky) || (nz != kz);
        // both home pairs are read before either is inspected
        const uint32_t ih = hash_key(kx, ky, kz) & mv.hash_mask, in = hash_key(nx, ny, nz) & mv.hash_mask;
        const SlotPair sh = load_pair(mv.slots, ih);
        const SlotPair sn = load_pair(mv.slots, in);
        root = resolve_pair(mv.slots, mv.hash_mask, ih, sh, kx, ky, kz);
        // the reference only looks at the neighbour when the home voxel exists
        near = (root >= 0 && differs) ? resolve_pair(mv.slots, mv.hash_mask, in, sn, nx, ny, nz) : -1;
    }
    // ---- stage the home records: one bulk copy per lane -------------------------------------------
    unsigned char* my_slot = ps->tile + (size_t)tid * TILE_STRIDE;
    const bool gather = root >= 0;
    const uint32_t valid = __ballot_sync(0xffffffffu, gather);
    if (valid) {
        if (lane == 0) mbar_expect_tx(&ps->bar[warp], 256u * (uint32_t)__popc(valid));
        __syncwarp();
        if (gather) bulk_g2s(my_slot, mv.nodes + root, 256u, &ps->bar[warp]);
        mbar_wait(&ps->bar[warp], phase);
        phase ^= 1u;
    }
    // ---- gates + row -----------------------------------------------------------------------------
    Row row;
    bool ok = false;
    if (root >= 0) {
        PlaneRec r;
        plane_from_smem(my_slot, r);
        ok = eval_record(mv.nodes, r, pc, sc, g, row);
    }
    fallback_list<NTHREADS>(ps, root >= 0 && !ok && near >= 0, pc, near, lane, warp);
    if (DEBUG) {
        if (active) {
            const size_t gi = base_idx + tid;
            dbg.ok[gi] = ok ? 1 : 0;
            for (int k = 0; k < 3; ++k) dbg.key[gi * 3 + k] = key[k];
            for (int k = 0; k < 6; ++k) dbg.h[gi * 6 + k] = ok ? row.h[k] : 0.0;
            dbg.z[gi] = ok ? row.z : 0.0;
            dbg.R[gi] = ok ? row.R : 0.0;
        }
    }
    __syncthreads();
    // ---- fallback round: the neighbour voxel of the points that failed at home ------------------------
    uint32_t fb_slot = 0;
    Row row2;
    bool ok2 = false;
    if (fallback_pick<NTHREADS>(ps, fb_slot)) {
        const typename PassSmem<NTHREADS>::Fallback& f = ps->fb[fb_slot];
        PointCtx fc;
        fc.pbx = f.pc[0]; fc.pby = f.pc[1]; fc.pbz = f.pc[2]; fc.pix = f.pc[3]; fc.piy = f.pc[4]; fc.piz = f.pc[5];
        fc.pwx = f.pc[6]; fc.pwy = f.pc[7]; fc.pwz = f.pc[8]; fc.r2 = f.pc[9]; fc.range2 = f.pc[10];
        PlaneRec r;
        load_plane(mv.nodes + f.near, r);
        ok2 = eval_record(mv.nodes, r, fc, sc, g, row2);
        if (ok2 && DEBUG) {
            const size_t gi = base_idx + f.idx;
            dbg.ok[gi] = 1;
            for (int k = 0; k < 6; ++k) dbg.h[gi * 6 + k] = row2.h[k];
            dbg.z[gi] = row2.z;
            dbg.R[gi] = row2.R;
        }
    }
    if (!DEBUG) {  // rows are folded in only now, so no accumulator is live across the evaluations
        if (ok) accumulate_row(row, acc);
        if (ok2) accumulate_row(row2, acc);
    }
    __syncthreads();
}

// =================================================================================================
// Cached pass of the fused per-scan kernel (one chunk per block, so a lane sees the same point in every iteration
// of a bucket). A lane keeps everything that does not depend on the state, plus the last voxel key with its lookup
// results; BOTH candidate records — the home voxel's and the one neighbour voxel's the reference falls back to
// (KILO.cc:156-178) — are staged into shared memory by TMA bulk copies issued together, and stay there: when the key
// is unchanged in a later iteration (the map is static within a bucket) the probes AND the gathers are skipped, and
// the fallback round reads its record from shared memory instead of paying another dependent global round trip.
// Arithmetic and accumulation order are those of block_points_pass (bitwise-equal sums).
// =================================================================================================
struct LaneCache {
    double pbx, pby, pbz, pix, piy, piz, r2, range2;
    int kx, ky, kz, nx, ny, nz, root, near;
    int have;  // 0 = nothing cached, 1 = point quantities cached, 2 = + keys / root / near / staged records
};

template <int NTHREADS>
struct CachedPassSmem {
    __align__(16) unsigned char tile[2][NTHREADS * TILE_STRIDE];  // [0] home records, [1] neighbour records
    struct __align__(8) Fallback {
        double pc[11];
        int near;
        uint32_t idx;
    } fb[NTHREADS];
    uint64_t bar[NTHREADS / 32];
    uint32_t wcnt[NTHREADS / 32];
};

template <int NTHREADS>
__device__ __forceinline__ void cached_pass_init(CachedPassSmem<NTHREADS>* ps) {
    const int tid = threadIdx.x;
    if ((tid & 31) == 0) mbar_init(&ps->bar[tid >> 5], 1);
    mbar_init_fence();
    __syncthreads();
}

template <int NTHREADS, bool COH = false>
__device__ __forceinline__ void cached_points_pass(CachedPassSmem<NTHREADS>* ps, uint32_t& phase, uint32_t count,
                                                   const ScanConst& sc, const MapView& mv, const Globals& g,
                                                   double (&acc)[32], LaneCache& lc, float4 pre) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool active = (uint32_t)tid < count;
    PointCtx pc;
    int root = -1, near = -1;
    bool gather_home = false, gather_near = false;
    if (active) {
        if (lc.have == 0) {
            const double bx = (double)pre.x, by = (double)pre.y, bz = (double)pre.z;
            lc.pix = g.Re[0] * bx + g.Re[1] * by + g.Re[2] * bz + g.te[0];
            lc.piy = g.Re[3] * bx + g.Re[4] * by + g.Re[5] * bz + g.te[1];
            lc.piz = g.Re[6] * bx + g.Re[7] * by + g.Re[8] * bz + g.te[2];
            // calcBodyCov mutates pb.z AFTER pi / pw were formed (voxel_map.cc:23, KILO.cc:134)
            lc.pbx = bx; lc.pby = by; lc.pbz = (bz == 0.0) ? 0.0001 : bz;
            lc.r2 = lc.pbx * lc.pbx + lc.pby * lc.pby + lc.pbz * lc.pbz;
            const float range = (float)sqrt(lc.r2);
            lc.range2 = (double)range * (double)range;
            lc.have = 1;
        }
        pc.pbx = lc.pbx; pc.pby = lc.pby; pc.pbz = lc.pbz; pc.pix = lc.pix; pc.piy = lc.piy; pc.piz = lc.piz;
        pc.r2 = lc.r2; pc.range2 = lc.range2;
        pc.pwx = sc.R[0] * pc.pix + sc.R[1] * pc.piy + sc.R[2] * pc.piz + sc.p[0];
        pc.pwy = sc.R[3] * pc.pix + sc.R[4] * pc.piy + sc.R[5] * pc.piz + sc.p[1];
        pc.pwz = sc.R[6] * pc.pix + sc.R[7] * pc.piy + sc.R[8] * pc.piz + sc.p[2];
        float lx, ly, lz;
        voxel_loc(pc, g, lx, ly, lz);
        const int kx = (int)lx, ky = (int)ly, kz = (int)lz;
        int nx, ny, nz;
        neighbour_key(g, lx, ly, lz, kx, ky, kz, nx, ny, nz);
        const bool differs = (nx != kx) || (ny != ky) || (nz != kz);
        const bool same_home = lc.have == 2 && lc.kx == kx && lc.ky == ky && lc.kz == kz;
        if (same_home) {
            root = lc.root;
            near = lc.near;
            if (lc.nx != nx || lc.ny != ny || lc.nz != nz) {  // same home voxel, different neighbour: redo that lookup only
                const uint32_t in = hash_key(nx, ny, nz) & mv.hash_mask;
                near = (root >= 0 && differs) ? resolve_pair<COH>(mv.slots, mv.hash_mask, in, load_pair<COH>(mv.slots, in), nx, ny, nz) : -1;
                gather_near = near >= 0;
            }
        } else {
            // both home pairs are read before either is inspected
            const uint32_t ih = hash_key(kx, ky, kz) & mv.hash_mask, in = hash_key(nx, ny, nz) & mv.hash_mask;
            const SlotPair sh = load_pair<COH>(mv.slots, ih);
            const SlotPair sn = load_pair<COH>(mv.slots, in);
            root = resolve_pair<COH>(mv.slots, mv.hash_mask, ih, sh, kx, ky, kz);
            // the reference only looks at the neighbour when the home voxel exists
            near = (root >= 0 && differs) ? resolve_pair<COH>(mv.slots, mv.hash_mask, in, sn, nx, ny, nz) : -1;
            gather_home = root >= 0;
            gather_near = near >= 0;
        }
        lc.kx = kx; lc.ky = ky; lc.kz = kz; lc.nx = nx; lc.ny = ny; lc.nz = nz;
        lc.root = root; lc.near = near;
        lc.have = 2;
    }
    // ---- stage the records: one bulk copy per lane and record ---------------------------------------
    unsigned char* home_slot = ps->tile[0] + (size_t)tid * TILE_STRIDE;
    unsigned char* near_slot = ps->tile[1] + (size_t)tid * TILE_STRIDE;
    const uint32_t vh = __ballot_sync(0xffffffffu, gather_home), vn = __ballot_sync(0xffffffffu, gather_near);
    if (vh | vn) {
        if (lane == 0) mbar_expect_tx(&ps->bar[warp], 256u * (uint32_t)(__popc(vh) + __popc(vn)));
        __syncwarp();
        if (gather_home) bulk_g2s(home_slot, mv.nodes + root, 256u, &ps->bar[warp]);
        if (gather_near) bulk_g2s(near_slot, mv.nodes + near, 256u, &ps->bar[warp]);
        mbar_wait(&ps->bar[warp], phase);
        phase ^= 1u;
    }
    // ---- gates + row -----------------------------------------------------------------------------
    Row row;
    bool ok = false;
    if (root >= 0) {
        PlaneRec r;
        plane_from_smem(home_slot, r);
        ok = eval_record<COH>(mv.nodes, r, pc, sc, g, row);
    }
    fallback_list<NTHREADS>(ps, root >= 0 && !ok && near >= 0, pc, near, lane, warp);
    __syncthreads();
    // ---- fallback round: the neighbour voxel of the points that failed at home, record already staged -------
    uint32_t fb_slot = 0;
    Row row2;
    bool ok2 = false;
    if (fallback_pick<NTHREADS>(ps, fb_slot)) {
        const typename CachedPassSmem<NTHREADS>::Fallback& f = ps->fb[fb_slot];
        PointCtx fc;
        fc.pbx = f.pc[0]; fc.pby = f.pc[1]; fc.pbz = f.pc[2]; fc.pix = f.pc[3]; fc.piy = f.pc[4]; fc.piz = f.pc[5];
        fc.pwx = f.pc[6]; fc.pwy = f.pc[7]; fc.pwz = f.pc[8]; fc.r2 = f.pc[9]; fc.range2 = f.pc[10];
        PlaneRec r;
        plane_from_smem(ps->tile[1] + (size_t)f.idx * TILE_STRIDE, r);
        ok2 = eval_record<COH>(mv.nodes, r, fc, sc, g, row2);
    }
    if (ok) accumulate_row(row, acc);
    if (ok2) accumulate_row(row2, acc);
    __syncthreads();  // the fallback list is rewritten by the next pass
}


}  // namespace lk
