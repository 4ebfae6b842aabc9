// lk_point.cuh — one LiDAR point through rows a3-a7 of SURVEY.md §8a, and the warp-level
// reduction / 6x6 solve primitives. Shared by the batched multi-kernel path (lk_residual.cu) and
// the fused per-scan persistent kernel (lk_fused.cu).
//
// Follows: a3 KILO.cc:127-140 + voxel_map.cc:22-40, a4 KILO.cc:143-149, a5 voxel_map.cc:363-427,
// a6 KILO.cc:156-178, a7 KILO.cc:187-210. Algebra is restructured (never the results' meaning):
//   * calcBodyCov's A*A^T is range^2 (I - u u^T) because {b1, b2, u} is orthonormal, so
//     n^T M Sigma_b M^T n = rv (u.w)^2 + range^2 dv (|w|^2 - (u.w)^2), w = M^T n;
//   * n^T (R[pi]x) P_tt (R[pi]x)^T n = h_t^T P_tt h_t with h_t = pi x (R^T n), the Jacobian row itself.
#pragma once
#include "lk_device.cuh"

namespace lk {

struct PlaneRec {
    double c[3], n[3], pv[21];
    float d, radius;
    uint32_t flags;
    int child_base;
};

// 15 x 128-bit loads cover the 232 bytes the path needs (lk_map_node). COH = the map may be written by this very kernel
// (persistent per-scan kernel with the map insert inside): read through L2 instead of the non-coherent path.
template <bool COH = false>
__device__ __forceinline__ void load_plane(const MapNode* __restrict__ nd, PlaneRec& r) {
    const double2* q = reinterpret_cast<const double2*>(nd);
    double2 v[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) v[i] = COH ? __ldcg(q + i) : __ldg(q + i);
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

struct PointCtx {
    double pbx, pby, pbz;  // lidar-frame point as calcBodyCov sees it (z == 0 -> 1e-4)
    double pix, piy, piz;  // IMU frame
    double pwx, pwy, pwz;  // world
    double r2;             // |pb|^2
    double range2;         // (double)(float range)^2   (voxel_map.cc:24)
};

struct Row {
    double h[6];
    double z;
    double R;
};

__device__ __forceinline__ double quad_sym3(const double* S, double a, double b, double c) {
    return S[0] * a * a + S[3] * b * b + S[5] * c * c + 2.0 * (S[1] * a * b + S[2] * a * c + S[4] * b * c);
}

// build_single_residual's plane branch (voxel_map.cc:370-411) + the row of KILO.cc:192-209.
__device__ __forceinline__ bool eval_plane(const PlaneRec& r, const PointCtx& pc, const ScanConst& sc,
                                           const Globals& g, bool need_prob, double& prob, Row& row) {
    double s = r.n[0] * pc.pwx + r.n[1] * pc.pwy + r.n[2] * pc.pwz + (double)r.d;
    float dis = (float)fabs(s);
    double ax = pc.pwx - r.c[0], ay = pc.pwy - r.c[1], az = pc.pwz - r.c[2];
    float dc = (float)(ax * ax + ay * ay + az * az);
    float rd = sqrtf(__fsub_rn(dc, __fmul_rn(dis, dis)));  // float arithmetic as in the reference
    if (!((double)rd <= 3.0 * (double)r.radius)) return false;

    // J_nq Sigma_plane J_nq^T, J_nq = [(pw - c)^T, -n^T]
    const double J0 = ax, J1 = ay, J2 = az, J3 = -r.n[0], J4 = -r.n[1], J5 = -r.n[2];
    const double* pv = r.pv;
    double t0 = pv[0] * J0 + 2.0 * (pv[1] * J1 + pv[2] * J2 + pv[3] * J3 + pv[4] * J4 + pv[5] * J5);
    double t1 = pv[6] * J1 + 2.0 * (pv[7] * J2 + pv[8] * J3 + pv[9] * J4 + pv[10] * J5);
    double t2 = pv[11] * J2 + 2.0 * (pv[12] * J3 + pv[13] * J4 + pv[14] * J5);
    double t3 = pv[15] * J3 + 2.0 * (pv[16] * J4 + pv[17] * J5);
    double t4 = pv[18] * J4 + 2.0 * (pv[19] * J5);
    double t5 = pv[20] * J5;
    double sigma_pl = J0 * t0 + J1 * t1 + J2 * t2 + J3 * t3 + J4 * t4 + J5 * t5;

    // q = R^T n ; h_theta = pi x q ; w = (R Re)^T n = Re^T q
    double qx = sc.R[0] * r.n[0] + sc.R[3] * r.n[1] + sc.R[6] * r.n[2];
    double qy = sc.R[1] * r.n[0] + sc.R[4] * r.n[1] + sc.R[7] * r.n[2];
    double qz = sc.R[2] * r.n[0] + sc.R[5] * r.n[1] + sc.R[8] * r.n[2];
    double hx = pc.piy * qz - pc.piz * qy;
    double hy = pc.piz * qx - pc.pix * qz;
    double hz = pc.pix * qy - pc.piy * qx;
    double wx = g.Re[0] * qx + g.Re[3] * qy + g.Re[6] * qz;
    double wy = g.Re[1] * qx + g.Re[4] * qy + g.Re[7] * qz;
    double wz = g.Re[2] * qx + g.Re[5] * qy + g.Re[8] * qz;
    double uw = pc.pbx * wx + pc.pby * wy + pc.pbz * wz;
    double ww = wx * wx + wy * wy + wz * wz;
    double uw2 = uw * uw / pc.r2;  // (u.w)^2
    double body = (double)g.rv * uw2 + pc.range2 * g.dv * (ww - uw2);
    double state = quad_sym3(sc.Pth, hx, hy, hz) + quad_sym3(sc.Ppp, r.n[0], r.n[1], r.n[2]);
    double sigma_l = sigma_pl + body + state;

    // gate 2: dis_to_plane < sigma_num * sqrt(sigma_l)   (voxel_map.cc:387), squared with an exact
    // fallback at the boundary so the decision equals the reference's comparison.
    double lhs = (double)dis * (double)dis;
    double rhs = g.sigma_num * g.sigma_num * sigma_l;
    bool pass;
    if (lhs < rhs * (1.0 - 1e-12)) pass = true;
    else if (lhs > rhs * (1.0 + 1e-12)) pass = false;
    else pass = (double)dis < g.sigma_num * sqrt(sigma_l);
    if (!pass) return false;
    if (need_prob) {
        double this_prob = 1.0 / sqrt(sigma_l) * exp(-0.5 * (double)dis * (double)dis / sigma_l);
        if (!(this_prob > prob)) return true;  // is_success without replacing the candidate
        prob = this_prob;
    }
    row.h[0] = hx; row.h[1] = hy; row.h[2] = hz;
    row.h[3] = r.n[0]; row.h[4] = r.n[1]; row.h[5] = r.n[2];
    row.z = -(double)(float)s;  // dis_to_plane_ is float (voxel_map.h:92)
    row.R = g.ratio * (sigma_pl + body);
    return true;
}

__device__ __forceinline__ int map_find(const HashSlot* __restrict__ slots, uint32_t mask, int kx, int ky, int kz) {
    uint32_t i = hash_key(kx, ky, kz) & mask;
    for (;;) {
        int4 s = __ldg(reinterpret_cast<const int4*>(slots + i));
        if (s.w < 0) return -1;
        if (s.x == kx && s.y == ky && s.z == kz) return s.w;
        i = (i + 1) & mask;
    }
}

// Rare path of build_single_residual (voxel_map.cc:412-424): the root is not a plane, so every
// initialised plane among ALL children of non-plane nodes down to max_layer is a candidate and the
// most probable one wins. Kept out of line so the common path does not carry its registers.
template <bool COH = false>
static __device__ __noinline__ bool visit_subtree(const MapNode* __restrict__ nodes, int child_base, uint32_t cmask,
                                           const PointCtx* pcp, const ScanConst* scp, const Globals* gp, double* probp,
                                           Row* rowp) {
    const PointCtx& pc = *pcp;
    const ScanConst& sc = *scp;
    const Globals& g = *gp;
    bool ok = false;
    int st_base[4];
    uint32_t st_mask[4];
    int sp = 1;
    st_base[0] = child_base;
    st_mask[0] = cmask;
    double prob = *probp;
    Row row = *rowp;
    while (sp > 0) {
        uint32_t m = st_mask[sp - 1];
        if (m == 0) { --sp; continue; }
        int c = __ffs(m) - 1;  // child order 0..7 as the reference's loop
        st_mask[sp - 1] = m & (m - 1);
        int layer = sp;  // children of a layer-(sp-1) node
        PlaneRec cr;
        load_plane<COH>(nodes + st_base[sp - 1] + c, cr);
        if (cr.flags & LK_NODE_IS_PLANE) {
            if (eval_plane(cr, pc, sc, g, true, prob, row)) ok = true;
        } else if (layer < g.max_layer && sp < 4) {
            uint32_t cm = (cr.flags >> LK_NODE_CHILDMASK_SHIFT) & 0xffu;
            if (cr.child_base >= 0 && cm) {
                st_base[sp] = cr.child_base;
                st_mask[sp] = cm;
                ++sp;
            }
        }
    }
    *probp = prob;
    *rowp = row;
    return ok;
}

// One point through rows a3-a7. Returns true when a residual row was produced.
__device__ __forceinline__ bool point_row(float4 pt, const ScanConst& sc, const MapView& a, const Globals& g, Row& row,
                                          int* key_out) {
    PointCtx pc;
    double bx = (double)pt.x, by = (double)pt.y, bz = (double)pt.z;
    pc.pix = g.Re[0] * bx + g.Re[1] * by + g.Re[2] * bz + g.te[0];
    pc.piy = g.Re[3] * bx + g.Re[4] * by + g.Re[5] * bz + g.te[1];
    pc.piz = g.Re[6] * bx + g.Re[7] * by + g.Re[8] * bz + g.te[2];
    pc.pwx = sc.R[0] * pc.pix + sc.R[1] * pc.piy + sc.R[2] * pc.piz + sc.p[0];
    pc.pwy = sc.R[3] * pc.pix + sc.R[4] * pc.piy + sc.R[5] * pc.piz + sc.p[1];
    pc.pwz = sc.R[6] * pc.pix + sc.R[7] * pc.piy + sc.R[8] * pc.piz + sc.p[2];

    // voxel key: float quotient, -1 shift for negatives, truncation (KILO.cc:143-148)
    float lx, ly, lz;
    if (g.voxel_pow2) {
        lx = (float)(pc.pwx * g.inv_voxel); ly = (float)(pc.pwy * g.inv_voxel); lz = (float)(pc.pwz * g.inv_voxel);
    } else {
        lx = (float)(pc.pwx / g.voxel); ly = (float)(pc.pwy / g.voxel); lz = (float)(pc.pwz / g.voxel);
    }
    if (lx < 0) lx = (float)((double)lx - 1.0);
    if (ly < 0) ly = (float)((double)ly - 1.0);
    if (lz < 0) lz = (float)((double)lz - 1.0);
    const int kx = (int)lx, ky = (int)ly, kz = (int)lz;
    if (key_out) { key_out[0] = kx; key_out[1] = ky; key_out[2] = kz; }
    int root = map_find(a.slots, a.hash_mask, kx, ky, kz);  // issue the probe before the fp64 work below
    if (root < 0) return false;

    // calcBodyCov mutates pb.z AFTER pi / pw were formed (voxel_map.cc:23, KILO.cc:134)
    pc.pbx = bx; pc.pby = by; pc.pbz = (bz == 0.0) ? 0.0001 : bz;
    pc.r2 = pc.pbx * pc.pbx + pc.pby * pc.pby + pc.pbz * pc.pbz;
    float range = (float)sqrt(pc.r2);
    pc.range2 = (double)range * (double)range;

    double prob = 0.0;
    bool ok = false;
    int nx = kx, ny = ky, nz = kz;
    // home voxel first; on failure ONE (possibly diagonal) neighbour (KILO.cc:156-178)
#pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
        PlaneRec r;
        load_plane(a.nodes + root, r);
        if (r.flags & LK_NODE_IS_PLANE) {
            ok = eval_plane(r, pc, sc, g, false, prob, row);
        } else {
            uint32_t cmask = (r.flags >> LK_NODE_CHILDMASK_SHIFT) & 0xffu;
            if (g.max_layer >= 1 && r.child_base >= 0 && cmask)
                ok = visit_subtree(a.nodes, r.child_base, cmask, &pc, &sc, &g, &prob, &row);
        }
        if (ok || attempt == 1) break;
        // loc in VOXEL units against a centre in METRES: the reference's own unit mismatch
        double q = (double)(g.voxel_f / 4.0f);
        double cx = (0.5 + kx) * (double)g.voxel_f, cy = (0.5 + ky) * (double)g.voxel_f, cz = (0.5 + kz) * (double)g.voxel_f;
        if ((double)lx > cx + q) nx++; else if ((double)lx < cx - q) nx--;
        if ((double)ly > cy + q) ny++; else if ((double)ly < cy - q) ny--;
        if ((double)lz > cz + q) nz++; else if ((double)lz < cz - q) nz--;
        if (nx == kx && ny == ky && nz == kz) break;
        root = map_find(a.slots, a.hash_mask, nx, ny, nz);
        if (root < 0) break;
    }
    return ok;
}

// Sum of 32 per-lane values over the warp with value/lane transposition: after the 5 exchange
// steps lane L holds the warp total of value L (31 exchanges instead of 32 x 5 shuffles).
__device__ __forceinline__ double warp_transpose_sum(double (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < off; ++i) {
            double send = upper ? v[i] : v[i + off];
            double keep = upper ? v[i + off] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// 6 x 6 solve [M | b | A] -> [I | y | W] by Gauss-Jordan with partial pivoting, one column per
// lane (lanes 0..12), executed by one full warp. Returns false on a zero pivot.
__device__ __forceinline__ bool warp_solve6(double (&col)[6], int lane) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double ck[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) ck[i] = __shfl_sync(0xffffffffu, col[i], k);
        int piv = k;
        double best = fabs(ck[k]);
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
            double a = fabs(ck[i]);
            if (a > best) { best = a; piv = i; }
        }
        if (best == 0.0) ok = false;
#pragma unroll
        for (int i = k + 1; i < 6; ++i)
            if (piv == i) {
                double t = col[k]; col[k] = col[i]; col[i] = t;
                t = ck[k]; ck[k] = ck[i]; ck[i] = t;
            }
        const double inv = 1.0 / ck[k];
        col[k] *= inv;
#pragma unroll
        for (int i = 0; i < 6; ++i)
            if (i != k) col[i] -= ck[i] * col[k];
    }
    return ok;
}

}  // namespace lk
