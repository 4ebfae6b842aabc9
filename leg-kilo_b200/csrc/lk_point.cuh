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

// The 232 bytes of a plane record (lk_map_node) the path reads, addressed in place — in shared memory where a record was
// staged, in global memory on the rare paths: centre 3 | normal 3 | Sigma_plane upper triangle 21 doubles, then
// {d, radius} as floats (double index 27), flags (byte 224), child_base (byte 228).
__device__ __forceinline__ uint32_t rec_flags(const void* rec) { return *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(rec) + 224); }
__device__ __forceinline__ int rec_child_base(const void* rec) { return *reinterpret_cast<const int*>(reinterpret_cast<const char*>(rec) + 228); }

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

// build_single_residual's plane branch (voxel_map.cc:370-411) + the row of KILO.cc:192-209, reading the record field
// by field (no 29-double register image): centre and normal first, the cheap float gate, and only then the 21
// plane-covariance terms. ONE evaluation for every path (latency and throughput families), so that they differ only in
// the order of their sums.
__device__ __forceinline__ bool eval_plane_at(const double* __restrict__ q, const PointCtx& pc, const ScanConst& sc,
                                              const Globals& g, bool need_prob, double& prob, Row& row) {
    const double2 v0 = *reinterpret_cast<const double2*>(q), v1 = *reinterpret_cast<const double2*>(q + 2),
                  v2 = *reinterpret_cast<const double2*>(q + 4);
    const double c0 = v0.x, c1 = v0.y, c2 = v1.x, n0 = v1.y, n1 = v2.x, n2 = v2.y;
    const float2 dr = *reinterpret_cast<const float2*>(q + 27);  // {d, radius}: floats in the reference (voxel_map.h:103,107)
    const double s = n0 * pc.pwx + n1 * pc.pwy + n2 * pc.pwz + (double)dr.x;
    const float dis = (float)fabs(s);
    const double ax = pc.pwx - c0, ay = pc.pwy - c1, az = pc.pwz - c2;
    const float dc = (float)(ax * ax + ay * ay + az * az);
    const float rd = sqrtf(__fsub_rn(dc, __fmul_rn(dis, dis)));  // float arithmetic as in the reference
    if (!((double)rd <= 3.0 * (double)dr.y)) return false;
    // J_nq Sigma_plane J_nq^T, J_nq = [(pw - c)^T, -n^T]
    const double J0 = ax, J1 = ay, J2 = az, J3 = -n0, J4 = -n1, J5 = -n2;
    const double* pv = q + 6;
    double sigma_pl = J0 * (pv[0] * J0 + 2.0 * (pv[1] * J1 + pv[2] * J2 + pv[3] * J3 + pv[4] * J4 + pv[5] * J5));
    sigma_pl += J1 * (pv[6] * J1 + 2.0 * (pv[7] * J2 + pv[8] * J3 + pv[9] * J4 + pv[10] * J5));
    sigma_pl += J2 * (pv[11] * J2 + 2.0 * (pv[12] * J3 + pv[13] * J4 + pv[14] * J5));
    sigma_pl += J3 * (pv[15] * J3 + 2.0 * (pv[16] * J4 + pv[17] * J5));
    sigma_pl += J4 * (pv[18] * J4 + 2.0 * (pv[19] * J5));
    sigma_pl += J5 * (pv[20] * J5);
    // q = R^T n ; h_theta = pi x q ; w = (R Re)^T n = Re^T q
    const double qx = sc.R[0] * n0 + sc.R[3] * n1 + sc.R[6] * n2;
    const double qy = sc.R[1] * n0 + sc.R[4] * n1 + sc.R[7] * n2;
    const double qz = sc.R[2] * n0 + sc.R[5] * n1 + sc.R[8] * n2;
    const double hx = pc.piy * qz - pc.piz * qy, hy = pc.piz * qx - pc.pix * qz, hz = pc.pix * qy - pc.piy * qx;
    const double wx = g.Re[0] * qx + g.Re[3] * qy + g.Re[6] * qz;
    const double wy = g.Re[1] * qx + g.Re[4] * qy + g.Re[7] * qz;
    const double wz = g.Re[2] * qx + g.Re[5] * qy + g.Re[8] * qz;
    const double uw = pc.pbx * wx + pc.pby * wy + pc.pbz * wz;
    const double ww = wx * wx + wy * wy + wz * wz;
    const double uw2 = uw * uw / pc.r2;  // (u.w)^2
    const double body = (double)g.rv * uw2 + pc.range2 * g.dv * (ww - uw2);
    const double state = quad_sym3(sc.Pth, hx, hy, hz) + quad_sym3(sc.Ppp, n0, n1, n2);
    const double sigma_l = sigma_pl + body + state;
    // gate 2: dis_to_plane < sigma_num * sqrt(sigma_l)   (voxel_map.cc:387), squared with an exact
    // fallback at the boundary so the decision equals the reference's comparison.
    const double lhs = (double)dis * (double)dis;
    const double rhs = g.sigma_num * g.sigma_num * sigma_l;
    bool pass;
    if (lhs < rhs * (1.0 - 1e-12)) pass = true;
    else if (lhs > rhs * (1.0 + 1e-12)) pass = false;
    else pass = (double)dis < g.sigma_num * sqrt(sigma_l);
    if (!pass) return false;
    if (need_prob) {
        const double this_prob = 1.0 / sqrt(sigma_l) * exp(-0.5 * (double)dis * (double)dis / sigma_l);
        if (!(this_prob > prob)) return true;  // is_success without replacing the candidate
        prob = this_prob;
    }
    row.h[0] = hx; row.h[1] = hy; row.h[2] = hz; row.h[3] = n0; row.h[4] = n1; row.h[5] = n2;
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
        const MapNode* cn = nodes + st_base[sp - 1] + c;
        const uint32_t cflags = rec_flags(cn);
        if (cflags & LK_NODE_IS_PLANE) {
            if (eval_plane_at(reinterpret_cast<const double*>(cn), pc, sc, g, true, prob, row)) ok = true;
        } else if (layer < g.max_layer && sp < 4) {
            const uint32_t cm = (cflags >> LK_NODE_CHILDMASK_SHIFT) & 0xffu;
            const int cb = rec_child_base(cn);
            if (cb >= 0 && cm) {
                st_base[sp] = cb;
                st_mask[sp] = cm;
                ++sp;
            }
        }
    }
    *probp = prob;
    *rowp = row;
    return ok;
}

// build_single_residual on one root record (in shared or global memory): the plane branch, or the descent.
__device__ __forceinline__ bool eval_record(const MapNode* __restrict__ nodes, const void* rec, const PointCtx& pc,
                                            const ScanConst& sc, const Globals& g, double& prob, Row& row) {
    const uint32_t flags = rec_flags(rec);
    if (flags & LK_NODE_IS_PLANE) return eval_plane_at(reinterpret_cast<const double*>(rec), pc, sc, g, false, prob, row);
    const uint32_t cmask = (flags >> LK_NODE_CHILDMASK_SHIFT) & 0xffu;
    const int cb = rec_child_base(rec);
    if (g.max_layer >= 1 && cb >= 0 && cmask) return visit_subtree(nodes, cb, cmask, &pc, &sc, &g, &prob, &row);
    return false;
}
__device__ __forceinline__ bool eval_record(const MapNode* __restrict__ nodes, const void* rec, const PointCtx& pc,
                                            const ScanConst& sc, const Globals& g, Row& row) {
    double prob = 0.0;
    return eval_record(nodes, rec, pc, sc, g, prob, row);
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
        ok = eval_record(a.nodes, a.nodes + root, pc, sc, g, prob, row);
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
