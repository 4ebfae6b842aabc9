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

namespace lk {

namespace {

constexpr int BLOCK = 256;
constexpr int WARPS = BLOCK / 32;

struct PlaneRec {
    double c[3], n[3], pv[21];
    float d, radius;
    uint32_t flags;
    int child_base;
};

// 15 x 128-bit read-only loads cover the 232 bytes the path needs (lk_map_node).
__device__ __forceinline__ void load_plane(const MapNode* __restrict__ nd, PlaneRec& r) {
    const double2* q = reinterpret_cast<const double2*>(nd);
    double2 v[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) v[i] = __ldg(q + i);
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

// build_single_residual over one root's octree (voxel_map.cc:363-427): the root's own plane, or
// every initialised plane among ALL children of non-plane nodes down to max_layer.
__device__ __forceinline__ bool visit_tree(const MapNode* __restrict__ nodes, int root, const PointCtx& pc,
                                           const ScanConst& sc, const Globals& g, double& prob, Row& row) {
    PlaneRec r;
    load_plane(nodes + root, r);
    if (r.flags & LK_NODE_IS_PLANE) return eval_plane(r, pc, sc, g, false, prob, row);
    uint32_t cmask = (r.flags >> LK_NODE_CHILDMASK_SHIFT) & 0xffu;
    if (g.max_layer < 1 || r.child_base < 0 || cmask == 0) return false;
    // rare path: iterative DFS, child order 0..7 as the reference's loop
    bool ok = false;
    int st_base[4];
    uint32_t st_mask[4];
    int sp = 0;
    st_base[0] = r.child_base;
    st_mask[0] = cmask;
    sp = 1;
    while (sp > 0) {
        uint32_t m = st_mask[sp - 1];
        if (m == 0) { --sp; continue; }
        int c = __ffs(m) - 1;
        st_mask[sp - 1] = m & (m - 1);
        int layer = sp;  // children of a layer-(sp-1) node
        PlaneRec cr;
        load_plane(nodes + st_base[sp - 1] + c, cr);
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
    return ok;
}

// One point through rows a3-a7. Returns true when a residual row was produced.
__device__ __forceinline__ bool point_row(float4 pt, const ScanConst& sc, const ResidualArgs& a, Row& row,
                                          int* key_out) {
    const Globals& g = a.g;
    PointCtx pc;
    double bx = (double)pt.x, by = (double)pt.y, bz = (double)pt.z;
    pc.pix = g.Re[0] * bx + g.Re[1] * by + g.Re[2] * bz + g.te[0];
    pc.piy = g.Re[3] * bx + g.Re[4] * by + g.Re[5] * bz + g.te[1];
    pc.piz = g.Re[6] * bx + g.Re[7] * by + g.Re[8] * bz + g.te[2];
    pc.pwx = sc.R[0] * pc.pix + sc.R[1] * pc.piy + sc.R[2] * pc.piz + sc.p[0];
    pc.pwy = sc.R[3] * pc.pix + sc.R[4] * pc.piy + sc.R[5] * pc.piz + sc.p[1];
    pc.pwz = sc.R[6] * pc.pix + sc.R[7] * pc.piy + sc.R[8] * pc.piz + sc.p[2];
    // calcBodyCov mutates pb.z AFTER pi / pw were formed (voxel_map.cc:23, KILO.cc:134)
    pc.pbx = bx; pc.pby = by; pc.pbz = (bz == 0.0) ? 0.0001 : bz;
    pc.r2 = pc.pbx * pc.pbx + pc.pby * pc.pby + pc.pbz * pc.pbz;
    float range = (float)sqrt(pc.r2);
    pc.range2 = (double)range * (double)range;

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
    int kx = (int)lx, ky = (int)ly, kz = (int)lz;
    if (key_out) { key_out[0] = kx; key_out[1] = ky; key_out[2] = kz; }

    int root = map_find(a.slots, a.hash_mask, kx, ky, kz);
    if (root < 0) return false;
    double prob = 0.0;
    bool ok = visit_tree(a.nodes, root, pc, sc, g, prob, row);
    if (!ok) {
        // neighbour fallback (KILO.cc:156-178): loc in VOXEL units against a centre in METRES
        double q = (double)(g.voxel_f / 4.0f);
        double cx = (0.5 + kx) * (double)g.voxel_f, cy = (0.5 + ky) * (double)g.voxel_f, cz = (0.5 + kz) * (double)g.voxel_f;
        int nx = kx, ny = ky, nz = kz;
        if ((double)lx > cx + q) nx++; else if ((double)lx < cx - q) nx--;
        if ((double)ly > cy + q) ny++; else if ((double)ly < cy - q) ny--;
        if ((double)lz > cz + q) nz++; else if ((double)lz < cz - q) nz--;
        if (nx != kx || ny != ky || nz != kz) {
            int near = map_find(a.slots, a.hash_mask, nx, ny, nz);
            if (near >= 0) ok = visit_tree(a.nodes, near, pc, sc, g, prob, row);
        }
    }
    return ok;
}

// eskf.cc:91-113 in information form + State::operator+= ; run by the last block of a scan.
__device__ void scan_solve(const ResidualArgs& a, uint32_t scan, double* sm /* >= 512 doubles */) {
    const int tid = threadIdx.x;
    ScanStep st = a.step[scan];
    double* sAcc = sm;        // NACC
    double* sY = sm + 32;     // y[6] | W[36] row-major 6x6 | flag
    double* sDelta = sm + 80; // 30
    double* sProw = sm + 112; // 6 x 30 old rows of P
    double* sKH = sm + 292;   // 30 x 6
    double* Pg = a.P + (size_t)scan * 900;
    if (tid < NACC) {
        double acc = 0.0;
        for (uint32_t c = st.chunk_begin; c < st.chunk_end; ++c) acc += __ldcg(a.partial + (size_t)c * PARTIAL_STRIDE + tid);
        sAcc[tid] = acc;
    }
    __syncthreads();
    const double cnt = sAcc[ACC_CNT];
    if (cnt > 0.5) {
        if (tid == 0) {
            double A[36], rhs[42], M[36];
            double scale = 1.0;
            if (cnt < 1.5) scale = sAcc[ACC_SUMR] / (sAcc[ACC_SUMR] + 0.0001);  // N == 1 adds 1e-4 to S (eskf.cc:100)
            int q = 0;
            for (int i = 0; i < 6; ++i)
                for (int j = i; j < 6; ++j) {
                    double v = sAcc[q++] * scale;
                    A[i * 6 + j] = v;
                    A[j * 6 + i] = v;
                }
            for (int i = 0; i < 6; ++i) {
                for (int j = 0; j < 6; ++j) {
                    double s = (i == j) ? 1.0 : 0.0;
                    for (int k = 0; k < 6; ++k) s += A[i * 6 + k] * Pg[k * 30 + j];
                    M[i * 6 + j] = s;
                    rhs[i * 7 + 1 + j] = A[i * 6 + j];
                }
                rhs[i * 7] = sAcc[21 + i] * scale;
            }
            bool okl = lu_solve_small<6>(M, rhs, 6, 7);
            for (int i = 0; i < 6; ++i) {
                sY[i] = okl ? rhs[i * 7] : 0.0;
                for (int j = 0; j < 6; ++j) sY[6 + i * 6 + j] = okl ? rhs[i * 7 + 1 + j] : 0.0;
            }
        }
        __syncthreads();
        if (tid < 30) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) s += Pg[tid * 30 + k] * sY[k];
            sDelta[tid] = s;
        }
        __syncthreads();
        if (tid == 0) {
            double* xs = a.x + (size_t)scan * 36;
            double xl[36], dl[30];
            for (int i = 0; i < 36; ++i) xl[i] = xs[i];
            for (int i = 0; i < 30; ++i) dl[i] = sDelta[i];
            state_boxplus(xl, dl);
            for (int i = 0; i < 36; ++i) xs[i] = xl[i];
            ScanConst* sc = a.sc + scan;
            for (int i = 0; i < 9; ++i) sc->R[i] = xl[i];
            for (int i = 0; i < 3; ++i) sc->p[i] = xl[9 + i];
            a.clk[scan].last_update_time = st.t_bucket;  // KILO.cc:212
        }
        if (a.last_iter) {
            // P <- P - (P6 W) P[0:6,:]
            if (tid < 180) {
                sProw[tid] = Pg[tid];  // rows 0..5 are contiguous
                int i = tid / 6, j = tid % 6;
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) s += Pg[i * 30 + k] * sY[6 + k * 6 + j];
                sKH[tid] = s;
            }
            __syncthreads();
            for (int e = tid; e < 900; e += BLOCK) {
                int i = e / 30, j = e % 30;
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < 6; ++k) s += sKH[i * 6 + k] * sProw[k * 30 + j];
                Pg[e] -= s;
            }
        }
    }
    __syncthreads();
    if (a.last_iter && tid == 0) {
        ScanConst* sc = a.sc + scan;
        // refresh the covariance blocks the re-projection / map insertion use (KILO.cc:225-228)
        const int ut[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
        for (int q = 0; q < 6; ++q) {
            int i = ut[q][0], j = ut[q][1];
            sc->Pth[q] = 0.5 * (Pg[i * 30 + j] + Pg[j * 30 + i]);
            sc->Ppp[q] = 0.5 * (Pg[(3 + i) * 30 + 3 + j] + Pg[(3 + j) * 30 + 3 + i]);
        }
    }
    if (tid == 0) {
        ScanStep* sp = a.step + scan;
        uint32_t n = (uint32_t)(cnt + 0.5);
        sp->n_eff_last = n;
        if (n > 0) sp->updated = 1;
        if (a.last_iter) a.n_eff[scan] += n;
        a.ticket[scan] = 0;
    }
}

template <bool DEBUG>
__global__ void __launch_bounds__(BLOCK) k_residual(const __grid_constant__ ResidualArgs a) {
    __shared__ ScanConst s_sc;
    __shared__ double s_red[512];
    __shared__ uint32_t s_last;
    const int tid = threadIdx.x;
    const ChunkDesc cd = a.chunks[a.chunk_first + blockIdx.x];
    if (tid < (int)(sizeof(ScanConst) / sizeof(double)))
        reinterpret_cast<double*>(&s_sc)[tid] = reinterpret_cast<const double*>(a.sc + cd.scan)[tid];
    __syncthreads();

    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0;

    const float4* __restrict__ pts = a.pts + cd.start;
    for (uint32_t i = tid; i < cd.count; i += BLOCK) {
        float4 pt = __ldg(pts + i);
        Row row;
        int key[3];
        bool ok = point_row(pt, s_sc, a, row, DEBUG ? key : nullptr);
        if (DEBUG) {
            size_t gi = (size_t)cd.start + i;
            a.dbg_ok[gi] = ok ? 1 : 0;
            for (int k = 0; k < 3; ++k) a.dbg_key[gi * 3 + k] = key[k];
            for (int k = 0; k < 6; ++k) a.dbg_h[gi * 6 + k] = ok ? row.h[k] : 0.0;
            a.dbg_z[gi] = ok ? row.z : 0.0;
            a.dbg_R[gi] = ok ? row.R : 0.0;
        } else if (ok) {
            double w = 1.0 / row.R;
            int q = 0;
#pragma unroll
            for (int r = 0; r < 6; ++r) {
                double hw = row.h[r] * w;
#pragma unroll
                for (int c = r; c < 6; ++c) acc[q++] += hw * row.h[c];
                acc[21 + r] += hw * row.z;
            }
            acc[ACC_SUMR] += row.R;
            acc[ACC_CNT] += 1.0;
        }
    }
    if (DEBUG) return;

    // block reduction: shuffle tree inside the warp, one row per warp in smem, fixed-order sum
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        double v = acc[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp * 32 + i] = v;
    }
    __syncthreads();
    if (tid < NACC) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) v += s_red[w * 32 + tid];
        a.partial[(size_t)(a.chunk_first + blockIdx.x) * PARTIAL_STRIDE + tid] = v;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const ScanStep* sp = a.step + cd.scan;
        uint32_t n_chunks = sp->chunk_end - sp->chunk_begin;
        uint32_t t = atomicAdd(a.ticket + cd.scan, 1u);
        s_last = (t == n_chunks - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        scan_solve(a, cd.scan, s_red);
    }
}

}  // namespace

void launch_residual(const ResidualArgs& a, uint32_t n_chunks, bool debug, int /*gather_mode*/, cudaStream_t s) {
    if (n_chunks == 0) return;
    if (debug)
        k_residual<true><<<n_chunks, BLOCK, 0, s>>>(a);
    else
        k_residual<false><<<n_chunks, BLOCK, 0, s>>>(a);
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
