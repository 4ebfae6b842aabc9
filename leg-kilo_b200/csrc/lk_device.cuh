// lk_device.cuh — device-side types and small dense math shared by the kernels of the
// B200-native Leg-KILO LiDAR update path. sm_100a only; fp64 throughout with the reference's
// float temporaries reproduced where they decide something.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/legkilo_b200.h"

namespace lk {

// ---- map in HBM --------------------------------------------------------------------------
// Open-addressed table of root voxels: 16-byte slots, linear probing, empty <=> node < 0.
struct __align__(16) HashSlot {
    int kx, ky, kz;
    int node;
};

__host__ __device__ __forceinline__ uint32_t hash_key(int x, int y, int z) {
    uint32_t h = (uint32_t)x * 73856093u ^ (uint32_t)y * 19349669u ^ (uint32_t)z * 83492791u;
    h ^= h >> 15;
    h *= 0x2c1b3c6du;
    h ^= h >> 12;
    h *= 0x297a2d39u;
    h ^= h >> 15;
    return h;
}

typedef lk_map_node MapNode;   // 256 B, first 232 B are what the residual kernel reads
typedef lk_map_aux MapAux;     // 64 B
typedef lk_map_point MapPoint; // 72 B

struct MapView {  // what the residual path reads of the map
    const HashSlot* slots;
    uint32_t hash_mask;
    const MapNode* nodes;
};

// ---- per-call constants (kernel parameter space) ------------------------------------------
struct Globals {
    double Re[9];      // extrinsic rotation  (KILO::ext_rot_)
    double te[3];      // extrinsic translation
    double voxel;      // max_voxel_size_ as double (KILO.cc:145)
    double inv_voxel;  // exact reciprocal when voxel is a power of two, else unused
    double sigma_num;
    double ratio;      // lidar_point_meas_ratio
    double dv;         // sin(DEG2RAD(beam_err))^2 with PCL's DEG2RAD constant (voxel_map.cc:27)
    float voxel_f;     // (float)max_voxel_size_ (voxel_map.cc:289,337)
    float rv;          // range_inc*range_inc in float (voxel_map.cc:25)
    float planer_threshold;  // (float)min_eigen_value
    int voxel_pow2;
    int max_layer;
    int max_points_num;
    int layer_init_num[5];
};

// What every residual thread needs of one scan's filter at the current linearisation point.
struct ScanConst {
    double R[9];    // rot
    double p[3];    // pos
    double Pth[6];  // sym(P[0:3,0:3]) upper: xx xy xz yy yz zz
    double Ppp[6];  // sym(P[3:6,3:6]) upper
};

// 32 doubles per chunk partial: A upper (21) | b (6) | sumR | count | pad
constexpr int NACC = 29;
constexpr int ACC_B = 21, ACC_SUMR = 27, ACC_CNT = 28;
constexpr int PARTIAL_STRIDE = 32;

struct ChunkDesc {
    uint32_t scan;   // scan index in the batch
    uint32_t start;  // first point (absolute)
    uint32_t count;  // points in this chunk
    uint32_t pad;
};

// Per-scan bookkeeping of the current bucket step.
struct ScanStep {
    uint32_t chunk_begin, chunk_end;  // chunks of this scan's current bucket
    uint32_t pt_begin, pt_end;        // point range of the current bucket
    double t_bucket;
    uint32_t active;  // 0 => scan has no bucket in this step
    uint32_t updated; // any iteration of this bucket produced an update
    uint32_t n_eff_last;  // residual count of the most recent iteration
    uint32_t pad;
};

// ---- tiny dense helpers ----------------------------------------------------------------------
__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

// math_utils.hpp:55-68 — Exp(v1,v2,v3) with the reference's 1e-5 identity threshold.
__device__ __forceinline__ void so3_exp3(double v1, double v2, double v3, double* E) {
    double norm = sqrt(v1 * v1 + v2 * v2 + v3 * v3);
    E[0] = 1; E[1] = 0; E[2] = 0; E[3] = 0; E[4] = 1; E[5] = 0; E[6] = 0; E[7] = 0; E[8] = 1;
    if (norm > 0.00001) {
        double kx = v1 / norm, ky = v2 / norm, kz = v3 / norm;
        double s, c;
        sincos(norm, &s, &c);
        double c1 = 1.0 - c;
        double K[9] = {0, -kz, ky, kz, 0, -kx, -ky, kx, 0};
        double K2[9];
        mat3_mul(K, K, K2);
#pragma unroll
        for (int i = 0; i < 9; ++i) E[i] += s * K[i] + c1 * K2[i];
    }
}

// math_utils.hpp:20-32 — Exp(vec) with the 1e-7 threshold (used by getFx).
__device__ __forceinline__ void so3_exp_vec(double v1, double v2, double v3, double* E) {
    double norm = sqrt(v1 * v1 + v2 * v2 + v3 * v3);
    E[0] = 1; E[1] = 0; E[2] = 0; E[3] = 0; E[4] = 1; E[5] = 0; E[6] = 0; E[7] = 0; E[8] = 1;
    if (norm > 0.0000001) {
        double kx = v1 / norm, ky = v2 / norm, kz = v3 / norm;
        double s, c;
        sincos(norm, &s, &c);
        double c1 = 1.0 - c;
        double K[9] = {0, -kz, ky, kz, 0, -kx, -ky, kx, 0};
        double K2[9];
        mat3_mul(K, K, K2);
#pragma unroll
        for (int i = 0; i < 9; ++i) E[i] += s * K[i] + c1 * K2[i];
    }
}

// State::operator+= (eskf.cc:18-29) on the 36-double lk_state layout.
__device__ __forceinline__ void state_boxplus(double* x36, const double* d30) {
    double E[9], Rn[9];
    so3_exp3(d30[0], d30[1], d30[2], E);
    mat3_mul(x36, E, Rn);
#pragma unroll
    for (int i = 0; i < 9; ++i) x36[i] = Rn[i];
#pragma unroll
    for (int i = 0; i < 27; ++i) x36[9 + i] += d30[3 + i];
}

// In-place LU with partial pivoting of an n x n system with m right-hand sides (n <= 18).
// M is n x n row-major (stride n), B is n x m row-major. Single thread.
template <int MAXN>
__device__ inline bool lu_solve_small(double* M, double* B, int n, int m) {
    for (int k = 0; k < n; ++k) {
        int piv = k;
        double best = fabs(M[k * n + k]);
        for (int i = k + 1; i < n; ++i) {
            double v = fabs(M[i * n + k]);
            if (v > best) { best = v; piv = i; }
        }
        if (best == 0.0) return false;
        if (piv != k) {
            for (int j = 0; j < n; ++j) { double t = M[k * n + j]; M[k * n + j] = M[piv * n + j]; M[piv * n + j] = t; }
            for (int j = 0; j < m; ++j) { double t = B[k * m + j]; B[k * m + j] = B[piv * m + j]; B[piv * m + j] = t; }
        }
        double inv = 1.0 / M[k * n + k];
        for (int i = k + 1; i < n; ++i) {
            double l = M[i * n + k] * inv;
            if (l != 0.0) {
                for (int j = k + 1; j < n; ++j) M[i * n + j] -= l * M[k * n + j];
                for (int j = 0; j < m; ++j) B[i * m + j] -= l * B[k * m + j];
            }
        }
    }
    for (int i = n - 1; i >= 0; --i) {
        double inv = 1.0 / M[i * n + i];
        for (int j = 0; j < m; ++j) {
            double s = B[i * m + j];
            for (int k = i + 1; k < n; ++k) s -= M[i * n + k] * B[k * m + j];
            B[i * m + j] = s * inv;
        }
    }
    return true;
}

}  // namespace lk
