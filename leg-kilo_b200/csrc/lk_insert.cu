// lk_insert.cu — device-side VoxelMapManager::UpdateVoxelMap (voxel_map.cc:336-361) for one
// bucket: step 4 of KILO::predictUpdatePoint (KILO.cc:215-231).
//
// The reference inserts the bucket's points one after another; points only interact when they
// fall into the same ROOT voxel (every octree node belongs to exactly one root). So:
//   P1  per point: world point + covariance with the UPDATED state (KILO.cc:218-228), insert key
//       (voxelKeyFloor), find-or-create the root (CAS on the open-addressed table), count the
//       point on its root, first toucher registers the root;
//   P2  per touched root: reserve a slice of the pending list;
//   P3  per point: drop the point index into its root's slice;
//   P4  one warp per touched root: order the slice by point index (= the reference's insertion
//       order) and run UpdateOctoTree sequentially on it, with the warp-cooperative plane refit.
// Small buckets (the 2 ms buckets of streaming mode: a few hundred points) skip P2 / P3 and the sort: the
// warp of a touched root simply scans the bucket's root-per-point array in index order (P4S), and P1 also
// stores the re-projected world point (KILO.cc:216-224), so a bucket costs two launches instead of six.
#include "lk_kernels.h"
#include "lk_mapdev.h"
#include "lk_insert.cuh"

namespace lk {

namespace {

struct InsertArgs {
    MapDev md;
    Globals g;
    const float4* pts;
    const ChunkDesc* chunks;
    uint32_t chunk_first;
    const ScanConst* sc;
    const ScanStep* step;
    DevPoint* ipts;      // [bucket points] DevPoint per point
    int* iroot;          // [bucket points] root node per point (-1 = dropped)
    uint32_t pt_base;    // absolute index of the first point covered by this launch's scratch
    int* pend;           // [node_cap * 3] count | offset | fill
    uint32_t* touched;   // [bucket points]
    uint32_t* counters;  // [0] n_touched  [1] list bump
    uint32_t* list;      // [2 * bucket points]  (second half = sort scratch)
    uint32_t n_pts;
    float4* world;       // non-null: P1 also writes cloud_down_world (x, y, z, intensity 0 | 255)
    uint32_t cslot;      // which of the two n_touched counters this bucket uses (small-bucket path)
};

// P1 — also the re-projection's covariance half (KILO.cc:225-228).
__global__ void __launch_bounds__(256) k_insert_p1(const __grid_constant__ InsertArgs a) {
    __shared__ ScanConst s_sc;
    const int tid = threadIdx.x;
    const ChunkDesc cd = a.chunks[a.chunk_first + blockIdx.x];
    if (tid < (int)(sizeof(ScanConst) / sizeof(double)))
        reinterpret_cast<double*>(&s_sc)[tid] = reinterpret_cast<const double*>(a.sc + cd.scan)[tid];
    __syncthreads();
    const Globals& g = a.g;
    MapDev md = a.md;
    for (uint32_t i = tid; i < cd.count; i += blockDim.x) {
        const float4 pt = __ldg(a.pts + cd.start + i);
        const double bx = pt.x, by = pt.y, bz0 = pt.z;
        const double pix = g.Re[0] * bx + g.Re[1] * by + g.Re[2] * bz0 + g.te[0];
        const double piy = g.Re[3] * bx + g.Re[4] * by + g.Re[5] * bz0 + g.te[1];
        const double piz = g.Re[6] * bx + g.Re[7] * by + g.Re[8] * bz0 + g.te[2];
        DevPoint p;
        make_insert_point(pix, piy, piz, bx, by, (bz0 == 0.0) ? 0.0001 : bz0, s_sc, g, p);  // calcBodyCov saw pb.z == 0 -> 1e-4
        const uint32_t li = cd.start + i - a.pt_base;
        a.ipts[li] = p;
        if (a.world) {
            float4 o;
            o.x = (float)p.pw[0]; o.y = (float)p.pw[1]; o.z = (float)p.pw[2];
            o.w = a.step[cd.scan].updated ? 255.0f : 0.0f;
            a.world[cd.start + i] = o;
        }
        a.iroot[li] = insert_register_point(md, g, p, a.pend, a.touched, &a.counters[a.cslot]);
    }
}

__global__ void k_insert_p2(const __grid_constant__ InsertArgs a) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.counters[0]) return;
    const uint32_t root = a.touched[t];
    const int cnt = a.pend[root * 3];
    a.pend[root * 3 + 1] = (int)atomicAdd(&a.counters[1], (uint32_t)cnt);
    a.pend[root * 3 + 2] = 0;
}

__global__ void k_insert_p3(const __grid_constant__ InsertArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_pts) return;
    const int root = a.iroot[i];
    if (root < 0) return;
    const int slot = atomicAdd(&a.pend[root * 3 + 2], 1);
    a.list[a.pend[root * 3 + 1] + slot] = i;
}

__global__ void __launch_bounds__(128) k_insert_p4(const __grid_constant__ InsertArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpTile* tiles = reinterpret_cast<WarpTile*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpTile* wt = tiles + warp;
    if (lane == 0) {
        mbar_init(&wt->bar, 1);
        wt->phase = 0;
        mbar_init_fence();
    }
    __syncwarp();
    const uint32_t t = blockIdx.x * (blockDim.x >> 5) + warp;
    if (t >= a.counters[0]) return;
    const uint32_t root = a.touched[t];
    const int cnt = a.pend[root * 3];
    const int off = a.pend[root * 3 + 1];
    uint32_t* src = a.list + off;
    uint32_t* dst = a.list + a.n_pts + off;
    // rank sort by point index = the order UpdateVoxelMap walks input_points
    for (int j = lane; j < cnt; j += 32) {
        const uint32_t v = src[j];
        int rank = 0;
        for (int k = 0; k < cnt; ++k) rank += (src[k] < v) ? 1 : 0;
        dst[rank] = v;
    }
    __syncwarp();
    MapDev md = a.md;
    for (int j = 0; j < cnt; ++j) {
        const DevPoint p = a.ipts[dst[j]];
        warp_update_octo_tree(md, a.g, wt, root, p, lane);
    }
    __syncwarp();
    if (lane == 0) a.pend[root * 3] = 0;
}

// P4S — small buckets: one warp per touched root walks the bucket's root-per-point array in index order
// (= the order UpdateVoxelMap walks input_points) and inserts its own points; no slices, no sort.
__global__ void __launch_bounds__(128) k_insert_p4_scan(const __grid_constant__ InsertArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpTile* tiles = reinterpret_cast<WarpTile*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpTile* wt = tiles + warp;
    if (lane == 0) {
        mbar_init(&wt->bar, 1);
        wt->phase = 0;
        mbar_init_fence();
    }
    __syncwarp();
    if (blockIdx.x == 0 && threadIdx.x == 0) a.counters[a.cslot ^ 1u] = 0;  // the next bucket's counter
    const uint32_t t = blockIdx.x * (blockDim.x >> 5) + warp;
    if (t >= a.counters[a.cslot]) return;
    const uint32_t root = a.touched[t];
    MapDev md = a.md;
    warp_insert_root_scan(md, a.g, wt, root, a.iroot, a.ipts, a.n_pts, a.pend, lane);
}

}  // namespace

// Scratch owned by the caller (lk_api): sized for the largest bucket.
int map_insert_bucket(MapDevHost& mh, const Globals& g, const float4* pts, const ChunkDesc* chunks, uint32_t chunk_first,
                      uint32_t n_chunks, uint32_t pt_begin, uint32_t n_pts, const ScanConst* sc, const ScanStep* step,
                      void* ipts, int* iroot, int* pend, uint32_t* touched, uint32_t* counters, uint32_t* list,
                      cudaStream_t s, float4* world, uint32_t* small_parity) {
    if (!n_pts || !n_chunks) return LK_OK;
    InsertArgs a;
    a.md = mh.dev();
    a.g = g;
    a.pts = pts;
    a.chunks = chunks;
    a.chunk_first = chunk_first;
    a.sc = sc;
    a.step = step;
    a.ipts = reinterpret_cast<DevPoint*>(ipts);
    a.iroot = iroot;
    a.pt_base = pt_begin;
    a.pend = pend;
    a.touched = touched;
    a.counters = counters;
    a.list = list;
    a.n_pts = n_pts;
    a.world = world;
    a.cslot = 0;
    const int wpb = 4;
    if (small_parity && n_pts <= 4096u) {
        // counters[2 + parity] is this bucket's n_touched; P4S zeroes the other one for the next bucket (the caller
        // zeroed both before the first bucket of the scan)
        a.counters = counters + 2;
        a.cslot = *small_parity;
        *small_parity ^= 1u;
        k_insert_p1<<<n_chunks, 256, 0, s>>>(a);
        k_insert_p4_scan<<<(n_pts + wpb - 1) / wpb, wpb * 32, wpb * sizeof(WarpTile), s>>>(a);
        return 2;
    }
    cudaMemsetAsync(counters, 0, 8, s);
    k_insert_p1<<<n_chunks, 256, 0, s>>>(a);
    k_insert_p2<<<(n_pts + 255) / 256, 256, 0, s>>>(a);
    k_insert_p3<<<(n_pts + 255) / 256, 256, 0, s>>>(a);
    k_insert_p4<<<(n_pts + wpb - 1) / wpb, wpb * 32, wpb * sizeof(WarpTile), s>>>(a);
    return 5;
}

size_t insert_point_bytes() { return sizeof(DevPoint); }

}  // namespace lk
