// lk_mapdev.h — host-side owner of the device map buffers (root table, node / aux / point pools,
// allocator counters) and the entry points of the map translation units.
#pragma once
#include <string>

#include "lk_device.cuh"

namespace lk {

struct DevPoint;
struct MapDev;

class MapDevHost {
   public:
    ~MapDevHost() { release(); }
    // (Re)create an EMPTY map sized for at least these counts plus the reserve; clears the table.
    int allocate(uint64_t roots, uint64_t nodes, uint64_t points, cudaStream_t s, std::string& err);
    // Make sure pools can take `extra_*` more items (grows by reallocation + copy). No-op if they fit.
    int ensure_headroom(uint64_t extra_roots, uint64_t extra_nodes, uint64_t extra_points, cudaStream_t s,
                        std::string& err);
    void release();
    MapDev dev() const;
    bool ready() const { return hash_cap != 0; }
    // pull the device allocator counters into the host mirrors
    int sync_counters(cudaStream_t s, std::string& err);
    int push_counters(cudaStream_t s, std::string& err);

    HashSlot* slots = nullptr;
    MapNode* nodes = nullptr;
    MapAux* aux = nullptr;
    HotRec* hot = nullptr;  // hot image of every node's plane (what the throughput kernel gathers)
    DevPoint* points = nullptr;
    uint32_t* counters = nullptr;  // [0] n_nodes [1] n_roots [2] overflow [4..5] n_points (u64)
    uint64_t hash_cap = 0, node_cap = 0, point_cap = 0;
    uint32_t n_roots = 0, n_nodes = 0;
    uint64_t n_points = 0;  // bump pointer (slots handed out), not the number of live points
    uint64_t reserve_roots = 0, reserve_nodes = 0, reserve_points = 0;
};

// lk_mapbuild.cu
int map_build_device(MapDevHost& mh, const Globals& g, const float* d_xyz_world, const float* d_xyz_body, uint32_t n,
                     const double* rot, const double* rot_cov, const double* pos_cov, cudaStream_t s, std::string& err);
// lk_insert.cu — UpdateVoxelMap for one bucket (scratch buffers owned by the caller). Returns the number of
// launches (0 = nothing to do). world != null: the bucket's re-projected cloud is written too (the caller then
// skips its own re-projection kernel). small_parity != null enables the two-launch path for buckets of up to
// 4 096 points; the caller zeroes counters[2..3] and *small_parity before the first bucket of a scan.
int map_insert_bucket(MapDevHost& mh, const Globals& g, const float4* pts, const ChunkDesc* chunks, uint32_t chunk_first,
                      uint32_t n_chunks, uint32_t pt_begin, uint32_t n_pts, const ScanConst* sc, const ScanStep* step,
                      void* ipts, int* iroot, int* pend, uint32_t* touched, uint32_t* counters, uint32_t* list,
                      cudaStream_t s, float4* world = nullptr, uint32_t* small_parity = nullptr);
size_t insert_point_bytes();
// lk_mapio.cu
int map_upload_blob(MapDevHost& mh, const Globals& g, const void* blob, size_t bytes, cudaStream_t s, std::string& err);
int map_download_blob(MapDevHost& mh, void* blob, size_t capacity, size_t* bytes_out, cudaStream_t s, std::string& err);
int map_count_planes(MapDevHost& mh, uint64_t* planes, uint64_t* live_points, cudaStream_t s, std::string& err);
// drop every root whose key is outside [lo, hi] on some axis (clearMemOutOfMap, voxel_map.cc:573-594); rebuilds the table
int map_clear_outside(MapDevHost& mh, const int lo[3], const int hi[3], uint64_t* removed, cudaStream_t s, std::string& err);

}  // namespace lk
