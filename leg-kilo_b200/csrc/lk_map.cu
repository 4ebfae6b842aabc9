// lk_map.cu — the voxel map in HBM: open-addressed root table (replaces
// std::unordered_map<Vector3i, VoxelOctoTree*>, voxel_map.h:186 — results never depend on the
// hash function, eigen_types.hpp:80-82, only on exact key match) plus node / point pools.
#include "lk_kernels.h"

namespace lk {

namespace {

__global__ void k_hash_clear(HashSlot* slots, uint64_t capacity) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < capacity) {
        HashSlot s;
        s.kx = 0; s.ky = 0; s.kz = 0; s.node = -1;
        slots[i] = s;
    }
}

// Claim a slot with CAS on the node word; key words are written by the claimant before the node
// index becomes visible to later kernels (the table is only read by OTHER kernels).
__device__ bool hash_insert(HashSlot* slots, uint32_t mask, int kx, int ky, int kz, int node) {
    uint32_t i = hash_key(kx, ky, kz) & mask;
    for (uint32_t probe = 0; probe <= mask; ++probe) {
        int* nodep = &slots[i].node;
        int old = atomicCAS(nodep, -1, -2);  // -2 = being written
        if (old == -1) {
            slots[i].kx = kx; slots[i].ky = ky; slots[i].kz = kz;
            __threadfence();
            atomicExch(nodep, node);
            return true;
        }
        i = (i + 1) & mask;
    }
    return false;
}

__global__ void k_hash_insert_roots(HashSlot* slots, uint32_t mask, const lk_map_root* roots, uint32_t n,
                                    uint32_t* fail) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lk_map_root r = roots[i];
    if (!hash_insert(slots, mask, r.key[0], r.key[1], r.key[2], r.node)) atomicExch(fail, 1u);
}

__global__ void k_hash_dump(const HashSlot* slots, uint64_t capacity, lk_map_root* roots, uint32_t* counter) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= capacity) return;
    HashSlot s = slots[i];
    if (s.node >= 0) {
        uint32_t o = atomicAdd(counter, 1u);
        lk_map_root r;
        r.key[0] = s.kx; r.key[1] = s.ky; r.key[2] = s.kz; r.node = s.node;
        roots[o] = r;
    }
}

}  // namespace

void launch_hash_clear(HashSlot* slots, uint64_t capacity, cudaStream_t s) {
    if (!capacity) return;
    k_hash_clear<<<(unsigned)((capacity + 255) / 256), 256, 0, s>>>(slots, capacity);
}

void launch_hash_insert_roots(HashSlot* slots, uint32_t mask, const lk_map_root* roots, uint32_t n_roots,
                              uint32_t* fail_flag, cudaStream_t s) {
    if (!n_roots) return;
    k_hash_insert_roots<<<(n_roots + 255) / 256, 256, 0, s>>>(slots, mask, roots, n_roots, fail_flag);
}

void launch_hash_dump_roots(const HashSlot* slots, uint64_t capacity, lk_map_root* roots, uint32_t* counter,
                            cudaStream_t s) {
    if (!capacity) return;
    k_hash_dump<<<(unsigned)((capacity + 255) / 256), 256, 0, s>>>(slots, capacity, roots, counter);
}

}  // namespace lk
