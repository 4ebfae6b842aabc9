// lk_mapio.cu — device map storage management and the blob import / export
// (lk_map_upload / lk_map_download of include/legkilo_b200.h).
#include <cstring>
#include <vector>

#include "lk_kernels.h"
#include "lk_mapdev.h"
#include "lk_octree.cuh"

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

__global__ void k_hash_insert_roots(HashSlot* slots, uint32_t mask, const lk_map_root* roots, uint32_t n, uint32_t* ovf) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lk_map_root r = roots[i];
    if (!hash_insert_dev(slots, mask, r.key[0], r.key[1], r.key[2], r.node)) atomicOr(ovf, 4u);
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

// every root key must resolve to its own node: a second root with the same key is unreachable (and means a corrupt blob)
__global__ void k_check_roots(const HashSlot* slots, uint32_t mask, const lk_map_root* roots, uint32_t n, uint32_t* ovf) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    lk_map_root r = roots[i];
    if (hash_find_dev(slots, mask, r.key[0], r.key[1], r.key[2]) != r.node) atomicOr(ovf, 8u);
}

// hot images of nodes [0, n) from their node records (after a blob upload)
__global__ void k_hot_from_nodes(const MapNode* nodes, HotRec* hot, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) hot_from_node(nodes[i], hot[i]);
}

__global__ void k_count_planes(const MapNode* nodes, const MapAux* aux, uint32_t n, unsigned long long* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (nodes[i].flags & LK_NODE_IS_PLANE) atomicAdd(out, 1ull);
    if (aux[i].pts_count > 0) atomicAdd(out + 1, (unsigned long long)aux[i].pts_count);
}

uint64_t next_pow2(uint64_t v) {
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

#define MI_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            cudaGetLastError();                                                                \
            err = std::string(#expr) + ": " + cudaGetErrorString(e__);                         \
            return e__ == cudaErrorMemoryAllocation ? LK_ERR_OUT_OF_MEMORY : LK_ERR_CUDA;      \
        }                                                                                      \
    } while (0)

}  // namespace

void MapDevHost::release() {
    void* ptrs[] = {slots, nodes, aux, hot, points, counters};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    slots = nullptr; nodes = nullptr; aux = nullptr; hot = nullptr; points = nullptr; counters = nullptr;
    hash_cap = node_cap = point_cap = 0;
    n_roots = n_nodes = 0;
    n_points = 0;
}

MapDev MapDevHost::dev() const {
    MapDev d;
    d.slots = slots;
    d.hash_mask = (uint32_t)(hash_cap - 1);
    d.nodes = nodes;
    d.aux = aux;
    d.hot = hot;
    d.points = points;
    d.node_cap = (uint32_t)node_cap;
    d.point_cap = point_cap;
    d.n_nodes = counters;
    d.n_roots = counters + 1;
    d.overflow = counters + 2;
    d.n_points = reinterpret_cast<unsigned long long*>(counters + 4);
    return d;
}

int MapDevHost::allocate(uint64_t roots, uint64_t nnodes, uint64_t npoints, cudaStream_t s, std::string& err) {
    uint64_t want_hash = next_pow2(std::max<uint64_t>(1024, 4 * (roots + reserve_roots)));  // load factor <= 0.25
    if (want_hash > (1ull << 31)) { err = "root table too large"; return LK_ERR_CAPACITY; }
    uint64_t want_nodes = std::max<uint64_t>(nnodes + reserve_nodes, 64);
    uint64_t want_points = std::max<uint64_t>(npoints + reserve_points, 64);
    if (want_nodes >= (1ull << 31)) { err = "node pool too large"; return LK_ERR_CAPACITY; }
    if (want_points >= (1ull << 32)) { err = "point pool too large"; return LK_ERR_CAPACITY; }
    if (!counters) MI_CUDA(cudaMalloc((void**)&counters, 64));
    if (want_hash != hash_cap) {
        if (slots) cudaFree(slots);
        slots = nullptr; hash_cap = 0;
        MI_CUDA(cudaMalloc((void**)&slots, want_hash * sizeof(HashSlot)));
        hash_cap = want_hash;
    }
    if (want_nodes > node_cap) {
        if (nodes) cudaFree(nodes);
        if (aux) cudaFree(aux);
        if (hot) cudaFree(hot);
        nodes = nullptr; aux = nullptr; hot = nullptr; node_cap = 0;
        MI_CUDA(cudaMalloc((void**)&nodes, want_nodes * sizeof(MapNode)));
        MI_CUDA(cudaMalloc((void**)&aux, want_nodes * sizeof(MapAux)));
        MI_CUDA(cudaMalloc((void**)&hot, want_nodes * sizeof(HotRec)));
        node_cap = want_nodes;
    }
    if (want_points > point_cap) {
        if (points) cudaFree(points);
        points = nullptr; point_cap = 0;
        MI_CUDA(cudaMalloc((void**)&points, want_points * sizeof(DevPoint)));
        point_cap = want_points;
    }
    k_hash_clear<<<(unsigned)((hash_cap + 255) / 256), 256, 0, s>>>(slots, hash_cap);
    MI_CUDA(cudaMemsetAsync(counters, 0, 64, s));
    MI_CUDA(cudaGetLastError());
    n_roots = n_nodes = 0;
    n_points = 0;
    return LK_OK;
}

int MapDevHost::ensure_headroom(uint64_t extra_roots, uint64_t extra_nodes, uint64_t extra_points, cudaStream_t s,
                                std::string& err) {
    if (!ready()) return allocate(extra_roots, extra_nodes, extra_points, s, err);
    // nodes
    if (n_nodes + extra_nodes > node_cap) {
        uint64_t want = std::max<uint64_t>(n_nodes + extra_nodes, node_cap + node_cap / 2);
        if (want >= (1ull << 31)) { err = "node pool too large"; return LK_ERR_CAPACITY; }
        MapNode* nn = nullptr;
        MapAux* na = nullptr;
        HotRec* nh = nullptr;
        MI_CUDA(cudaMalloc((void**)&nn, want * sizeof(MapNode)));
        MI_CUDA(cudaMalloc((void**)&na, want * sizeof(MapAux)));
        MI_CUDA(cudaMalloc((void**)&nh, want * sizeof(HotRec)));
        MI_CUDA(cudaMemcpyAsync(nn, nodes, (size_t)n_nodes * sizeof(MapNode), cudaMemcpyDeviceToDevice, s));
        MI_CUDA(cudaMemcpyAsync(na, aux, (size_t)n_nodes * sizeof(MapAux), cudaMemcpyDeviceToDevice, s));
        MI_CUDA(cudaMemcpyAsync(nh, hot, (size_t)n_nodes * sizeof(HotRec), cudaMemcpyDeviceToDevice, s));
        MI_CUDA(cudaStreamSynchronize(s));
        cudaFree(nodes); cudaFree(aux); cudaFree(hot);
        nodes = nn; aux = na; hot = nh; node_cap = want;
    }
    if (n_points + extra_points > point_cap) {
        uint64_t want = std::max<uint64_t>(n_points + extra_points, point_cap + point_cap / 2);
        if (want >= (1ull << 32)) { err = "point pool too large"; return LK_ERR_CAPACITY; }
        DevPoint* np = nullptr;
        MI_CUDA(cudaMalloc((void**)&np, want * sizeof(DevPoint)));
        MI_CUDA(cudaMemcpyAsync(np, points, (size_t)n_points * sizeof(DevPoint), cudaMemcpyDeviceToDevice, s));
        MI_CUDA(cudaStreamSynchronize(s));
        cudaFree(points);
        points = np; point_cap = want;
    }
    if (4 * (n_roots + extra_roots) > hash_cap) {
        // rehash: dump roots, rebuild a bigger table
        uint64_t want = next_pow2(4 * (n_roots + extra_roots) + 4 * reserve_roots);
        if (want > (1ull << 31)) { err = "root table too large"; return LK_ERR_CAPACITY; }
        lk_map_root* tmp = nullptr;
        MI_CUDA(cudaMalloc((void**)&tmp, std::max<size_t>(n_roots, 1) * sizeof(lk_map_root)));
        uint32_t* cnt = counters + 8;
        MI_CUDA(cudaMemsetAsync(cnt, 0, 4, s));
        k_hash_dump<<<(unsigned)((hash_cap + 255) / 256), 256, 0, s>>>(slots, hash_cap, tmp, cnt);
        HashSlot* ns = nullptr;
        MI_CUDA(cudaMalloc((void**)&ns, want * sizeof(HashSlot)));
        k_hash_clear<<<(unsigned)((want + 255) / 256), 256, 0, s>>>(ns, want);
        if (n_roots) k_hash_insert_roots<<<(n_roots + 255) / 256, 256, 0, s>>>(ns, (uint32_t)(want - 1), tmp, n_roots, counters + 2);
        MI_CUDA(cudaStreamSynchronize(s));
        cudaFree(slots); cudaFree(tmp);
        slots = ns; hash_cap = want;
    }
    return LK_OK;
}

int MapDevHost::sync_counters(cudaStream_t s, std::string& err) {
    uint32_t h[6] = {0, 0, 0, 0, 0, 0};
    MI_CUDA(cudaMemcpyAsync(h, counters, 24, cudaMemcpyDeviceToHost, s));
    MI_CUDA(cudaStreamSynchronize(s));
    n_nodes = h[0];
    n_roots = h[1];
    unsigned long long np;
    std::memcpy(&np, &h[4], 8);
    n_points = np;
    return LK_OK;
}

int MapDevHost::push_counters(cudaStream_t s, std::string& err) {
    uint32_t h[6] = {n_nodes, n_roots, 0, 0, 0, 0};
    unsigned long long np = n_points;
    std::memcpy(&h[4], &np, 8);
    MI_CUDA(cudaMemcpyAsync(counters, h, 24, cudaMemcpyHostToDevice, s));
    MI_CUDA(cudaStreamSynchronize(s));
    return LK_OK;
}

int map_upload_blob(MapDevHost& mh, const Globals& g, const void* blob, size_t bytes, cudaStream_t s, std::string& err) {
    if (bytes < sizeof(lk_map_blob_header)) { err = "blob shorter than its header"; return LK_ERR_BAD_BLOB; }
    lk_map_blob_header hd;
    std::memcpy(&hd, blob, sizeof(hd));
    if (hd.magic != LK_MAP_MAGIC || hd.version != 1) { err = "bad magic / version"; return LK_ERR_BAD_BLOB; }
    size_t need = sizeof(hd) + (size_t)hd.n_roots * sizeof(lk_map_root) + (size_t)hd.n_nodes * (sizeof(lk_map_node) + sizeof(lk_map_aux)) +
                  (size_t)hd.n_points * sizeof(lk_map_point);
    if (bytes < need) { err = "blob truncated"; return LK_ERR_BAD_BLOB; }
    const char* p = (const char*)blob + sizeof(hd);
    const lk_map_root* roots = (const lk_map_root*)p;
    p += (size_t)hd.n_roots * sizeof(lk_map_root);
    const lk_map_node* nodes = (const lk_map_node*)p;
    p += (size_t)hd.n_nodes * sizeof(lk_map_node);
    const lk_map_aux* aux = (const lk_map_aux*)p;
    p += (size_t)hd.n_nodes * sizeof(lk_map_aux);
    const lk_map_point* pts = (const lk_map_point*)p;
    for (uint32_t r = 0; r < hd.n_roots; ++r)
        if (roots[r].node < 0 || (uint32_t)roots[r].node >= hd.n_nodes) { err = "root node index out of range"; return LK_ERR_BAD_BLOB; }

    // Re-pack retained points into growable 80-byte tiles. A node may still take points when it is
    // not initialised, or is an update-enabled leaf (plane, or non-plane at max_layer).
    std::vector<lk_map_aux> aux2(aux, aux + hd.n_nodes);
    std::vector<DevPoint> dpts;
    const int P2 = g.max_points_num + 2;
    uint64_t slots_needed = 0;
    for (uint32_t i = 0; i < hd.n_nodes; ++i) {
        const lk_map_node& n = nodes[i];
        lk_map_aux& a = aux2[i];
        if ((uint64_t)a.pts_base + (uint64_t)std::max(a.pts_count, 0) > hd.n_points) { err = "node point range out of bounds"; return LK_ERR_BAD_BLOB; }
        int layer = (n.flags >> LK_NODE_LAYER_SHIFT) & 0xff;
        const uint32_t cmask = (n.flags >> LK_NODE_CHILDMASK_SHIFT) & 0xffu;
        // the residual kernels follow child_base / the child mask without further checks
        if (layer > g.max_layer || layer > 4) { err = "node layer exceeds max_layer"; return LK_ERR_BAD_BLOB; }
        if (n.child_base < -1 || (n.child_base >= 0 && (uint64_t)n.child_base + 8 > hd.n_nodes)) { err = "node child_base out of range"; return LK_ERR_BAD_BLOB; }
        if (cmask && n.child_base < 0) { err = "node has a child mask but no children"; return LK_ERR_BAD_BLOB; }
        bool init = n.flags & LK_NODE_INIT_OCTO, plane = n.flags & LK_NODE_IS_PLANE, upd = n.flags & LK_NODE_UPDATE_ENABLE;
        bool interior = init && !plane && layer < g.max_layer;
        bool can_grow = !init || (upd && !interior);
        int cnt = interior ? 0 : std::max(a.pts_count, 0);
        int cap = 0;
        if (cnt > 0) cap = ((std::max(cnt + 1, can_grow ? P2 : cnt) + 1) & ~1);
        a.pts_count = cnt;
        a.pts_cap = cap;
        const uint32_t src = a.pts_base;
        a.pts_base = (uint32_t)slots_needed;
        if (cap) {
            size_t at = dpts.size();
            dpts.resize(at + cap);
            std::memset(&dpts[at], 0, sizeof(DevPoint) * cap);
            for (int j = 0; j < cnt; ++j) {
                std::memcpy(dpts[at + j].pw, pts[src + j].pw, 24);
                std::memcpy(dpts[at + j].var, pts[src + j].var, 48);
            }
        }
        slots_needed += cap;
    }
    int rc = mh.allocate(hd.n_roots, hd.n_nodes, slots_needed, s, err);
    if (rc) return rc;
    lk_map_root* d_roots = nullptr;
    MI_CUDA(cudaMalloc((void**)&d_roots, std::max<size_t>(hd.n_roots, 1) * sizeof(lk_map_root)));
    cudaError_t e = cudaSuccess;
    if (hd.n_nodes) {
        e = cudaMemcpyAsync(mh.nodes, nodes, (size_t)hd.n_nodes * sizeof(MapNode), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(mh.aux, aux2.data(), (size_t)hd.n_nodes * sizeof(MapAux), cudaMemcpyHostToDevice, s);
    }
    if (e == cudaSuccess && !dpts.empty()) e = cudaMemcpyAsync(mh.points, dpts.data(), dpts.size() * sizeof(DevPoint), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && hd.n_roots) {
        e = cudaMemcpyAsync(d_roots, roots, (size_t)hd.n_roots * sizeof(lk_map_root), cudaMemcpyHostToDevice, s);
        k_hash_insert_roots<<<(hd.n_roots + 255) / 256, 256, 0, s>>>(mh.slots, (uint32_t)(mh.hash_cap - 1), d_roots, hd.n_roots, mh.counters + 2);
        k_check_roots<<<(hd.n_roots + 255) / 256, 256, 0, s>>>(mh.slots, (uint32_t)(mh.hash_cap - 1), d_roots, hd.n_roots, mh.counters + 2);
    }
    if (e == cudaSuccess && hd.n_nodes) k_hot_from_nodes<<<(hd.n_nodes + 255) / 256, 256, 0, s>>>(mh.nodes, mh.hot, hd.n_nodes);
    uint32_t ovf = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&ovf, mh.counters + 2, 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_roots);
    if (e != cudaSuccess) { cudaGetLastError(); err = cudaGetErrorString(e); return LK_ERR_CUDA; }
    if (ovf & 8u) { err = "duplicate root keys in the blob"; return LK_ERR_BAD_BLOB; }
    if (ovf) { err = "root table overflow"; return LK_ERR_CAPACITY; }
    mh.n_roots = hd.n_roots;
    mh.n_nodes = hd.n_nodes;
    mh.n_points = slots_needed;
    return mh.push_counters(s, err);
}

int map_download_blob(MapDevHost& mh, void* blob, size_t capacity, size_t* bytes_out, cudaStream_t s, std::string& err) {
    if (!mh.ready()) {
        lk_map_blob_header hd;
        std::memset(&hd, 0, sizeof(hd));
        hd.magic = LK_MAP_MAGIC; hd.version = 1;
        if (bytes_out) *bytes_out = sizeof(hd);
        if (blob) {
            if (capacity < sizeof(hd)) { err = "blob buffer too small"; return LK_ERR_CAPACITY; }
            std::memcpy(blob, &hd, sizeof(hd));
        }
        return LK_OK;
    }
    int rc = mh.sync_counters(s, err);
    if (rc) return rc;
    std::vector<lk_map_aux> aux(mh.n_nodes);
    if (mh.n_nodes) MI_CUDA(cudaMemcpyAsync(aux.data(), mh.aux, (size_t)mh.n_nodes * sizeof(MapAux), cudaMemcpyDeviceToHost, s));
    MI_CUDA(cudaStreamSynchronize(s));
    uint64_t live = 0;
    for (auto& a : aux) live += (uint64_t)std::max(a.pts_count, 0);
    size_t need = sizeof(lk_map_blob_header) + (size_t)mh.n_roots * sizeof(lk_map_root) +
                  (size_t)mh.n_nodes * (sizeof(lk_map_node) + sizeof(lk_map_aux)) + (size_t)live * sizeof(lk_map_point);
    if (bytes_out) *bytes_out = need;
    if (!blob) return LK_OK;
    if (capacity < need) { err = "blob buffer too small"; return LK_ERR_CAPACITY; }
    lk_map_blob_header hd;
    std::memset(&hd, 0, sizeof(hd));
    hd.magic = LK_MAP_MAGIC; hd.version = 1;
    hd.n_roots = mh.n_roots; hd.n_nodes = mh.n_nodes; hd.n_points = live;
    char* p = (char*)blob;
    std::memcpy(p, &hd, sizeof(hd));
    p += sizeof(hd);
    lk_map_root* d_roots = nullptr;
    MI_CUDA(cudaMalloc((void**)&d_roots, std::max<size_t>(mh.n_roots, 1) * sizeof(lk_map_root)));
    uint32_t* cnt = mh.counters + 8;
    cudaMemsetAsync(cnt, 0, 4, s);
    k_hash_dump<<<(unsigned)((mh.hash_cap + 255) / 256), 256, 0, s>>>(mh.slots, mh.hash_cap, d_roots, cnt);
    cudaError_t e = cudaMemcpyAsync(p, d_roots, (size_t)mh.n_roots * sizeof(lk_map_root), cudaMemcpyDeviceToHost, s);
    p += (size_t)mh.n_roots * sizeof(lk_map_root);
    if (e == cudaSuccess && mh.n_nodes) e = cudaMemcpyAsync(p, mh.nodes, (size_t)mh.n_nodes * sizeof(MapNode), cudaMemcpyDeviceToHost, s);
    p += (size_t)mh.n_nodes * sizeof(MapNode);
    lk_map_aux* out_aux = (lk_map_aux*)p;
    p += (size_t)mh.n_nodes * sizeof(MapAux);
    lk_map_point* out_pts = (lk_map_point*)p;
    std::vector<DevPoint> dpts((size_t)mh.n_points);
    if (e == cudaSuccess && mh.n_points) e = cudaMemcpyAsync(dpts.data(), mh.points, (size_t)mh.n_points * sizeof(DevPoint), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFree(d_roots);
    if (e != cudaSuccess) { cudaGetLastError(); err = cudaGetErrorString(e); return LK_ERR_CUDA; }
    uint64_t at = 0;
    for (uint32_t i = 0; i < mh.n_nodes; ++i) {
        lk_map_aux a = aux[i];
        int cnt_i = std::max(a.pts_count, 0);
        for (int j = 0; j < cnt_i; ++j) {
            std::memcpy(out_pts[at + j].pw, dpts[(size_t)a.pts_base + j].pw, 24);
            std::memcpy(out_pts[at + j].var, dpts[(size_t)a.pts_base + j].var, 48);
        }
        a.pts_base = (uint32_t)at;
        a.pts_count = cnt_i;
        out_aux[i] = a;
        at += cnt_i;
    }
    return LK_OK;
}

int map_clear_outside(MapDevHost& mh, const int lo[3], const int hi[3], uint64_t* removed, cudaStream_t s, std::string& err) {
    if (removed) *removed = 0;
    if (!mh.ready()) return LK_OK;
    int rc = mh.sync_counters(s, err);
    if (rc) return rc;
    if (mh.n_roots == 0) return LK_OK;
    lk_map_root* d_roots = nullptr;
    uint32_t* d_cnt = nullptr;
    MI_CUDA(cudaMalloc((void**)&d_roots, (size_t)mh.n_roots * sizeof(lk_map_root)));
    MI_CUDA(cudaMalloc((void**)&d_cnt, 4));
    MI_CUDA(cudaMemsetAsync(d_cnt, 0, 4, s));
    k_hash_dump<<<(unsigned)((mh.hash_cap + 255) / 256), 256, 0, s>>>(mh.slots, mh.hash_cap, d_roots, d_cnt);
    std::vector<lk_map_root> roots(mh.n_roots);
    uint32_t n = 0;
    cudaError_t e = cudaMemcpyAsync(&n, d_cnt, 4, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e == cudaSuccess && n) e = cudaMemcpy(roots.data(), d_roots, (size_t)std::min(n, mh.n_roots) * sizeof(lk_map_root), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { cudaFree(d_roots); cudaFree(d_cnt); cudaGetLastError(); err = cudaGetErrorString(e); return LK_ERR_CUDA; }
    n = std::min(n, mh.n_roots);
    uint32_t keep = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const lk_map_root& r = roots[i];
        const bool out = r.key[0] > hi[0] || r.key[0] < lo[0] || r.key[1] > hi[1] || r.key[1] < lo[1] || r.key[2] > hi[2] || r.key[2] < lo[2];
        if (!out) roots[keep++] = r;
    }
    if (removed) *removed = n - keep;
    if (keep != n) {
        k_hash_clear<<<(unsigned)((mh.hash_cap + 255) / 256), 256, 0, s>>>(mh.slots, mh.hash_cap);
        if (keep) {
            e = cudaMemcpyAsync(d_roots, roots.data(), (size_t)keep * sizeof(lk_map_root), cudaMemcpyHostToDevice, s);
            k_hash_insert_roots<<<(keep + 255) / 256, 256, 0, s>>>(mh.slots, (uint32_t)(mh.hash_cap - 1), d_roots, keep, mh.counters + 2);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        mh.n_roots = keep;
    }
    cudaFree(d_roots); cudaFree(d_cnt);
    if (e != cudaSuccess) { cudaGetLastError(); err = cudaGetErrorString(e); return LK_ERR_CUDA; }
    return keep != n ? mh.push_counters(s, err) : LK_OK;
}

int map_count_planes(MapDevHost& mh, uint64_t* planes, uint64_t* live_points, cudaStream_t s, std::string& err) {
    *planes = 0;
    *live_points = 0;
    if (!mh.ready()) return LK_OK;
    int rc = mh.sync_counters(s, err);
    if (rc) return rc;
    unsigned long long* out = reinterpret_cast<unsigned long long*>(mh.counters + 10);
    MI_CUDA(cudaMemsetAsync(out, 0, 16, s));
    if (mh.n_nodes) k_count_planes<<<(mh.n_nodes + 255) / 256, 256, 0, s>>>(mh.nodes, mh.aux, mh.n_nodes, out);
    unsigned long long h[2] = {0, 0};
    MI_CUDA(cudaMemcpyAsync(h, out, 16, cudaMemcpyDeviceToHost, s));
    MI_CUDA(cudaStreamSynchronize(s));
    *planes = h[0];
    *live_points = h[1];
    return LK_OK;
}

}  // namespace lk
