// lk_llsync.cuh — grid-wide all-reduce of one 32-double row per block WITHOUT a barrier: rows travel
// through global memory in a flagged ("low-latency") format — every double is split into two 8-byte
// words {low 32 bits | tag << 32, high 32 bits | tag << 32} stored with ONE 16-byte store. Aligned 8-byte
// stores are single-copy atomic, so a reader that finds the expected tag in both words holds the whole
// value: data and "ready" flag arrive together, there is no release fence, no counter and no second
// round trip to read the rows after a barrier. The tag is a per-handle epoch that grows with every
// (launch, iteration), so stale rows of earlier iterations never match; two buffers alternate by
// iteration parity (a block writes iteration i+2 only after it has seen every row of iteration i+1,
// i.e. after every block has finished reading iteration i).
//
// Summation order (shared with the multi-kernel path, lk_solve.cuh: block_sum_partials):
//   total = sum over groups g ascending of ( sum over the rows of group g ascending ),
//   group g = chunks [g*LK_GROUP, (g+1)*LK_GROUP) of the bucket — a function of the bucket alone.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lk {

constexpr int LK_GROUP = 8;          // chunk rows per group (level-1 fan-in)
constexpr int LL_ROW = 32;           // slots (doubles) per row
constexpr int LL_MAX_CHUNKS = 160;   // >= SM count: the fused kernel runs one chunk per block
constexpr int LL_MAX_GROUPS = (LL_MAX_CHUNKS + LK_GROUP - 1) / LK_GROUP;

struct LLView {
    ulonglong2* chunk_rows;  // [2][LL_MAX_CHUNKS][LL_ROW]
    ulonglong2* group_rows;  // [2][LL_MAX_GROUPS][LL_ROW]
    uint32_t* stall;         // [8] watchdog record: [0] != 0 once a poll gave up | block | tag | first row | rows | lane
};
constexpr size_t LL_ROWS_BYTES = (size_t)2 * (LL_MAX_CHUNKS + LL_MAX_GROUPS) * LL_ROW * sizeof(ulonglong2);
constexpr size_t LL_BYTES = LL_ROWS_BYTES + 64;
// A poll that sees nothing for this many rounds (seconds) gives up, records who waited for what and lets the kernel run
// to its end with garbage sums; the host then reports LK_ERR_CUDA instead of hanging. It means the blocks of the grid were
// not all resident (another process holds SMs: see INTEGRATION.md "Sharing a device") — or a bug.
constexpr uint32_t LL_SPIN_LIMIT = 1u << 23;

__device__ __forceinline__ void ll_store(ulonglong2* p, double v, uint32_t tag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned long long t = (unsigned long long)tag << 32;
    const unsigned long long w0 = (b & 0xffffffffull) | t, w1 = (b >> 32) | t;
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}

// Element `lane` of rows [r0, r0 + n) (n <= N), summed in ascending row order starting from 0.0. Every poll round
// issues ALL n loads back to back (independent, so they overlap in the memory system: one L2 round trip per round,
// not one per row) and only then inspects the tags; the round repeats until every row carries the tag (rows never
// change once published within an epoch, so re-reading the ones that already matched is harmless). One full warp.
template <int N>
__device__ __forceinline__ double ll_sum_rows(const ulonglong2* rows, uint32_t r0, uint32_t n, uint32_t tag, int lane,
                                              uint32_t* stall, double first = 0.0, bool have_first = false) {
    const ulonglong2* p = rows + (size_t)r0 * LL_ROW + lane;
    unsigned long long w0[N], w1[N];
    const unsigned long long want = ((unsigned long long)tag << 32);
    bool all;
    uint32_t spins = 0;
    do {
#pragma unroll
        for (int k = 0; k < N; ++k)
            if ((uint32_t)k < n && !(have_first && k == 0))
                asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0[k]), "=l"(w1[k]) : "l"(p + (size_t)k * LL_ROW));
        all = true;
#pragma unroll
        for (int k = 0; k < N; ++k)
            if ((uint32_t)k < n && !(have_first && k == 0))
                all = all && ((w0[k] & 0xffffffff00000000ull) == want) && ((w1[k] & 0xffffffff00000000ull) == want);
        if (!all && ((++spins & 0xfffu) == 0u)) {  // watchdog, off the fast path
            if (spins >= LL_SPIN_LIMIT || *reinterpret_cast<volatile uint32_t*>(stall) != 0u) {
                if (atomicCAS(stall, 0u, 1u) == 0u) {
                    stall[1] = blockIdx.x; stall[2] = tag; stall[3] = r0; stall[4] = n; stall[5] = (uint32_t)lane;
                    __threadfence();
                }
                break;
            }
        }
    } while (!all);
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k)
        if ((uint32_t)k < n) {
            const double v = (have_first && k == 0) ? first : __longlong_as_double((long long)((w0[k] & 0xffffffffull) | (w1[k] << 32)));
            s += v;
        }
    return s;
}

// The all-reduce. `v` = this block's row element `lane` (warp 0 calls, all 32 lanes). Block b owns
// chunk b of the n_chunks chunks of the bucket (blocks with b >= n_chunks contribute nothing but still receive the
// total). Returns the total of element `lane` in the fixed grouped order. Level 1: rows travel through global memory in
// the flagged format, the group's first block adds them; level 2: it publishes the group row, every block polls the
// (<= 20) group rows.
__device__ __forceinline__ double ll_allreduce(const LLView& ll, uint32_t parity, uint32_t tag, uint32_t b, uint32_t n_chunks,
                                               double v, int lane) {
    ulonglong2* crows = ll.chunk_rows + (size_t)parity * LL_MAX_CHUNKS * LL_ROW;
    ulonglong2* grows = ll.group_rows + (size_t)parity * LL_MAX_GROUPS * LL_ROW;
    const uint32_t n_groups = (n_chunks + LK_GROUP - 1) / LK_GROUP;
    if (b < n_chunks) {
        if ((b % LK_GROUP) == 0) {
            const uint32_t n = min((uint32_t)LK_GROUP, n_chunks - b);
            const double s = ll_sum_rows<LK_GROUP>(crows, b, n, tag, lane, ll.stall, v, true);
            ll_store(grows + (size_t)(b / LK_GROUP) * LL_ROW + lane, s, tag);
        } else {
            ll_store(crows + (size_t)b * LL_ROW + lane, v, tag);
        }
    }
    return ll_sum_rows<LL_MAX_GROUPS>(grows, 0, n_groups, tag, lane, ll.stall);
}

}  // namespace lk
