// lk_kernels.h — host-callable launchers of the device kernels (one translation unit each).
#pragma once
#include <cuda_runtime.h>

#include "lk_device.cuh"

namespace lk {

// Per (step, scan) description of the bucket processed in that step; built on the host at
// stage time from scan_offsets / bucket_offsets / bucket_times.
struct StepInit {
    uint32_t chunk_begin, chunk_end;
    uint32_t pt_begin, pt_end;
    double t_bucket;
    uint32_t active;
    uint32_t pad;
};

struct ResidualArgs {
    const float4* pts;
    const HashSlot* slots;
    uint32_t hash_mask;
    const MapNode* nodes;
    const ChunkDesc* chunks;  // chunk table of the whole staged batch
    uint32_t chunk_first;     // first chunk of this launch (grid.x = number of chunks)
    ScanConst* sc;            // [batch]
    ScanStep* step;           // [batch]
    double* partial;          // [total_chunks * PARTIAL_STRIDE]
    uint32_t* ticket;         // [batch]
    double* x;                // [batch * 36]
    double* P;                // [batch * 900]
    lk_stream_clock* clk;     // [batch]
    uint32_t* n_eff;          // [batch] accumulated success_pts_size_out
    int last_iter;
    // debug (lk_debug_residuals): per-point rows instead of accumulation
    uint8_t* dbg_ok;
    double* dbg_h;
    double* dbg_z;
    double* dbg_R;
    int32_t* dbg_key;
    Globals g;
};

void launch_residual(const ResidualArgs& a, uint32_t n_chunks, bool debug, int gather_mode, cudaStream_t s);

struct PredictArgs {
    const StepInit* init;  // [batch] for this step
    ScanStep* step;        // [batch]
    ScanConst* sc;         // [batch]
    double* x;             // [batch*36]
    double* P;             // [batch*900]
    const double* Q;       // [900]
    lk_stream_clock* clk;  // [batch]
    uint32_t* ticket;
    int batch;
};
void launch_predict_prepare(const PredictArgs& a, cudaStream_t s);

// Plain ESKF::predict on `batch` filters with explicit dt (lk_predict).
void launch_predict_dt(double* x, double* P, const double* Q, const double* dt, int batch, int prop_state,
                       int prop_cov, cudaStream_t s);

struct ReprojectArgs {
    const float4* pts;
    float4* world;
    const ChunkDesc* chunks;
    uint32_t chunk_first;
    const ScanConst* sc;
    const ScanStep* step;
    Globals g;
};
void launch_reproject(const ReprojectArgs& a, uint32_t n_chunks, cudaStream_t s);

}  // namespace lk
