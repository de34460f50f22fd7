// lanes.cu — the lane-per-voice interpreter kernels (lanes.cuh), one unit per CTA, and the host-side
// entry points of the lane path.
#include <algorithm>

#include <cstdio>
#include "lanes.cuh"

// grid = ceil(n_voices / LT) CTAs of LT voices; P.n_samples is a multiple of TB_LS; P.out points at the
// first sample of this launch.
extern "C" __global__ void __launch_bounds__(TB_LANE_THREADS, TB_LANE_MIN_BLOCKS)
tb_render_lanes_kernel(const tb_launch P) { lanes_body<false, false>(P, blockIdx.x, 0, P.n_samples, P.accumulate != 0); }
// The same program with the rows summed on the chip instead of stored (tb_launch::mix_partial).
extern "C" __global__ void __launch_bounds__(TB_LANE_THREADS, TB_LANE_MIN_BLOCKS)
tb_render_lanes_mix_kernel(const tb_launch P) { lanes_body<true, false>(P, blockIdx.x, 0, P.n_samples, P.accumulate != 0); }

extern "C" void tb_lanes_queue_kernels(const void** plain, const void** mix);  // lanes_queue.cu
extern "C" void tb_lanes_fm_kernels(const void** plain, const void** mix);     // lanes_fm.cu
extern "C" void tb_lanes_split_kernels(const void** plain);                    // lanes_split.cu
extern "C" void tb_lanes_fm_split_kernels(const void** plain, const void** sums);  // lanes_fm_split.cu
extern "C" void tb_lanes_split_run(const tb_launch* P, uint32_t grid, size_t smem, cudaStream_t stream);
extern "C" void tb_lanes_fm_split_run(const tb_launch* P, uint32_t grid, size_t smem, cudaStream_t stream);
extern "C" void tb_lanes_queue_run(const tb_launch* P, uint32_t grid, size_t smem, cudaStream_t stream);
extern "C" void tb_lanes_fm_run(const tb_launch* P, uint32_t grid, size_t smem, cudaStream_t stream);

extern "C" size_t tb_lanes_smem_bytes(uint32_t n_lane_code, uint32_t w_words, uint32_t q_units, uint32_t slots) {
    return (size_t)n_lane_code * sizeof(tb_insn) + (size_t)(q_units + 4 * slots) * LT * 16 +
           (size_t)8 * AS * 16 + (((size_t)w_words * LT * 4 + 15) & ~(size_t)15);
}

// kind: 0 = interpreter kernels, one unit per CTA; 1 = the same behind the work queue (tb_launch::lane_queue);
//       2 = the fused-FM-voice kernels (lanes_fm.cu).
// Every kernel that may take a lane program is opted in to the largest shared-memory size any program of this
// process has needed on the device (never lowered: a program created later with a smaller footprint must not
// take the opt-in away from one that is still rendering).
static cudaError_t ensure_lane_smem(size_t smem) {
    static size_t configured[64] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    size_t& have = configured[dev & 63];
    if (smem <= 48 * 1024 || smem <= have) return cudaSuccess;
    const void* ks[9] = {(const void*)tb_render_lanes_kernel, (const void*)tb_render_lanes_mix_kernel};
    tb_lanes_queue_kernels(&ks[2], &ks[3]);
    tb_lanes_fm_kernels(&ks[4], &ks[5]);
    tb_lanes_split_kernels(&ks[6]);
    tb_lanes_fm_split_kernels(&ks[7], &ks[8]);
    for (const void* k : ks) {
        e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    have = smem;
    return cudaSuccess;
}

extern "C" cudaError_t tb_lanes_launch(const tb_launch* P, size_t smem, int kind, cudaStream_t stream) {
    cudaError_t se = ensure_lane_smem(smem);
    if (se != cudaSuccess) return se;
    const uint32_t groups = (P->n_voices + LT - 1) / LT;
    const bool vsplit = P->vsplit_total > 1;  // virtual voices (time-axis split): the kernels of lanes_*split.cu; no mixdown
    if (vsplit && P->mix_partial) return cudaErrorInvalidValue;
    if (vsplit && kind == 2) tb_lanes_fm_split_run(P, groups, smem, stream);
    else if (vsplit) tb_lanes_split_run(P, groups, smem, stream);
    else if (kind == 1) tb_lanes_queue_run(P, std::min(groups * P->lane_segs, P->lane_grid), smem, stream);
    else if (kind == 2) tb_lanes_fm_run(P, groups, smem, stream);
    else if (P->mix_partial) tb_render_lanes_mix_kernel<<<groups, LT, smem, stream>>>(*P);
    else tb_render_lanes_kernel<<<groups, LT, smem, stream>>>(*P);
    return cudaGetLastError();
}

// Resident CTAs per SM at this shared-memory size (kind as in tb_lanes_launch).
extern "C" cudaError_t tb_lanes_occupancy(size_t smem, int kind, int* blocks_per_sm, int* n_sm) {
    const void* k = (const void*)tb_render_lanes_kernel;
    const void* dummy = nullptr;
    if (kind == 1) tb_lanes_queue_kernels(&k, &dummy);
    if (kind == 2) tb_lanes_fm_kernels(&k, &dummy);
    cudaError_t e = ensure_lane_smem(smem);
    if (e != cudaSuccess) return e;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, LT, smem);
    if (e != cudaSuccess) return e;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    return cudaDeviceGetAttribute(n_sm, cudaDevAttrMultiProcessorCount, dev);
}
