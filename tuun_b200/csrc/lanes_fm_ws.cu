// lanes_fm_ws.cu — the kernels of lanes_fm_ws.cuh over real voices (rows, or the on-chip mixdown), and the host-side
// entry points of the two-warps-a-voice form.
#include "lanes_fm_ws.cuh"

#ifndef TB_FM_WS_MAXNREG
#define TB_FM_WS_MAXNREG 64
#endif

extern "C" __global__ void __maxnreg__(TB_FM_WS_MAXNREG) tb_render_lanes_fm_ws_kernel(const tb_launch P) { fm_ws_body<false>(P); }
extern "C" __global__ void __maxnreg__(TB_FM_WS_MAXNREG) tb_render_lanes_fm_ws_mix_kernel(const tb_launch P) { fm_ws_body<true>(P); }

extern "C" void tb_lanes_fm_ws_kernels(const void** plain, const void** mix) {
    *plain = (const void*)tb_render_lanes_fm_ws_kernel;
    *mix = (const void*)tb_render_lanes_fm_ws_mix_kernel;
}
// Shared memory of one CTA (32 columns).
extern "C" size_t tb_lanes_fm_ws_smem_bytes(uint32_t n_lane_code, uint32_t w_words, uint32_t q_units, uint32_t slots) {
    return (size_t)n_lane_code * sizeof(tb_insn) + (size_t)(q_units + 4 * slots) * LT * 16 + (size_t)8 * AS * 16 +
           (((size_t)w_words * LT * 4 + 15) & ~(size_t)15);
}
extern "C" void tb_lanes_fm_ws_run(const tb_launch* P, size_t smem, cudaStream_t stream) {
    const uint32_t grid = (P->n_voices + 31u) / 32u;
    if (P->mix_partial) tb_render_lanes_fm_ws_mix_kernel<<<grid, WS_THREADS, smem, stream>>>(*P);
    else tb_render_lanes_fm_ws_kernel<<<grid, WS_THREADS, smem, stream>>>(*P);
}
extern "C" cudaError_t tb_lanes_fm_ws_occupancy(size_t smem, int* blocks_per_sm) {
    const void* k = (const void*)tb_render_lanes_fm_ws_kernel;
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, (int)WS_THREADS, smem);
}
