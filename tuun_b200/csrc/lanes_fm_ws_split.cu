// lanes_fm_ws_split.cu — the two-warps-a-voice kernel of lanes_fm_ws.cuh over VIRTUAL voices: the segments of a time-axis
// split (abi.cpp split_pass; program.h tb_launch::vsplit*).  Parameters and the row belong to the voice, the state block
// and out_len to the segment; the warm-up and the samples pass of a batch of fused FM voices cut in time (abi.cpp
// render_split_fm: one GPU's share of the 65,536-voice batch under strong scaling, few voices over long renders, the
// ranks of a time-sharded render) run here.  Compiled apart so that the kernels for real voices keep their code.
#define TB_LANES_VSPLIT 1
#include "lanes_fm_ws.cuh"

#ifndef TB_FM_WS_MAXNREG
#define TB_FM_WS_MAXNREG 64
#endif

extern "C" __global__ void __maxnreg__(TB_FM_WS_MAXNREG) tb_render_lanes_fm_ws_split_kernel(const tb_launch P) { fm_ws_body<false>(P); }

extern "C" void tb_lanes_fm_ws_split_run(const tb_launch* P, size_t smem, cudaStream_t stream) {
    const uint32_t grid = (P->n_voices + 31u) / 32u;
    tb_render_lanes_fm_ws_split_kernel<<<grid, WS_THREADS, smem, stream>>>(*P);
}
extern "C" cudaError_t tb_lanes_fm_ws_split_occupancy(size_t smem, int* blocks_per_sm) {
    const void* k = (const void*)tb_render_lanes_fm_ws_split_kernel;
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k, (int)WS_THREADS, smem);
}
