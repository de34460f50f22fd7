// lanes_fm_split.cu — the fused-FM-voice kernel (lanes_fm.cu) for launches of VIRTUAL voices (time-axis
// split, program.h tb_launch::vsplit*).  Compiled apart so that the plain kernel keeps its code.
#define TB_LANES_VSPLIT 1
#include "lanes.cuh"

extern "C" __global__ void __launch_bounds__(TB_LANE_THREADS, TB_LANE_MIN_BLOCKS)
tb_render_lanes_fm_split_kernel(const tb_launch P) { lanes_body<false, true>(P, blockIdx.x, 0, P.n_samples, P.accumulate != 0); }

extern "C" void tb_lanes_fm_split_kernels(const void** plain) { *plain = (const void*)tb_render_lanes_fm_split_kernel; }
extern "C" void tb_lanes_fm_split_run(const tb_launch* P, uint32_t grid, size_t smem, cudaStream_t stream) {
    tb_render_lanes_fm_split_kernel<<<grid, LT, smem, stream>>>(*P);
}
