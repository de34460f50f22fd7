// lanes_fm_split.cu — the fused-FM-voice kernel (lanes_fm.cu) for launches of VIRTUAL voices (time-axis
// split, program.h tb_launch::vsplit*).  Compiled apart so that the plain kernel keeps its code.
#define TB_LANES_VSPLIT 1
#include "lanes.cuh"

// 65,536 voices are 6.9 CTAs of 64 threads an SM; measured on B200, a kernel of more than 128 registers holds only
// 6 such CTAs an SM (130 and 144 registers: tb_lanes_occupancy reports 6), so 128 it is — without
// __launch_bounds__(64, 7), under which ptxas spilled the rotation table it now keeps in registers.
#ifndef TB_FM_MAXNREG
#define TB_FM_MAXNREG 128
#endif

extern "C" __global__ void __maxnreg__(TB_FM_MAXNREG)
tb_render_lanes_fm_split_kernel(const tb_launch P) { lanes_body<false, true>(P, blockIdx.x, 0, P.n_samples, P.accumulate != 0); }

// The summary pass (lanes.cuh run_fm_sums): phase sums only.
extern "C" __global__ void __maxnreg__(TB_FM_MAXNREG)
tb_render_lanes_fm_sums_kernel(const tb_launch P) { lanes_body<false, true, true>(P, blockIdx.x, 0, P.n_samples, P.accumulate != 0); }

extern "C" void tb_lanes_fm_split_kernels(const void** plain, const void** sums) {
    *plain = (const void*)tb_render_lanes_fm_split_kernel;
    *sums = (const void*)tb_render_lanes_fm_sums_kernel;
}
extern "C" void tb_lanes_fm_split_run(const tb_launch* P, uint32_t grid, size_t smem, cudaStream_t stream) {
    if (P->fm_sums) tb_render_lanes_fm_sums_kernel<<<grid, LT, smem, stream>>>(*P);
    else tb_render_lanes_fm_split_kernel<<<grid, LT, smem, stream>>>(*P);
}
