// lanes_fm.cu — the lane-per-voice kernel of a program that is ONE fused FM voice (LN_FM, program.h):
// the 65,536-voice FM + low-pass batch.  No interpreter in the kernel: the thread keeps the voice in
// registers and runs the software-pipelined loop of lanes.cuh (run_fm_voice).  A kernel of its own
// because inside the interpreter kernels the register allocation and schedule of that loop move with
// every unrelated edit (+-5 %); the host (abi.cpp launch_lanes) picks it from the lane program.
#include "lanes.cuh"

// 65,536 voices are 6.9 CTAs of 64 threads an SM; measured on B200, a kernel of more than 128 registers holds only
// 6 such CTAs an SM (130 and 144 registers: tb_lanes_occupancy reports 6), so 128 it is — without
// __launch_bounds__(64, 7), under which ptxas spilled the rotation table it now keeps in registers.
#ifndef TB_FM_MAXNREG
#define TB_FM_MAXNREG 128
#endif

extern "C" __global__ void __maxnreg__(TB_FM_MAXNREG)
tb_render_lanes_fm_kernel(const tb_launch P) { lanes_body<false, true>(P, blockIdx.x, 0, P.n_samples, P.accumulate != 0); }
extern "C" __global__ void __maxnreg__(TB_FM_MAXNREG)
tb_render_lanes_fm_mix_kernel(const tb_launch P) { lanes_body<true, true>(P, blockIdx.x, 0, P.n_samples, P.accumulate != 0); }

extern "C" void tb_lanes_fm_kernels(const void** plain, const void** mix) {
    *plain = (const void*)tb_render_lanes_fm_kernel;
    *mix = (const void*)tb_render_lanes_fm_mix_kernel;
}
extern "C" void tb_lanes_fm_run(const tb_launch* P, uint32_t grid, size_t smem, cudaStream_t stream) {
    if (P->mix_partial) tb_render_lanes_fm_mix_kernel<<<grid, LT, smem, stream>>>(*P);
    else tb_render_lanes_fm_kernel<<<grid, LT, smem, stream>>>(*P);
}
