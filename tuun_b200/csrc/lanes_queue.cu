// lanes_queue.cu — the lane-per-voice interpreter kernels behind a work queue.
//
// When the device cannot hold ceil(n_voices / 64) CTAs at once, a plain grid runs a second, mostly empty
// wave for the whole duration of the render.  Here a persistent grid takes (time segment, voice group)
// units off a counter in segment-major order, so every SM stays busy to the end.  Unit (s, g) continues
// the state that unit (s-1, g) wrote: taken earlier by construction, normally long finished, else
// awaited.  (Separate kernels, separate file: in one kernel with the plain form the register
// allocation of the tile loops suffers.)
#include "lanes.cuh"

namespace {
template <bool MIX>
__device__ __forceinline__ void lanes_queue_kernel(const tb_launch& P) {
    __shared__ uint32_t unit_s;
    uint32_t* counter = P.lane_queue;
    uint32_t* progress = P.lane_queue + 1;  // [lane_groups]: segments finished
    const uint32_t n_units = P.lane_groups * P.lane_segs;
    for (;;) {
        __syncthreads();  // the previous unit is done with shared memory
        if (threadIdx.x == 0) unit_s = atomicAdd(counter, 1u);
        __syncthreads();
        const uint32_t u = unit_s;
        if (u >= n_units) break;
        const uint32_t g = u % P.lane_groups, s = u / P.lane_groups;
        if (s > 0) {
            if (threadIdx.x == 0) {
                while (*reinterpret_cast<volatile uint32_t*>(progress + g) < s) __nanosleep(200);
                __threadfence();
            }
            __syncthreads();
        }
        const u64 s0 = (u64)s * P.lane_seg_samples;
        const u64 ns = s0 + P.lane_seg_samples <= P.n_samples ? P.lane_seg_samples : P.n_samples - s0;
        lanes_body<MIX, false>(P, g, s0, ns, P.accumulate != 0 || s > 0);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicExch(progress + g, s + 1u);
    }
}
}  // namespace

extern "C" __global__ void __launch_bounds__(TB_LANE_THREADS, TB_LANE_MIN_BLOCKS)
tb_render_lanes_queue_kernel(const tb_launch P) { lanes_queue_kernel<false>(P); }
extern "C" __global__ void __launch_bounds__(TB_LANE_THREADS, TB_LANE_MIN_BLOCKS)
tb_render_lanes_mix_queue_kernel(const tb_launch P) { lanes_queue_kernel<true>(P); }

extern "C" void tb_lanes_queue_kernels(const void** plain, const void** mix) {
    *plain = (const void*)tb_render_lanes_queue_kernel;
    *mix = (const void*)tb_render_lanes_mix_queue_kernel;
}
extern "C" void tb_lanes_queue_run(const tb_launch* P, uint32_t grid, size_t smem, cudaStream_t stream) {
    if (P->mix_partial) tb_render_lanes_mix_queue_kernel<<<grid, LT, smem, stream>>>(*P);
    else tb_render_lanes_queue_kernel<<<grid, LT, smem, stream>>>(*P);
}
