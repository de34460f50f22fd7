// slots.cu — where do the warps of 64-thread CTAs land?  For every warp of a grid shaped like the
// two-warps-a-voice FM kernel (lanes_fm_ws.cu: 2,048 CTAs of 64 threads, 72 registers, 14 CTAs an SM) record
// (%smid, %warpid); the scheduler a warp belongs to is %warpid mod 4.  Prints, for a few SMs, the warp
// slots of the resident CTAs, and over all SMs how often the two warps of a CTA share a scheduler and how
// the "warp 0" role is spread over the four schedulers.
//   nvcc -arch=sm_100a -o slots slots.cu && ./slots
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(64, 14) probe(uint32_t* rec, long long spin) {
    extern __shared__ unsigned char smem[];
    uint32_t smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    const long long t0 = clock64();
    while (clock64() - t0 < spin) { smem[threadIdx.x] = (unsigned char)warpid; }   // keep every CTA resident at once
    if ((threadIdx.x & 31) == 0) rec[blockIdx.x * 2 + (threadIdx.x >> 5)] = (smid << 8) | warpid;
}

int main() {
    const int grid = 2048;
    uint32_t* d;
    cudaMalloc(&d, grid * 2 * 4);
    probe<<<grid, 64, 13616>>>(d, 2000000);
    cudaDeviceSynchronize();
    std::vector<uint32_t> h(grid * 2);
    cudaMemcpy(h.data(), d, grid * 2 * 4, cudaMemcpyDeviceToHost);
    int same = 0, w0_sched[4] = {0, 0, 0, 0}, w1_sched[4] = {0, 0, 0, 0}, role_sched[2][4] = {};
    for (int c = 0; c < grid; c++) {
        const uint32_t a = h[2 * c] & 255, b = h[2 * c + 1] & 255;
        if ((a & 3) == (b & 3)) same++;
        w0_sched[a & 3]++;
        w1_sched[b & 3]++;
        const uint32_t pw = (a >> 2) & 1u;   // lanes_fm_ws.cu: which warp takes the phase role
        role_sched[0][(pw == 0 ? a : b) & 3]++;
        role_sched[1][(pw == 0 ? b : a) & 3]++;
    }
    printf("CTAs whose two warps share a scheduler: %d of %d\n", same, grid);
    printf("warp 0 by scheduler: %d %d %d %d   warp 1: %d %d %d %d\n", w0_sched[0], w0_sched[1], w0_sched[2], w0_sched[3],
           w1_sched[0], w1_sched[1], w1_sched[2], w1_sched[3]);
    printf("role swap by (slot >> 2) & 1: phase warps by scheduler %d %d %d %d, tone warps %d %d %d %d\n", role_sched[0][0],
           role_sched[0][1], role_sched[0][2], role_sched[0][3], role_sched[1][0], role_sched[1][1], role_sched[1][2],
           role_sched[1][3]);
    for (int sm = 0; sm < 3; sm++) {
        printf("SM %d:", sm);
        for (int c = 0; c < grid; c++)
            if ((h[2 * c] >> 8) == (uint32_t)sm) printf(" cta%d(%u,%u)", c, h[2 * c] & 255, h[2 * c + 1] & 255);
        printf("\n");
    }
    int per_sm[256] = {};
    for (int c = 0; c < grid; c++) per_sm[h[2 * c] >> 8]++;
    int mn = 1 << 30, mx = 0;
    for (int s = 0; s < 148; s++) { if (per_sm[s] < mn) mn = per_sm[s]; if (per_sm[s] > mx) mx = per_sm[s]; }
    printf("CTAs per SM: min %d max %d\n", mn, mx);
    return 0;
}
