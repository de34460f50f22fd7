// How fast can 65,536 rows be written when every row receives PIECE contiguous bytes per visit and all
// rows advance in lockstep (the store side of the lane-per-voice kernel)?  One warp owns 32 rows.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rowstore rowstore.cu && ./rowstore
#include <cstdio>
#include <cuda_runtime.h>
template <int PIECE>  // bytes per row per visit: 64, 128, 256, 512
__global__ void k(float* out, size_t stride, int visits) {
    const int l = threadIdx.x & 31;
    const size_t warp = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
    constexpr int LPR = PIECE / 16;       // lanes per row
    constexpr int RPI = 32 / LPR;         // rows per instruction
    float* base = out + (warp * 32 + l / LPR) * stride + (l % LPR) * 4;
    const float4 v = make_float4(1.f, 2.f, 3.f, (float)l);
    for (int t = 0; t < visits; t++) {
#pragma unroll
        for (int i = 0; i < LPR; i++) __stcg(reinterpret_cast<float4*>(base + (size_t)i * RPI * stride), v);
        base += PIECE / 4;
    }
}
template <int PIECE> void run(float* out, size_t stride, size_t n) {
    const int visits = (int)(n * 4 / PIECE);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<PIECE><<<1024, 64>>>(out, stride, visits);
    cudaEventRecord(e0);
    k<PIECE><<<1024, 64>>>(out, stride, visits);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("piece %4d B: %7.2f ms  %6.0f GB/s\n", PIECE, ms, 65536.0 * visits * PIECE / (ms * 1e-3) / 1e9);
}
int main() {
    const size_t n = 220512, stride = n;  // 5 s rows, 128-byte aligned
    float* out; cudaMalloc(&out, 65536 * stride * 4);
    run<64>(out, stride, n); run<128>(out, stride, n); run<256>(out, stride, n); run<512>(out, stride, n);
    return 0;
}
