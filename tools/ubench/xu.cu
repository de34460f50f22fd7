// Throughput of the conversion / special-function instructions the renderer leans on (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o xu xu.cu && ./xu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters, float seed) {
    float a[8];
    double d[8];
    for (int i = 0; i < 8; i++) { a[i] = seed + i + threadIdx.x * 1e-3f; d[i] = a[i]; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) a[i] = __sinf(a[i]);                              // FMUL.RZ + MUFU.SIN
            if (OP == 1) { d[i] = (double)a[i]; a[i] = __double_as_longlong(d[i]) & 0xffff; }  // F2F.F64.F32 (+ I2F)
            if (OP == 2) { a[i] = (float)d[i]; d[i] = __longlong_as_double(__float_as_int(a[i]) | 0x3ff0000000000000LL); }  // F2F.F32.F64
            if (OP == 3) a[i] = (float)__float_as_int(a[i]);              // I2F
            if (OP == 4) d[i] = fma(d[i], 1.0000001, 1e-9);               // DFMA
            if (OP == 5) a[i] = fmaf(a[i], 1.0000001f, 1e-9f);            // FFMA
            if (OP == 6) a[i] = __fdividef(1.0f, a[i]);                   // MUFU.RCP + FMUL
        }
    }
    float s = 0; for (int i = 0; i < 8; i++) s += a[i] + (float)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name, int per_iter_ops) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    const int iters = 20000;
    k<OP><<<148 * 8, 256>>>(out, 100, 1.0f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<148 * 8, 256>>>(out, iters, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 8 * 256 * (double)iters * 8 * per_iter_ops;
    printf("%-28s %8.3f ms  %7.2f thread-ops/clk/SM (at 1.965 GHz)\n", name, ms, ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(out);
}
int main() {
    run<0>("sin.approx (FMUL+MUFU.SIN)", 1);
    run<1>("F2F.F64.F32 (+LOP,I2F)", 1);
    run<2>("F2F.F32.F64 (+LOP)", 1);
    run<3>("I2F.F32.S32", 1);
    run<4>("DFMA", 1);
    run<5>("FFMA", 1);
    run<6>("MUFU.RCP+FMUL", 1);
    return 0;
}
