// split.h — arguments of the time-axis split kernels (split.cu), shared with abi.cpp.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "program.h"

struct tb_split_args {
    const tb_cexpr* cexpr;
    uint32_t n_cval;
    const tb_split_entry* entries;
    uint32_t n_entries;
    const tb_filter_tab* filt;
    const float* params;  // [n_real][n_params] or NULL
    uint32_t n_params;
    uint32_t sample_rate, state_words;
    uint32_t n_real;      // real voices
    uint32_t n_seg;       // segments per voice (S)
    uint64_t seg;         // samples per segment
    const uint32_t* real_state;  // [n_real][state_words]
    uint32_t* vi;         // [n_real * n_seg][state_words]: initial state of every segment
    uint32_t* vs;         // same shape: the state the render kernels advance (final state after a pass)
    float* cval;          // [n_real][n_cval] scratch: the voices' constant tables
    unsigned long long* inc;  // [n_real][n_entries] scratch: per-sample advance of the analytic entries
    // fused FM voice with a biquad (abi.cpp render_split_fm): the filter's history at a segment's start comes from
    // a warm-up of `warm` samples from zero state instead of a summary pass
    uint32_t* vw;             // [n_real * n_seg][state_words]: states the warm-up starts from / ends in
    unsigned long long* snap; // [n_real * n_seg]: carrier accumulator `seg - warm` samples into the summary pass
    uint64_t warm;
    uint32_t* warm_need;      // device word: the longest warm-up any voice's filter needs (0xffffffff: not contractive)
    int32_t fm_carrier, fm_filter;  // entries of the carrier (SP_SINE_VAR) and of the filter (SP_FILTER)
};

extern "C" cudaError_t tb_split_seed(const tb_split_args* A, cudaStream_t stream);
extern "C" cudaError_t tb_split_fix(const tb_split_args* A, uint32_t level, cudaStream_t stream);
extern "C" cudaError_t tb_split_fm_need(const tb_split_args* A, cudaStream_t stream);
extern "C" cudaError_t tb_split_fm_warm_seed(const tb_split_args* A, cudaStream_t stream);
extern "C" cudaError_t tb_split_fm_adopt(const tb_split_args* A, cudaStream_t stream);
extern "C" cudaError_t tb_split_finish(const tb_split_args* A, uint32_t* real_state, unsigned long long* out_len,
                                       unsigned long long n, int accumulate, cudaStream_t stream);
