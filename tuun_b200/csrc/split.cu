// split.cu — time-axis split of a steady program (program.h tb_split_entry; abi.cpp render_split_round, render_split_fm).
//
// The reference streams a waveform block by block and keeps the cursor in the tree's own State
// (generator.rs:12-35, :76-85).  For a steady program every piece of that state is either a
// position (Time, Noise), a sum of phase increments (Sine, generator.rs:212-218) or the history of
// a constant-coefficient Filter (generator.rs:382-515) — all of which obey an associative law over
// time.  So ONE voice can be rendered as S segments side by side ("virtual voices" of the render
// kernels) as soon as each segment knows the state it starts from:
//   tb_split_seed   copies the voice's state to its S segments and advances the analytic entries
//                   (positions + s L, constant-rate accumulators + inc s L, exact in 2^-64 turns);
//   tb_split_fix    after a render pass, turns the segments' (initial, final) states of one level
//                   into right initial states: exclusive u64 prefix sums for Sine accumulators —
//                   exact and associative, the result does not depend on S — and a scan of affine
//                   maps X' = M^L X + z for Filter histories (f64);
//   tb_split_finish hands the last segment's final state back to the voice.
// One warp per real voice.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tuun_b200.h"
#include "program.h"
#include "split.h"

namespace {

#include "common.cuh"

constexpr int MAXD = TB_MAX_K - 1 + TB_MAX_J;  // 12: dimension of a filter's state (K-1 inputs, J outputs)

__device__ __forceinline__ u64 ldu64(const uint32_t* s, uint32_t w) { return (u64)s[w] | ((u64)s[w + 1] << 32); }
__device__ __forceinline__ void stu64(uint32_t* s, uint32_t w, u64 v) {
    s[w] = (uint32_t)v;
    s[w + 1] = (uint32_t)(v >> 32);
}

// Per real voice: the constant table (is_const folding, generator.rs:574-612 — the arithmetic of the
// render kernels' setup) and the phase increment of every constant-rate sine.
__global__ void split_prepare_kernel(const tb_split_args A) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= A.n_real) return;
    float* cv = A.cval + (size_t)v * A.n_cval;
    const float* prow = A.params ? A.params + (size_t)v * A.n_params : nullptr;
    for (uint32_t k = 0; k < A.n_cval; k++) {
        const tb_cexpr e = A.cexpr[k];
        float x;
        if (e.kind == CE_LIT) x = e.value;
        else if (e.kind == CE_PARAM) x = prow ? prow[e.a] : e.value;
        else if (e.kind == CE_NEG) x = -cv[e.a];
        else x = apply1(e.op, cv[e.a], cv[e.b]);
        cv[k] = x;
    }
    for (uint32_t k = 0; k < A.n_entries; k++) {
        const tb_split_entry e = A.entries[k];
        u64 inc = 0;
        if (e.kind == SP_POS || (e.kind == SP_CLK && e.a < 0)) inc = 1ull;
        else if (e.kind == SP_SINE_CONST || e.kind == SP_CLK)
            inc = turns_to_fx_slow((double)cv[e.a] / (TB_TAU * (double)A.sample_rate));
        A.inc[(size_t)v * A.n_entries + k] = inc;
    }
}

// Per virtual voice: the real voice's state, analytic entries advanced to the segment's first sample.
__global__ void split_seed_kernel(const tb_split_args A) {
    const uint32_t vv = blockIdx.x * blockDim.x + threadIdx.x;
    if (vv >= A.n_real * A.n_seg) return;
    const uint32_t v = vv / A.n_seg, s = vv - v * A.n_seg;
    const uint32_t* src = A.real_state + (size_t)v * A.state_words;
    uint32_t* dst = A.vi + (size_t)vv * A.state_words;
    for (uint32_t k = 0; k < A.state_words; k++) dst[k] = src[k];
    const u64 n = (u64)s * A.seg;
    for (uint32_t k = 0; k < A.n_entries; k++) {
        const tb_split_entry e = A.entries[k];
        if (e.kind == SP_POS || e.kind == SP_SINE_CONST)
            stu64(dst, e.state_off, ldu64(dst, e.state_off) + A.inc[(size_t)v * A.n_entries + k] * n);
        else if (e.kind == SP_RESET_SIGN)
            dst[e.state_off + 1] = 0u;  // "class of the first sample seen": recorded anew by every pass
    }
}

// ---- small dense f64 matrices in shared memory, one warp ------------------------------------------
__device__ __forceinline__ void mat_mul(double* C, const double* Am, const double* Bm, int D, int l) {
    for (int e = l; e < D * D; e += 32) {
        const int r = e / D, c = e % D;
        double s = 0.0;
        for (int k = 0; k < D; k++) s = fma(Am[r * D + k], Bm[k * D + c], s);
        C[e] = s;
    }
    __syncwarp();
}
__device__ __forceinline__ void mat_copy(double* C, const double* Am, int D, int l) {
    for (int e = l; e < D * D; e += 32) C[e] = Am[e];
    __syncwarp();
}
__device__ __forceinline__ void mat_vec(double (&y)[MAXD], const double* Mx, const double (&x)[MAXD], int D) {
    for (int r = 0; r < D; r++) {
        double s = 0.0;
        for (int k = 0; k < D; k++) s = fma(Mx[r * D + k], x[k], s);
        y[r] = s;
    }
}

// One warp per real voice: every entry of `level`.
__global__ void __launch_bounds__(32) split_fix_kernel(const tb_split_args A, uint32_t level) {
    __shared__ double pw[6][MAXD * MAXD];  // pw[k] = M^(L 2^k), k < 5; pw[5] scratch
    __shared__ double tmp[2][MAXD * MAXD];
    const uint32_t v = blockIdx.x;
    const int l = threadIdx.x;
    const uint32_t S = A.n_seg;
    const size_t vv0 = (size_t)v * S;
    // (SP_CLK reads the sign words the segments STARTED from: SP_RESET_SIGN overwrites them in a second round)
    for (uint32_t round = 0; round < 2; round++)
    for (uint32_t k = 0; k < A.n_entries; k++) {
        const tb_split_entry e = A.entries[k];
        if (e.level != level || (e.kind == SP_RESET_SIGN) != (round == 1)) continue;
        if (e.kind == SP_SINE_VAR) {
            // accumulator at the start of segment s = accumulator of the voice + the increments of segments < s
            u64 carry = ldu64(A.vi + vv0 * A.state_words, e.state_off);
            for (uint32_t c0 = 0; c0 < S; c0 += 32) {
                const uint32_t s = c0 + l;
                u64 d = 0;
                if (s < S) {
                    const size_t o = (vv0 + s) * A.state_words;
                    d = ldu64(A.vs + o, e.state_off) - ldu64(A.vi + o, e.state_off);
                }
                const u64 incl = warp_incl_sum(d);
                __syncwarp();
                // every lane has read its segment's old initial state: the chunk's first segment gets the carry
                // (written here, not by lane 31 of the previous chunk, which would have overwritten it unread)
                if (l == 0 && c0 > 0) stu64(A.vi + (vv0 + s) * A.state_words, e.state_off, carry);
                if (l < 31 && s + 1 < S) stu64(A.vi + (vv0 + s + 1) * A.state_words, e.state_off, carry + incl);
                carry += __shfl_sync(FULL, incl, 31);
                __syncwarp();
            }
        } else if (e.kind == SP_RESET_SIGN) {
            // the class of the last trigger sample carries over (the sign word does not depend on where it started)
            for (uint32_t s = 1 + l; s < S; s += 32)
                A.vi[(vv0 + s) * A.state_words + e.state_off] = A.vs[(vv0 + s - 1) * A.state_words + e.state_off];
            __syncwarp();
        } else if (e.kind == SP_CLK) {
            // A segment that restarted its clock ends at an absolute value; one that did not added rate x L to what
            // it started from.  Inclusive "last set" scan: (set, value) o (set', value') = set' ? (1, value')
            // : (set, value + value').
            const u64 step = A.inc[(size_t)v * A.n_entries + k] * A.seg;
            u64 carry = ldu64(A.vi + vv0 * A.state_words, e.state_off);  // the clock at the start of the chunk
            for (uint32_t c0 = 0; c0 < S; c0 += 32) {
                const uint32_t s = c0 + l;
                u64 val = 0;
                bool set = false;
                if (s < S) {
                    const size_t o = (vv0 + s) * A.state_words;
                    const u64 f = ldu64(A.vs + o, e.state_off), i0 = ldu64(A.vi + o, e.state_off);
                    // a restart on the first sample: what the pass saw (from the class it was started with) and
                    // what is true (from the class the previous segment really ended in)
                    const bool first_nonneg = A.vs[o + e.b + 1] == 2u;
                    const bool guess_neg = A.vi[o + e.b] == 0u;
                    const bool true_neg = s == 0 ? guess_neg : A.vs[o - A.state_words + e.b] == 0u;
                    const bool saw0 = guess_neg && first_nonneg, true0 = true_neg && first_nonneg;
                    if (true0) {         // the clock restarted on the first sample: the final value is absolute —
                        set = true;      // what the pass computed if it saw that restart or a later one, else L steps
                        val = (saw0 || (f - i0) != step) ? f : step;
                    } else if (saw0) {   // spurious: undone, unless a later restart made the final value absolute anyway
                        set = f != step;
                        val = set ? f : step;
                    } else {
                        set = (f - i0) != step;
                        val = set ? f : step;
                    }
                }
                if (l == 0 && !set) {  // the chunk's first segment continues the carried clock
                    val += carry;
                    set = true;
                }
                UNROLL for (int d = 1; d < 32; d <<= 1) {
                    const u64 pv = __shfl_up_sync(FULL, val, d);
                    const bool ps = __shfl_up_sync(FULL, set ? 1 : 0, d) != 0;
                    if (l >= d && !set) {
                        val += pv;
                        set = ps;
                    }
                }
                __syncwarp();
                if (l == 0 && c0 > 0) stu64(A.vi + (vv0 + s) * A.state_words, e.state_off, carry);
                if (l < 31 && s + 1 < S) stu64(A.vi + (vv0 + s + 1) * A.state_words, e.state_off, val);
                carry = __shfl_sync(FULL, val, 31);
                __syncwarp();
            }
        } else if (e.kind == SP_FILTER) {
            const tb_filter_tab* ft = &A.filt[e.a];
            const int K = (int)ft->K, J = (int)ft->J, D = K - 1 + J;
            if (D <= 0 || D > MAXD) continue;
            const float* cv = A.cval + (size_t)v * A.n_cval;
            // M: one step of the homogeneous system on X = [x[n-K+1..n-1], y[n-J..n-1]] (oldest first,
            // the layout of the state block behind its two header words), generator.rs:482-507.
            for (int en = l; en < D * D; en += 32) {
                const int r = en / D, c = en % D;
                double m = 0.0;
                if (r < K - 1) {
                    m = (r + 1 < K - 1 && c == r + 1) ? 1.0 : 0.0;  // shift; the newest input of the homogeneous system is 0
                } else if (r < D - 1) {
                    m = (c == r + 1) ? 1.0 : 0.0;
                } else if (J > 0) {  // the new output
                    if (c < K - 1) m = (double)cv[~ft->coef[K - 1 - c]];         // b_i x[n-i], i = K-1-c
                    else m = -(double)cv[~ft->coef[K + (J - 1 - (c - (K - 1)))]];  // -a_j y[n-j], j = J - (c - (K-1))
                }
                tmp[0][en] = m;
            }
            __syncwarp();
            // pw[0] = M^L by binary exponentiation (pw[5] = running result, tmp[0] = running square)
            for (int en = l; en < D * D; en += 32) pw[5][en] = (en / D == en % D) ? 1.0 : 0.0;
            __syncwarp();
            for (u64 n = A.seg; n != 0; n >>= 1) {
                if (n & 1ull) {
                    mat_mul(tmp[1], pw[5], tmp[0], D, l);
                    mat_copy(pw[5], tmp[1], D, l);
                }
                if (n > 1ull) {
                    mat_mul(tmp[1], tmp[0], tmp[0], D, l);
                    mat_copy(tmp[0], tmp[1], D, l);
                }
            }
            mat_copy(pw[0], pw[5], D, l);
            for (int q = 1; q < 5; q++) mat_mul(pw[q], pw[q - 1], pw[q - 1], D, l);
            const uint32_t so = e.state_off + 2;  // behind the header words (initialised, inputs held)
            double X0[MAXD];                      // state at the start of the chunk's first segment
            {
                const uint32_t* st = A.vi + vv0 * A.state_words;
                for (int r = 0; r < D; r++) X0[r] = (double)__uint_as_float(st[so + r]);
            }
            for (uint32_t c0 = 0; c0 < S; c0 += 32) {
                const uint32_t s = c0 + l;
                double W[MAXD], t[MAXD];
                for (int r = 0; r < D; r++) W[r] = 0.0;
                if (s < S) {  // z = final - M^L initial: the segment's response from zero state
                    const size_t o = (vv0 + s) * A.state_words;
                    double I[MAXD];
                    for (int r = 0; r < D; r++) I[r] = (double)__uint_as_float(A.vi[o + so + r]);
                    mat_vec(t, pw[0], I, D);
                    for (int r = 0; r < D; r++) W[r] = (double)__uint_as_float(A.vs[o + so + r]) - t[r];
                }
                if (l == 0) {
                    mat_vec(t, pw[0], X0, D);
                    for (int r = 0; r < D; r++) W[r] += t[r];
                }
                for (int q = 0; q < 5; q++) {  // inclusive scan of X' = M^L X + z over the 32 segments of the chunk
                    const int d = 1 << q;
                    double U[MAXD];
                    for (int r = 0; r < D; r++) U[r] = __shfl_up_sync(FULL, W[r], d);
                    if (l >= d) {
                        mat_vec(t, pw[q], U, D);
                        for (int r = 0; r < D; r++) W[r] += t[r];
                    }
                }
                // W = state at the start of segment s + 1.  The chunk's first segment gets the carried state here
                // (every lane has read its old initial state by now), the others from the lane before them.
                __syncwarp();
                if (l == 0 && c0 > 0) {
                    uint32_t* st = A.vi + (vv0 + s) * A.state_words;
                    for (int r = 0; r < D; r++) st[so + r] = __float_as_uint((float)X0[r]);
                }
                if (l < 31 && s + 1 < S) {
                    uint32_t* st = A.vi + (vv0 + s + 1) * A.state_words;
                    for (int r = 0; r < D; r++) st[so + r] = __float_as_uint((float)W[r]);
                }
                for (int r = 0; r < D; r++) X0[r] = __shfl_sync(FULL, W[r], 31);
                __syncwarp();
            }
        }
    }
}

// ---- fused FM voice with a biquad: filter histories by warm-up (abi.cpp render_split_fm) ---------------------
// How many samples until a biquad's zero-input response has decayed below 1e-11 of where it started: the filter's
// history at a segment's start then follows, to f32 precision, from rendering that many samples before it from
// ZERO history — no summary pass over the whole segment.  Poles of 1 + a1 z^-1 + a2 z^-2; the bound leaves room
// for the 1 / sin(pole angle) a resonant pair's response carries (35 for a 200 Hz pair at 44.1 kHz).
__global__ void split_fm_need_kernel(const tb_split_args A) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= A.n_real) return;
    const tb_filter_tab* ft = &A.filt[A.entries[A.fm_filter].a];
    const float* cv = A.cval + (size_t)v * A.n_cval;
    const double a1 = (double)cv[~ft->coef[ft->K]], a2 = (double)cv[~ft->coef[ft->K + 1]];
    const double disc = a1 * a1 - 4.0 * a2;
    double r;
    if (disc < 0.0) r = sqrt(a2);
    else r = 0.5 * (fabs(a1) + sqrt(disc));
    uint32_t need = 0xffffffffu;
    if (r < 0.9999 && r == r) need = r <= 1e-3 ? 16u : (uint32_t)ceil(log(1e-11) / log(r));
    atomicMax(A.warm_need, need);
}
// The state a segment's warm-up starts from, `warm` samples before the segment: positions and constant-rate sines
// moved back analytically, the carrier's accumulator = (true start of the PREVIOUS segment) + (what that segment
// had added `seg - warm` samples in, from the summary pass's snapshot), the filter with zero history.
__global__ void split_fm_warm_seed_kernel(const tb_split_args A) {
    const uint32_t vv = blockIdx.x * blockDim.x + threadIdx.x;
    if (vv >= A.n_real * A.n_seg) return;
    const uint32_t v = vv / A.n_seg, s = vv - v * A.n_seg;
    const uint32_t* src = A.vi + (size_t)vv * A.state_words;
    uint32_t* dst = A.vw + (size_t)vv * A.state_words;
    for (uint32_t k = 0; k < A.state_words; k++) dst[k] = src[k];
    if (s == 0) return;  // the first segment starts from the voice's own state: its warm-up result is not used
    for (uint32_t k = 0; k < A.n_entries; k++) {
        const tb_split_entry e = A.entries[k];
        if (e.kind == SP_POS || e.kind == SP_SINE_CONST)
            stu64(dst, e.state_off, ldu64(dst, e.state_off) - A.inc[(size_t)v * A.n_entries + k] * A.warm);
    }
    const tb_split_entry ec = A.entries[A.fm_carrier], ef = A.entries[A.fm_filter];
    const u64 guess = ldu64(A.real_state + (size_t)v * A.state_words, ec.state_off);  // what every segment of the summary pass started from
    const u64 prev_start = ldu64(A.vi + (size_t)(vv - 1) * A.state_words, ec.state_off);
    // (the snapshot of the summary pass sits in the first two history words of the previous segment's final state)
    const u64 snap = ldu64(A.vs + (size_t)(vv - 1) * A.state_words, ef.state_off + 2);
    stu64(dst, ec.state_off, prev_start + (snap - guess));
    const tb_filter_tab* ft = &A.filt[ef.a];
    for (uint32_t k = 0; k < ft->K - 1 + ft->J; k++) dst[ef.state_off + 2 + k] = 0u;
}
// The warm-up's final filter history is the segment's initial one.
__global__ void split_fm_adopt_kernel(const tb_split_args A) {
    const uint32_t vv = blockIdx.x * blockDim.x + threadIdx.x;
    if (vv >= A.n_real * A.n_seg) return;
    if (vv % A.n_seg == 0) return;
    const tb_split_entry ef = A.entries[A.fm_filter];
    const tb_filter_tab* ft = &A.filt[ef.a];
    for (uint32_t k = 0; k < ft->K - 1 + ft->J; k++)
        A.vi[(size_t)vv * A.state_words + ef.state_off + 2 + k] = A.vw[(size_t)vv * A.state_words + ef.state_off + 2 + k];
}

// The voice's state is the final state of its last segment; out_len (+)= what the split rendered.
__global__ void split_finish_kernel(const tb_split_args A, uint32_t* real_state, unsigned long long* out_len,
                                    unsigned long long n, int accumulate) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= A.n_real) return;
    const uint32_t* src = A.vs + ((size_t)v * A.n_seg + (A.n_seg - 1u)) * A.state_words;
    uint32_t* dst = real_state + (size_t)v * A.state_words;
    for (uint32_t k = 0; k < A.state_words; k++) dst[k] = src[k];
    if (out_len) out_len[v] = (accumulate ? out_len[v] : 0ull) + n;
}

}  // namespace

// out_len[v] = (accumulate ? out_len[v] : 0) + add  (abi.cpp launch_sequence)
namespace {
__global__ void len_set_kernel(unsigned long long* out_len, uint32_t n, unsigned long long add, int accumulate) {
    const uint32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) out_len[v] = (accumulate ? out_len[v] : 0ull) + add;
}
}  // namespace
extern "C" cudaError_t tb_len_set(unsigned long long* out_len, uint32_t n, unsigned long long add, int accumulate,
                                  cudaStream_t stream) {
    len_set_kernel<<<(n + 127) / 128, 128, 0, stream>>>(out_len, n, add, accumulate);
    return cudaGetLastError();
}

extern "C" cudaError_t tb_split_seed(const tb_split_args* A, cudaStream_t stream) {
    split_prepare_kernel<<<(A->n_real + 127) / 128, 128, 0, stream>>>(*A);
    const uint32_t nv = A->n_real * A->n_seg;
    split_seed_kernel<<<(nv + 127) / 128, 128, 0, stream>>>(*A);
    return cudaGetLastError();
}
extern "C" cudaError_t tb_split_fix(const tb_split_args* A, uint32_t level, cudaStream_t stream) {
    split_fix_kernel<<<A->n_real, 32, 0, stream>>>(*A, level);
    return cudaGetLastError();
}
extern "C" cudaError_t tb_split_fm_need(const tb_split_args* A, cudaStream_t stream) {
    split_fm_need_kernel<<<(A->n_real + 127) / 128, 128, 0, stream>>>(*A);
    return cudaGetLastError();
}
extern "C" cudaError_t tb_split_fm_warm_seed(const tb_split_args* A, cudaStream_t stream) {
    const uint32_t nv = A->n_real * A->n_seg;
    split_fm_warm_seed_kernel<<<(nv + 127) / 128, 128, 0, stream>>>(*A);
    return cudaGetLastError();
}
extern "C" cudaError_t tb_split_fm_adopt(const tb_split_args* A, cudaStream_t stream) {
    const uint32_t nv = A->n_real * A->n_seg;
    split_fm_adopt_kernel<<<(nv + 127) / 128, 128, 0, stream>>>(*A);
    return cudaGetLastError();
}
extern "C" cudaError_t tb_split_finish(const tb_split_args* A, uint32_t* real_state, unsigned long long* out_len,
                                       unsigned long long n, int accumulate, cudaStream_t stream) {
    split_finish_kernel<<<(A->n_real + 127) / 128, 128, 0, stream>>>(*A, real_state, out_len, n, accumulate);
    return cudaGetLastError();
}
