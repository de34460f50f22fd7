// render.cu — sm_100a kernels of the B200-native Tuun renderer.
//
// One warp renders one voice.  Per pass it evaluates the whole lowered Waveform program
// (program.h) over a tile of 32 x C samples: lane l owns the C consecutive positions
// [l*C, l*C+C), the running result is C registers per lane, named temporaries live in
// per-warp shared-memory slots (conflict-free float4 layout), and every piece of sequential
// state of the reference generator becomes a warp-shuffle scan:
//   Sine phase            (generator.rs:206-219)  exclusive prefix sum of 64-bit fixed-point turns
//   IIR feedback          (generator.rs:500-507)  scan of affine maps (constant companion powers)
//   Reset restarts        (generator.rs:290-316)  "last sign event" + running-max scans -> origins
//   Fin / Append lengths  (generator.rs:133-188)  warp-uniform window arithmetic
// The carried state of every stateful node (generator.rs:12-35) is loaded once per launch
// into shared memory and written back at the end, so successive launches continue the stream.
// Output rows are written with float4 stores.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tuun_b200.h"
#include "program.h"

namespace {

#include "common.cuh"


// ------------------------------------------------------------------------------------------
// per-warp machine state
// ------------------------------------------------------------------------------------------
struct WarpMem {
    uint32_t voice;
    float* cval;
    u64* aux;
    uint32_t* state;
    int* slot_len;
    uint8_t* slot_vm;  // [n_slots][32] validity bits of segmented temporaries
    float* slots;
};

struct Ctx {
    int w0, w1;     // current window [w0, w1) in tile positions
    int L;          // length of the last evaluated node, counted from w0
    int seg_slot;   // slot holding per-sample run origins when evaluating under a Reset, else -1
    int o_last;     // origin of the run that is live at the end of the window
    uint32_t vm;    // validity bits of the accumulator (segmented mode)
    bool first_tile;
};

__device__ __forceinline__ u64 ld_state64(const uint32_t* st, int off) {
    return (u64)st[off] | ((u64)st[off + 1] << 32);
}
__device__ __forceinline__ void st_state64(uint32_t* st, int off, u64 v) {
    st[off] = (uint32_t)v;
    st[off + 1] = (uint32_t)(v >> 32);
}

#define APPLY_OP(OPV, DST, A, B)                                                     \
    switch (OPV) {                                                                   \
        case TB_ADD:                                                                 \
        case TB_MERGE: UNROLL for (int j = 0; j < C; j++) DST[j] = __fadd_rn(A, B); break; \
        case TB_SUBTRACT: UNROLL for (int j = 0; j < C; j++) DST[j] = __fsub_rn(A, B); break; \
        case TB_MULTIPLY: UNROLL for (int j = 0; j < C; j++) DST[j] = __fmul_rn(A, B); break; \
        case TB_DIVIDE:                                                              \
            UNROLL for (int j = 0; j < C; j++) {                                     \
                float bb_ = (B);                                                     \
                DST[j] = bb_ == 0.0f ? 0.0f : __fdiv_rn(A, bb_);                     \
            }                                                                        \
            break;                                                                   \
        default: UNROLL for (int j = 0; j < C; j++) DST[j] = powf(A, B); break;      \
    }

// greater_or_equals_at(length, 0.0, max)  (generator.rs:787-862) over the flattened chain.
// kind: 0 = Some(len), 1 = None, 2 = Maybe.
__device__ void goe_eval(const tb_launch& P, const WarpMem& M, const Ctx& cx, int gi, int n,
                         int* kind, int* len) {
    const tb_goe g = P.goe[gi];
    float value = 0.0f;
    for (uint32_t s = 0; s < g.n_steps; s++) {
        const int sign = P.goe_steps[g.step_off + 2 * s];
        const float c = M.cval[P.goe_steps[g.step_off + 2 * s + 1]];
        value = sign > 0 ? __fadd_rn(value, c) : __fsub_rn(value, c);
    }
    int k = 2, l = 0;
    if (g.term == GOE_TIME) {
        const u64 pos = ld_state64(M.state, g.term_arg);
        const float srf = (float)P.sample_rate;
        const u64 target = f32_as_usize(ceilf(__fmul_rn(value, srf)));
        // The reference compares pos/sr >= value once per generate call (:808-809); inside a
        // call (later tiles) the cut stays where that call put it, i.e. at `target`.
        const bool reached = cx.first_tile ? (__fdiv_rn((float)pos, srf) >= value) : (pos >= target);
        k = 0;
        if (reached) l = 0;
        else {
            u64 rem = target > pos ? target - pos : 0ull;
            l = rem < (u64)n ? (int)rem : n;
        }
    } else if (g.term == GOE_CONST) {
        k = M.cval[g.term_arg] >= value ? 0 : 1;
        l = 0;
    }
    if (k == 1 && g.through_append) k = 2;
    *kind = k;
    *len = l;
}

// ------------------------------------------------------------------------------------------
// Sine kernels (window mode)
// ------------------------------------------------------------------------------------------
// Frequencies in `f` (valid for positions [w0, w0+f_len)), phase offsets already in fixed point.
template <bool UNIFORM_PH>
__device__ __forceinline__ void sine_window(float (&acc)[C], const float (&f)[C], const u64 (&phfx)[C], u64 ph0,
                                            int w0, int f_len, uint32_t* state, int st,
                                            const SineK& sk, int fast) {
    const int l = lane_id();
    const u64 acc0 = ld_state64(state, st);
    u64 inc[C];
    freq_to_inc_vec(inc, f, sk);
    if (!(w0 == 0 && f_len == TILE)) {  // partial window: samples outside it do not advance the phase
        UNROLL for (int j = 0; j < C; j++) {
            const int i = l * C + j;
            inc[j] = (i >= w0 && i < w0 + f_len) ? inc[j] : 0ull;
        }
    }
    u64 ph[C];
    u64 run = 0;
    UNROLL for (int j = 0; j < C; j++) {
        ph[j] = UNIFORM_PH ? run : run + phfx[j];
        run += inc[j];
    }
    const u64 incl = warp_incl_sum(run);
    const u64 base = acc0 + (incl - run) + (UNIFORM_PH ? ph0 : 0ull);
    const u64 total = __shfl_sync(FULL, incl, 31);
    UNROLL for (int j = 0; j < C; j++) ph[j] += base;
    sin_turns_vec(acc, ph, fast);
    __syncwarp();
    st_state64(state, st, acc0 + total);
}
// Constant frequency: phase(i) = acc0 + (i - w0) * inc, no scan.
template <bool UNIFORM_PH>
__device__ __forceinline__ void sine_window_cf(float (&acc)[C], u64 inc, const u64 (&phfx)[C], u64 ph0, int w0,
                                               int f_len, uint32_t* state, int st, int fast) {
    const int l = lane_id();
    const u64 acc0 = ld_state64(state, st);
    u64 ph[C];
    u64 p = acc0 + inc * (u64)(i64)(l * C - w0) + (UNIFORM_PH ? ph0 : 0ull);
    UNROLL for (int j = 0; j < C; j++) {
        ph[j] = UNIFORM_PH ? p : p + phfx[j];
        p += inc;
    }
    sin_turns_vec(acc, ph, fast);
    __syncwarp();
    st_state64(state, st, acc0 + inc * (u64)f_len);
}

// ------------------------------------------------------------------------------------------
// Filter (generator.rs:223-258, 382-515)
// State block at `st`: [0] initialised flag, [1] h = input.len(), [2..2+K-1) input deque
// (oldest first), then J outputs (oldest first).
// ------------------------------------------------------------------------------------------
template <int J>
__device__ __forceinline__ void iir_scan_const(float (&acc)[C], const float (&u)[C], const float* a,
                                               const double* mpow, const float* hy, int w0,
                                               int out_len) {
    // Feedback part of the filter for constant coefficients: y[m] = u[m] - sum_j a_j y[m-1-j].
    // Pass 1: every lane runs its chunk from a zero state (the lane holding w0 from the carried
    // state).  A warp scan with the constant companion-matrix powers A^(C*2^k) turns the chunk
    // responses into the true state at each chunk start.  Pass 2 re-runs the chunk from that
    // state with the reference's exact f32 operation order (generator.rs:500-502).
    const int l = lane_id();
    const int l0 = w0 >> 3;
    float s[J > 0 ? J : 1];
    UNROLL for (int jj = 0; jj < J; jj++) s[jj] = (l == l0) ? hy[J - 1 - jj] : 0.0f;
    UNROLL for (int j = 0; j < C; j++) {
        const int i = l * C + j;
        if (i >= w0 && i < w0 + out_len) {
            float y = u[j];
            UNROLL for (int jj = 0; jj < J; jj++) y = fmaf(-a[jj], s[jj], y);
            UNROLL for (int jj = J - 1; jj > 0; jj--) s[jj] = s[jj - 1];
            s[0] = y;
        }
    }
    double v[J > 0 ? J : 1];
    UNROLL for (int jj = 0; jj < J; jj++) v[jj] = (double)s[jj];
    UNROLL for (int k = 0; k < 5; k++) {
        const int d = 1 << k;
        double t[J > 0 ? J : 1];
        UNROLL for (int jj = 0; jj < J; jj++) t[jj] = __shfl_up_sync(FULL, v[jj], d);
        if (l >= d) {
            UNROLL for (int r = 0; r < J; r++) {
                double accv = v[r];
                UNROLL for (int c = 0; c < J; c++) accv = fma(mpow[(k * J + r) * J + c], t[c], accv);
                v[r] = accv;
            }
        }
    }
    UNROLL for (int jj = 0; jj < J; jj++) {
        double up = __shfl_up_sync(FULL, v[jj], 1);
        s[jj] = (l == l0) ? hy[J - 1 - jj] : (float)up;
    }
    UNROLL for (int j = 0; j < C; j++) {
        const int i = l * C + j;
        float y = u[j];
        if (i >= w0 && i < w0 + out_len) {
            UNROLL for (int jj = 0; jj < J; jj++) y = __fsub_rn(y, __fmul_rn(a[jj], s[jj]));
            UNROLL for (int jj = J - 1; jj > 0; jj--) s[jj] = s[jj - 1];
            s[0] = y;
        }
        acc[j] = y;
    }
}

// The same feedback with no scan: lane after lane, each running the reference's recurrence
// (generator.rs:500-502) over its chunk from the state the previous lane left.  32 dependent steps
// per tile — used for the few general tiles that bracket a lane-per-voice render (tb_launch::
// exact_fb), so that the whole call carries the reference's own f32 recurrence.
template <int J>
__device__ __forceinline__ void iir_serial_exact(float (&acc)[C], const float (&u)[C], const float* a,
                                                 const float* hy, int w0, int out_len) {
    const int l = lane_id();
    float s[J > 0 ? J : 1];
    UNROLL for (int jj = 0; jj < J; jj++) s[jj] = hy[J - 1 - jj];
    UNROLL for (int j = 0; j < C; j++) acc[j] = u[j];
    _Pragma("unroll 1") for (int L = 0; L < 32; L++) {
        if (l == L) {
            UNROLL for (int j = 0; j < C; j++) {
                const int i = l * C + j;
                if (i >= w0 && i < w0 + out_len) {
                    float y = u[j];
                    UNROLL for (int jj = 0; jj < J; jj++) y = __fsub_rn(y, __fmul_rn(a[jj], s[jj]));
                    UNROLL for (int jj = J - 1; jj > 0; jj--) s[jj] = s[jj - 1];
                    s[0] = y;
                    acc[j] = y;
                }
            }
        }
        UNROLL for (int jj = 0; jj < J; jj++) s[jj] = __shfl_sync(FULL, s[jj], L);
    }
}

// Feedback with coefficient WAVEFORMS (generator.rs:467-507: one value per output sample; filter_1_1_linear of
// benches/tracker_benches.rs:36-67): the same lane-after-lane recurrence, every lane holding its own samples of
// every coefficient in registers.  (The general form below — one lane walking the tile through shared memory —
// remains for more than four feedback taps.)
template <int J>
__device__ __forceinline__ void iir_serial_varying(float (&acc)[C], const float (&u)[C], const float (&a)[J][C],
                                                   const float* hy, int w0, int out_len) {
    const int l = lane_id();
    float s[J];
    UNROLL for (int jj = 0; jj < J; jj++) s[jj] = hy[J - 1 - jj];
    UNROLL for (int j = 0; j < C; j++) acc[j] = u[j];
    _Pragma("unroll 1") for (int L = 0; L < 32; L++) {
        if (l == L) {
            UNROLL for (int j = 0; j < C; j++) {
                const int i = l * C + j;
                if (i >= w0 && i < w0 + out_len) {
                    float y = u[j];
                    UNROLL for (int jj = 0; jj < J; jj++) y = __fsub_rn(y, __fmul_rn(a[jj][j], s[jj]));
                    UNROLL for (int jj = J - 1; jj > 0; jj--) s[jj] = s[jj - 1];
                    s[0] = y;
                    acc[j] = y;
                }
            }
        }
        UNROLL for (int jj = 0; jj < J; jj++) s[jj] = __shfl_sync(FULL, s[jj], L);
    }
}
template <int J>
__device__ __forceinline__ void iir_varying(const WarpMem& M, const tb_filter_tab* ft, float (&acc)[C], const float (&u)[C],
                                            const float* hy, int w0, int out_len) {
    float a[J][C];
    UNROLL for (int jj = 0; jj < J; jj++) {
        const int oj = ft->coef[ft->K + jj];
        if (oj < 0) {
            const float c = M.cval[~oj];
            UNROLL for (int j = 0; j < C; j++) a[jj][j] = c;
        } else {
            slot_load(M.slots, oj, a[jj]);
        }
    }
    iir_serial_varying<J>(acc, u, a, hy, w0, out_len);
}

// Full-tile fast path of the constant-coefficient filter: window = whole tile, history complete,
// nothing finishing.  Same arithmetic as filter_run below without any per-sample window test.
template <int J>
__device__ __forceinline__ void filter_full_tile(const WarpMem& M, const tb_filter_tab* ft, float (&acc)[C],
                                                 float* hx, float* hy) {
    const int l = lane_id();
    const int K = ft->K;
    float u[C];
    {
        const float b0 = M.cval[~ft->coef[0]];
        UNROLL for (int j = 0; j < C; j++) u[j] = __fmul_rn(acc[j], b0);
    }
    UNROLL for (int k = 1; k < TB_MAX_K; k++) {
        if (k < K) {
            const float bk = M.cval[~ft->coef[k]];
            float pv[C];
            UNROLL for (int j = 0; j < C; j++) {
                if (j < k) {
                    float t = __shfl_up_sync(FULL, acc[C - k + j], 1);
                    if (l == 0) t = hx[K - 1 - k + j];
                    pv[j] = t;
                }
            }
            UNROLL for (int j = 0; j < C; j++) u[j] = __fadd_rn(u[j], __fmul_rn(bk, (j >= k) ? acc[j - k] : pv[j]));
        }
    }
    float xin[C];
    UNROLL for (int j = 0; j < C; j++) xin[j] = acc[j];
    if (J > 0) {
        float a[J > 0 ? J : 1];
        UNROLL for (int jj = 0; jj < J; jj++) a[jj] = M.cval[~ft->coef[K + jj]];
        const double* mpow = reinterpret_cast<const double*>(M.aux + ft->pow_aux);
        float s[J > 0 ? J : 1];
        UNROLL for (int jj = 0; jj < J; jj++) s[jj] = (l == 0) ? hy[J - 1 - jj] : 0.0f;
        UNROLL for (int j = 0; j < C; j++) {  // pass 1: chunk response (zero state except lane 0)
            float y = u[j];
            UNROLL for (int jj = 0; jj < J; jj++) y = fmaf(-a[jj], s[jj], y);
            UNROLL for (int jj = J - 1; jj > 0; jj--) s[jj] = s[jj - 1];
            s[0] = y;
        }
        double v[J > 0 ? J : 1];
        UNROLL for (int jj = 0; jj < J; jj++) v[jj] = (double)s[jj];
        UNROLL for (int k = 0; k < 5; k++) {  // scan with A^(8*2^k)
            const int d = 1 << k;
            double t[J > 0 ? J : 1];
            UNROLL for (int jj = 0; jj < J; jj++) t[jj] = __shfl_up_sync(FULL, v[jj], d);
            if (l >= d) {
                UNROLL for (int r = 0; r < J; r++) {
                    double accv = v[r];
                    UNROLL for (int c = 0; c < J; c++) accv = fma(mpow[(k * J + r) * J + c], t[c], accv);
                    v[r] = accv;
                }
            }
        }
        UNROLL for (int jj = 0; jj < J; jj++) {
            const double up = __shfl_up_sync(FULL, v[jj], 1);
            s[jj] = (l == 0) ? hy[J - 1 - jj] : (float)up;
        }
        UNROLL for (int j = 0; j < C; j++) {  // pass 2: the reference's exact f32 order (generator.rs:500-502)
            float y = u[j];
            UNROLL for (int jj = 0; jj < J; jj++) y = __fsub_rn(y, __fmul_rn(a[jj], s[jj]));
            UNROLL for (int jj = J - 1; jj > 0; jj--) s[jj] = s[jj - 1];
            s[0] = y;
            acc[j] = y;
        }
    } else {
        UNROLL for (int j = 0; j < C; j++) acc[j] = u[j];
    }
    __syncwarp();
    if (l == 31) {  // the deques keep the last K-1 inputs and J outputs of the tile
        UNROLL for (int j = 0; j < C; j++) {
            const int ex = j - (C - (K - 1));
            if (ex >= 0) hx[ex] = xin[j];
            const int ey = j - (C - J);
            if (ey >= 0) hy[ey] = acc[j];
        }
    }
}

__device__ void filter_run(const tb_launch& P, const WarpMem& M, Ctx& cx, float (&acc)[C], int st,
                           int fi, bool had_begin) {
    const tb_filter_tab* ft = &P.filt[fi];
    const int K = ft->K, J = ft->J;
    const int l = lane_id();
    uint32_t* S = M.state + st;
    float* hx = reinterpret_cast<float*>(S + 2);
    float* hy = hx + (K - 1);
    int n, inner_len, out_len, h;
    int w0 = cx.w0;
    if (!had_begin && K <= TB_MAX_K && S[0] != 0u && cx.w0 == 0 && cx.w1 == TILE && cx.L == TILE && (int)S[1] == K - 1 &&
        (J == 0 || (ft->fb_const && !P.exact_fb))) {
        switch (J) {
            case 0: filter_full_tile<0>(M, ft, acc, hx, hy); break;
            case 1: filter_full_tile<1>(M, ft, acc, hx, hy); break;
            case 2: filter_full_tile<2>(M, ft, acc, hx, hy); break;
            case 3: filter_full_tile<3>(M, ft, acc, hx, hy); break;
            default: filter_full_tile<4>(M, ft, acc, hx, hy); break;
        }
        cx.L = TILE;
        return;
    }
    if (had_begin) {
        // G_FILT_BEGIN already narrowed the window to out_len and stored the zero-extended input.
        out_len = cx.w1 - cx.w0;
        n = (int)S[2 + (K - 1) + J];      // scratch words written by G_FILT_BEGIN
        inner_len = (int)S[3 + (K - 1) + J];
        h = (int)S[1];
        slot_load(M.slots, ft->x_slot, acc);
    } else {
        if (S[0] == 0u) {  // K == 1: nothing to pre-read (generator.rs:234-243)
            __syncwarp();
            S[0] = 1u;
            S[1] = 0u;
            for (int e = l; e < J; e += 32) hy[e] = 0.0f;
            __syncwarp();
        }
        n = cx.w1 - cx.w0;
        inner_len = cx.L;
        h = (int)S[1];
        out_len = min(n, inner_len + h);
        UNROLL for (int j = 0; j < C; j++) {
            const int i = l * C + j;
            if (i >= w0 + inner_len) acc[j] = 0.0f;  // zero-extend (generator.rs:404-405)
        }
    }
    const int pad = (h == K - 1) ? 0 : (K - 1) - h;  // generator.rs:408-418
    // History as the deque would read after padding: hx[0..h) then zeros.
    auto hist_at = [&](int idx) -> float { return (idx >= 0 && idx < h) ? hx[idx] : 0.0f; };

    // xe: the input with the history injected at positions just before the window.
    float xe[C];
    UNROLL for (int j = 0; j < C; j++) {
        const int i = l * C + j;
        xe[j] = (i >= w0) ? acc[j] : hist_at(K - 1 + (i - w0));
    }
    // Feed-forward: u = x*b0 + b1*x[-1] + ... each product and sum rounded (generator.rs:496-499).
    float u[C];
    {
        float bk[C];
        const int o0 = ft->coef[0];
        if (o0 < 0) { const float c = M.cval[~o0]; UNROLL for (int j = 0; j < C; j++) bk[j] = c; }
        else slot_load(M.slots, o0, bk);
        UNROLL for (int j = 0; j < C; j++) u[j] = __fmul_rn(xe[j], bk[j]);
    }
    UNROLL for (int k = 1; k < TB_MAX_K; k++) {
        if (k < K) {
            float bk[C];
            const int ok = ft->coef[k];
            if (ok < 0) { const float c = M.cval[~ok]; UNROLL for (int j = 0; j < C; j++) bk[j] = c; }
            else slot_load(M.slots, ok, bk);
            // value k positions back: own registers, or the previous lane's tail.
            float pv[C];  // pv[j] for j < k
            UNROLL for (int j = 0; j < C; j++) {
                if (j < k) {
                    float t = __shfl_up_sync(FULL, xe[C - k + j], 1);
                    if (l == 0) t = hist_at(K - 1 + (j - k - w0));
                    pv[j] = t;
                }
            }
            UNROLL for (int j = 0; j < C; j++) {
                const float xv = (j >= k) ? xe[j - k] : pv[j];
                u[j] = __fadd_rn(u[j], __fmul_rn(bk[j], xv));
            }
        }
    }
    if (K > TB_MAX_K) {  // long FIR (constant taps): taps 9.. read the staged input tile, in the reference's order
        const float* xs = M.slots + (size_t)ft->u_slot * TILE;
        __syncwarp();
        slot_store(M.slots, ft->u_slot, xe);
        __syncwarp();
        for (int k = TB_MAX_K; k < K; k++) {
            const float bk = M.cval[~ft->coef[k]];
            UNROLL for (int j = 0; j < C; j++) {
                const int idx = l * C + j - k;
                const float xv = idx >= 0 ? xs[slot_index(idx)] : hist_at(K - 1 + (idx - w0));
                u[j] = __fadd_rn(u[j], __fmul_rn(bk, xv));
            }
        }
        __syncwarp();
    }
    // Feedback.
    if (J == 0) {
        UNROLL for (int j = 0; j < C; j++) acc[j] = u[j];
    } else if (ft->fb_const) {
        float a[TB_MAX_J];
        UNROLL for (int jj = 0; jj < TB_MAX_J; jj++) a[jj] = jj < J ? M.cval[~ft->coef[K + jj]] : 0.0f;
        const double* mp = reinterpret_cast<const double*>(M.aux + ft->pow_aux);
        if (P.exact_fb) {
            switch (J) {
                case 1: iir_serial_exact<1>(acc, u, a, hy, w0, out_len); break;
                case 2: iir_serial_exact<2>(acc, u, a, hy, w0, out_len); break;
                case 3: iir_serial_exact<3>(acc, u, a, hy, w0, out_len); break;
                default: iir_serial_exact<4>(acc, u, a, hy, w0, out_len); break;
            }
        } else
        switch (J) {
            case 1: iir_scan_const<1>(acc, u, a, mp, hy, w0, out_len); break;
            case 2: iir_scan_const<2>(acc, u, a, mp, hy, w0, out_len); break;
            case 3: iir_scan_const<3>(acc, u, a, mp, hy, w0, out_len); break;
            default: iir_scan_const<4>(acc, u, a, mp, hy, w0, out_len); break;
        }
    } else if (J <= TB_MAX_J) {
        // Time-varying feedback coefficients: the recurrence lane after lane, coefficients in registers.
        switch (J) {
            case 1: iir_varying<1>(M, ft, acc, u, hy, w0, out_len); break;
            case 2: iir_varying<2>(M, ft, acc, u, hy, w0, out_len); break;
            case 3: iir_varying<3>(M, ft, acc, u, hy, w0, out_len); break;
            default: iir_varying<4>(M, ft, acc, u, hy, w0, out_len); break;
        }
    } else {
        // ... with more than four taps: serial recurrence by lane 0 over the tile, through shared memory.
        float* us = M.slots + (size_t)ft->u_slot * TILE;
        slot_store(M.slots, ft->u_slot, u);
        __syncwarp();
        if (l == 0) {
            float yh[TB_MAX_J_GEN];
            for (int jj = 0; jj < J; jj++) yh[jj] = hy[J - 1 - jj];
            for (int m = 0; m < out_len; m++) {
                const int si = slot_index(w0 + m);
                float y = us[si];
                for (int jj = 0; jj < J; jj++) {
                    const int oj = ft->coef[K + jj];
                    const float aj = oj < 0 ? M.cval[~oj] : M.slots[(size_t)oj * TILE + si];
                    y = __fsub_rn(y, __fmul_rn(aj, yh[jj]));
                }
                for (int jj = J - 1; jj > 0; jj--) yh[jj] = yh[jj - 1];
                yh[0] = y;
                us[si] = y;
            }
        }
        __syncwarp();
        slot_load(M.slots, ft->u_slot, acc);
    }
    // New deques: last K-1 of (history ++ x[0..out_len)), last J of (outputs ++ y[0..out_len)).
    float keep_x = 0.0f, keep_y = 0.0f;
    if (l < K - 1 && out_len + l < K - 1) keep_x = hist_at(out_len + l);
    if (l < J && out_len + l < J) keep_y = hy[out_len + l];
    __syncwarp();
    if (l < K - 1 && out_len + l < K - 1) hx[l] = keep_x;
    if (l < J && out_len + l < J) hy[l] = keep_y;
    UNROLL for (int j = 0; j < C; j++) {
        const int m = l * C + j - w0;  // sample index inside the window
        if (m >= 0 && m < out_len) {
            const int ex = m - (out_len - (K - 1));
            if (ex >= 0) hx[ex] = xe[j];
            const int ey = m - (out_len - J);
            if (ey >= 0) hy[ey] = acc[j];
        }
    }
    // truncate(len - (padding + extra)): wraps (no-op) when it would underflow (appendix A7).
    const int drop = pad + (n - inner_len);
    __syncwarp();
    if (l == 0) S[1] = (uint32_t)(drop <= K - 1 ? (K - 1) - drop : K - 1);
    cx.L = out_len;
}

// ------------------------------------------------------------------------------------------
// Reset: restart positions and per-sample run origins (generator.rs:281-318)
// ------------------------------------------------------------------------------------------
// acc holds the trigger for positions [w0, w0 + t_len).  Writes, into `org_slot`, for every
// position the tile position at which its run started (-1: the run continues from the previous
// tile) and returns the origin that is live at the end of the window.
__device__ int reset_origins(const WarpMem& M, const float (&trig)[C], uint32_t trig_vm, int w0,
                             int t_len, int st, int org_slot, int outer_slot) {
    const int l = lane_id();
    float oo[C];
    if (outer_slot >= 0) slot_load(M.slots, outer_slot, oo);
    int outer[C];
    UNROLL for (int j = 0; j < C; j++) outer[j] = outer_slot >= 0 ? __float_as_int(oo[j]) : -1;
    const bool neg_carried = M.state[st] == 0u;  // state word: 1 = last class non-negative
    // Event per sample: 0 keep, 1 set non-negative, 2 set negative.  A run start of the enclosing
    // Reset re-initialises this Reset (signum = -1) before the sample is looked at.
    bool any_set = false, last_neg = false;
    UNROLL for (int j = 0; j < C; j++) {
        const int i = l * C + j;
        const bool on = i >= w0 && i < w0 + t_len && ((trig_vm >> j) & 1u);
        if (outer[j] == i) { any_set = true; last_neg = true; }
        if (on) {
            const float x = trig[j];
            if (x < 0.0f) { any_set = true; last_neg = true; }
            else if (x >= 0.0f && !(x == 0.0f && signbit(x))) { any_set = true; last_neg = false; }
        }
    }
    const unsigned dmask = __ballot_sync(FULL, any_set);
    const unsigned nmask = __ballot_sync(FULL, last_neg);
    const unsigned below = dmask & ((1u << l) - 1u);
    bool neg = below ? ((nmask >> (31 - __clz(below))) & 1u) : neg_carried;
    int last_restart = -1;
    bool restart[C];
    UNROLL for (int j = 0; j < C; j++) {
        const int i = l * C + j;
        const bool on = i >= w0 && i < w0 + t_len && ((trig_vm >> j) & 1u);
        restart[j] = false;
        if (outer[j] == i) neg = true;
        if (on) {
            const float x = trig[j];
            if (neg && x >= 0.0f) {  // generator.rs:296-300
                restart[j] = true;
                last_restart = i;
                neg = signbit(x);
            } else if (!neg && x < 0.0f) {
                neg = true;
            }
        }
    }
    const unsigned rmask = __ballot_sync(FULL, last_restart >= 0);
    const unsigned rbelow = rmask & ((1u << l) - 1u);
    const int src = rbelow ? 31 - __clz(rbelow) : 0;
    const int cand = __shfl_sync(FULL, last_restart, src);
    int org = rbelow ? cand : -1;
    float ov[C];
    UNROLL for (int j = 0; j < C; j++) {
        const int i = l * C + j;
        if (restart[j]) org = i;
        const int o = max(org, outer[j]);
        ov[j] = __int_as_float(o);
    }
    slot_store(M.slots, org_slot, ov);
    const bool neg_end = __shfl_sync(FULL, (int)neg, 31) != 0;
    const int last_pos = w0 + t_len - 1;
    int mine = -1;
    UNROLL for (int j = 0; j < C; j++) if (l * C + j == last_pos) mine = __float_as_int(ov[j]);
    const int o_last = __shfl_sync(FULL, mine, (last_pos >> 3) & 31);
    __syncwarp();
    if (l == 0) M.state[st] = neg_end ? 0u : 1u;
    return t_len > 0 ? o_last : -1;
}

// ------------------------------------------------------------------------------------------
// The interpreter
// ------------------------------------------------------------------------------------------
__device__ void run_program(const tb_launch& P, const tb_insn* code, const WarpMem& M, Ctx& cx,
                            float (&acc)[C], int pc, const SineK& sk, int* ctl) {
    const int l = lane_id();
    const float srf = (float)P.sample_rate;
    int sp = 0;
#define PUSH(v) (ctl[sp++] = (v))
#define POP() (ctl[--sp])
    for (;;) {
        __syncwarp();
        const tb_insn in = code[pc++];
        const uint32_t op = in.op & 0xffu;
        const int fast = ((in.op >> 8) & 0xffu) == TB_SINE_FAST ? (int)P.fast_mode : 0;
        const int n = cx.w1 - cx.w0;
        switch (op) {
            case OP_END: return;

            // ------------------------------ generate ------------------------------
            case G_CONST: {  // generator.rs:97-100
                const float c = M.cval[in.a];
                UNROLL for (int j = 0; j < C; j++) acc[j] = c;
                cx.L = n;
                break;
            }
            case G_TIME: {  // generator.rs:101-111
                const u64 pos = ld_state64(M.state, in.a);
                UNROLL for (int j = 0; j < C; j++) {
                    const int i = l * C + j;
                    acc[j] = __fdiv_rn(__ull2float_rn(pos + (u64)(i - cx.w0)), srf);
                }
                __syncwarp();
                st_state64(M.state, in.a, pos + (u64)n);
                cx.L = n;
                break;
            }
            case G_FIXED: {  // generator.rs:119-131
                const u64 pos = ld_state64(M.state, in.a);
                const tb_fixed_tab ft = P.fixed[in.b];
                int len = 0;
                if (pos < ft.len) {
                    const u64 rem = ft.len - pos;
                    len = rem < (u64)n ? (int)rem : n;
                    const float* src = P.pool + ft.off + pos;
                    UNROLL for (int j = 0; j < C; j++) {
                        const int m = l * C + j - cx.w0;
                        acc[j] = (m >= 0 && m < len) ? __ldg(src + m) : 0.0f;
                    }
                }
                __syncwarp();
                st_state64(M.state, in.a, pos + (u64)len);
                cx.L = len;
                break;
            }
            case G_NOISE:
            case S_NOISE: {  // generator.rs:113-118; not restarted by an enclosing Reset (no tree state)
                const u64 pos = ld_state64(M.state, in.a);
                const u64 stream = noise_stream(P, M.voice, in.b);
                UNROLL for (int j = 0; j < C; j++) acc[j] = noise_at(stream, pos + (u64)(i64)(l * C + j - cx.w0));
                __syncwarp();
                st_state64(M.state, in.a, pos + (u64)n);
                if (op == G_NOISE) cx.L = n;
                else cx.vm = 0xffu;
                break;
            }
            case G_BINC: {  // generator.rs:538-549 (constant right-hand side)
                const float c = M.cval[in.b];
                if (in.c) {  // Merge: zero-extend to the window
                    if (cx.L == 0) {  // generator.rs:531-535
                        UNROLL for (int j = 0; j < C; j++) acc[j] = c;
                    } else {
                        UNROLL for (int j = 0; j < C; j++) {
                            const int i = l * C + j;
                            acc[j] = __fadd_rn(i < cx.w0 + cx.L ? acc[j] : 0.0f, c);
                        }
                    }
                    cx.L = n;
                } else {
                    APPLY_OP((uint32_t)in.a, acc, acc[j], c)
                }
                break;
            }
            case G_BIN_BEGIN: {  // generator.rs:530-553
                slot_store(M.slots, in.a, acc);
                M.slot_len[in.a] = cx.L;
                PUSH(cx.w1);
                if (!in.b) {
                    cx.w1 = cx.w0 + cx.L;
                    if (cx.L == 0) pc = in.c;
                }
                break;
            }
            case G_BIN_END: {  // generator.rs:555-567
                cx.w1 = POP();
                const int La = M.slot_len[in.a], Lb = cx.L;
                float av[C];
                slot_load(M.slots, in.a, av);
                if (!in.c) {
                    APPLY_OP((uint32_t)in.b, acc, av[j], acc[j])
                    cx.L = Lb;
                } else if (La != 0) {
                    UNROLL for (int j = 0; j < C; j++) {
                        const int i = l * C + j;
                        acc[j] = __fadd_rn(i < cx.w0 + La ? av[j] : 0.0f, i < cx.w0 + Lb ? acc[j] : 0.0f);
                    }
                    cx.L = max(La, Lb);
                }
                break;
            }
            case G_SINE_CC: {  // constant frequency and phase
                u64 ph[C];
                sine_window_cf<true>(acc, M.aux[in.b], ph, M.aux[in.c], cx.w0, n, M.state, in.a, fast);
                cx.L = n;
                break;
            }
            case G_SINE_AC: {  // frequency in acc (length L), constant phase
                u64 ph[C];
                float f[C];
                UNROLL for (int j = 0; j < C; j++) f[j] = acc[j];
                sine_window<true>(acc, f, ph, M.aux[in.c], cx.w0, cx.L, M.state, in.a, sk, fast);
                break;
            }
            case G_SINE_CA: {  // constant frequency, phase in acc (length L, zeros beyond)
                u64 ph[C];
                phase_to_fx_vec(ph, acc, sk);
                UNROLL for (int j = 0; j < C; j++) {
                    const int i = l * C + j;
                    ph[j] = i < cx.w0 + cx.L ? ph[j] : 0ull;
                }
                sine_window_cf<false>(acc, M.aux[in.b], ph, 0ull, cx.w0, n, M.state, in.a, fast);
                break;
            }
            case G_SINE_BEGIN: {  // generator.rs:206-210
                slot_store(M.slots, in.a, acc);
                M.slot_len[in.a] = cx.L;
                PUSH(cx.w1);
                cx.w1 = cx.w0 + cx.L;
                if (cx.L == 0) pc = in.c;
                break;
            }
            case G_SINE_END: {
                cx.w1 = POP();
                const int f_len = M.slot_len[in.b];
                float f[C];
                slot_load(M.slots, in.b, f);
                u64 ph[C];
                phase_to_fx_vec(ph, acc, sk);
                UNROLL for (int j = 0; j < C; j++) {
                    const int i = l * C + j;
                    ph[j] = i < cx.w0 + cx.L ? ph[j] : 0ull;
                }
                sine_window<false>(acc, f, ph, 0ull, cx.w0, f_len, M.state, in.a, sk, fast);
                break;
            }
            case G_ALT_CC: {  // generator.rs:335-341 with constant branches
                const float cp = M.cval[in.a], cn = M.cval[in.b];
                UNROLL for (int j = 0; j < C; j++) acc[j] = acc[j] >= 0.0f ? cp : cn;
                break;
            }
            case G_ALT_BEGIN: {
                slot_store(M.slots, in.a, acc);
                M.slot_len[in.a] = cx.L;
                PUSH(cx.w1);
                cx.w1 = cx.w0 + cx.L;
                if (cx.L == 0) pc = in.c;
                break;
            }
            case G_ALT_POS: {
                UNROLL for (int j = 0; j < C; j++) {
                    const int i = l * C + j;
                    if (i >= cx.w0 + cx.L) acc[j] = 0.0f;
                }
                slot_store(M.slots, in.a, acc);
                break;
            }
            case G_ALT_END: {
                cx.w1 = POP();
                const int t_len = M.slot_len[in.a];
                float t[C], pv[C];
                slot_load(M.slots, in.a, t);
                if (in.b >= 0) slot_load(M.slots, in.b, pv);
                else { const float c = M.cval[~in.b]; UNROLL for (int j = 0; j < C; j++) pv[j] = c; }
                if (in.c < 0) { const float c = M.cval[~in.c]; UNROLL for (int j = 0; j < C; j++) acc[j] = c; }
                else {
                    UNROLL for (int j = 0; j < C; j++) {
                        const int i = l * C + j;
                        if (i >= cx.w0 + cx.L) acc[j] = 0.0f;
                    }
                }
                UNROLL for (int j = 0; j < C; j++) acc[j] = t[j] >= 0.0f ? pv[j] : acc[j];
                cx.L = t_len;
                break;
            }
            case G_FILT_PRE: {  // generator.rs:223-252: first call reads K-1 inputs ahead
                if (M.state[in.a] != 0u || in.b <= 1) { pc = in.c; break; }
                PUSH(cx.w0);
                PUSH(cx.w1);
                cx.w0 = 0;
                cx.w1 = in.b - 1;
                break;
            }
            case G_FILT_PRE_END: {
                uint32_t* S = M.state + in.a;
                float* hx = reinterpret_cast<float*>(S + 2);
                float* hy = hx + (in.b - 1);
                UNROLL for (int j = 0; j < C; j++) if (l * C + j < cx.L) hx[l * C + j] = acc[j];
                if (l == 0) {
                    S[0] = 1u;
                    S[1] = (uint32_t)cx.L;
                }
                for (int e = l; e < in.c; e += 32) hy[e] = 0.0f;
                cx.w1 = POP();
                cx.w0 = POP();
                break;
            }
            case G_FILT_BEGIN: {
                const tb_filter_tab* ft = &P.filt[in.b];
                const int K = ft->K, J = ft->J;
                uint32_t* S = M.state + in.a;
                if (S[0] == 0u) {
                    __syncwarp();
                    if (l == 0) { S[0] = 1u; S[1] = 0u; }
                    for (int e = l; e < J; e += 32) reinterpret_cast<float*>(S + 2 + (K - 1))[e] = 0.0f;
                    __syncwarp();
                }
                const int inner_len = cx.L, h = (int)S[1];
                const int out_len = min(n, inner_len + h);
                UNROLL for (int j = 0; j < C; j++) {
                    const int i = l * C + j;
                    if (i >= cx.w0 + inner_len) acc[j] = 0.0f;
                }
                slot_store(M.slots, ft->x_slot, acc);
                if (l == 0) { S[2 + (K - 1) + J] = (uint32_t)n; S[3 + (K - 1) + J] = (uint32_t)inner_len; }
                PUSH(cx.w1);
                cx.w1 = cx.w0 + out_len;
                if (out_len == 0) pc = in.c;
                break;
            }
            case G_FILT_COEF: {  // coefficient waveforms read as 0 past their end (generator.rs:468-477)
                UNROLL for (int j = 0; j < C; j++) {
                    const int i = l * C + j;
                    if (i >= cx.w0 + cx.L) acc[j] = 0.0f;
                }
                slot_store(M.slots, in.a, acc);
                break;
            }
            case G_FILT_RUN: {
                filter_run(P, M, cx, acc, in.a, in.b, in.c != 0);
                if (in.c) cx.w1 = POP();
                break;
            }
            case G_FIN_HEAD: {  // also heads the length-mode Fin
                int kind, len;
                goe_eval(P, M, cx, in.a, n, &kind, &len);
                PUSH(kind);
                PUSH(len);
                if (kind != 2) pc = in.c;
                break;
            }
            case G_FIN_SCAN: {  // generator.rs:672-687 with the dummy inner: first v >= 0
                int cand = 0x7fffffff;
                UNROLL for (int j = C - 1; j >= 0; j--) {
                    const int m = l * C + j - cx.w0;
                    if (m >= 0 && m < cx.L && acc[j] >= 0.0f) cand = m;
                }
                cand = __reduce_min_sync(FULL, cand);
                const int len = min(min(cand, cx.L), n);
                sp -= 2;
                PUSH(len);
                pc = in.c;
                break;
            }
            case G_FIN_STATIC: {  // generator.rs:658-671 with the dummy inner
                const int len_g = POP();
                const int kind = POP();
                PUSH(kind == 0 ? min(len_g, n) : n);
                break;
            }
            case G_FIN_INNER: {  // generator.rs:164
                const int len = ctl[sp - 1];
                PUSH(cx.w1);
                cx.w1 = cx.w0 + len;
                if (len == 0) { cx.L = 0; pc = in.c; }
                break;
            }
            case G_FIN_ADV: {  // generator.rs:166: advance inner over the rest of the block
                const int inner_len = cx.L;
                cx.w1 = POP();
                const int len = POP();
                if (in.a && len > 0 && cx.w0 + len == cx.w1) {  // the Fin does not cut this block: `length(inner, 0)`
                    cx.L = inner_len;                          // would change nothing (lower.cpp has_unreached_filter)
                    pc = in.c;
                    break;
                }
                PUSH(inner_len);
                PUSH(cx.w0);
                cx.w0 = cx.w0 + len;
                break;
            }
            case G_FIN_END: {
                cx.w0 = POP();
                cx.L = POP();
                break;
            }
            case G_APP_BEGIN: {  // generator.rs:169-184
                const int fin = (int)M.state[in.a];
                PUSH(fin);
                if (fin) { cx.L = 0; pc = in.c; }
                break;
            }
            case G_APP_MID: {
                const int fin = POP();
                const int a_len = fin ? 0 : cx.L;
                if (!fin && a_len == n) { pc = in.c; break; }
                __syncwarp();
                if (l == 0) M.state[in.a] = 1u;
                PUSH(a_len);
                if (a_len > 0) slot_store(M.slots, in.b, acc);
                PUSH(cx.w0);
                cx.w0 += a_len;
                break;
            }
            case G_APP_END: {  // generator.rs:186-187
                const int w0n = cx.w0;
                cx.w0 = POP();
                const int a_len = POP();
                if (a_len > 0) {
                    float av[C];
                    slot_load(M.slots, in.b, av);
                    UNROLL for (int j = 0; j < C; j++) {
                        const int i = l * C + j;
                        if (i < w0n) acc[j] = av[j];
                    }
                }
                cx.L = a_len + cx.L;
                break;
            }
            case G_RESET_BEGIN: {
                const int t_len = cx.L;
                PUSH(cx.w1);
                PUSH(cx.seg_slot);
                PUSH(cx.o_last);
                PUSH(t_len);
                if (t_len == 0) { pc = in.c; break; }
                cx.w1 = cx.w0 + t_len;
                cx.o_last = reset_origins(M, acc, 0xffu, cx.w0, t_len, in.a, in.b, -1);
                cx.seg_slot = in.b;
                break;
            }
            case G_RESET_END: {  // zero-fill where the inner waveform ended early (generator.rs:309)
                const int t_len = POP();
                cx.o_last = POP();
                cx.seg_slot = POP();
                cx.w1 = POP();
                if (t_len > 0) {
                    UNROLL for (int j = 0; j < C; j++) if (!((cx.vm >> j) & 1u)) acc[j] = 0.0f;
                }
                cx.L = t_len;
                break;
            }
            case G_RUNS_BEGIN: {  // generator.rs:281-318, one run at a time: acc = trigger over [w0, w0 + t_len)
                const int t_len = cx.L;
                if (t_len == 0) {  // skip the inner tree, G_RUNS_END and its data words
                    pc = in.c + 1;
                    while (!code[pc++].c) {}
                    break;
                }
                PUSH(cx.w0);
                PUSH(cx.w1);
                PUSH(t_len);
                const int end = cx.w0 + t_len;
                (void)reset_origins(M, acc, 0xffu, cx.w0, t_len, in.a, in.b, -1);
                __syncwarp();
                float ov[C];
                slot_load(M.slots, in.b, ov);
                bool at0 = false;
                int nxt = end;
                UNROLL for (int j = 0; j < C; j++) {
                    const int i = l * C + j;
                    if (i >= cx.w0 && i < end && __float_as_int(ov[j]) == i) {
                        if (i == cx.w0) at0 = true;
                        else nxt = min(nxt, i);
                    }
                }
                nxt = __reduce_min_sync(FULL, nxt);
                if (__any_sync(FULL, at0)) {  // a restart on the first sample: the (empty) run before it renders nothing
                    for (int q = in.c + 1;; q++) {
                        const tb_insn z = code[q];
                        for (int k = l; k < z.b; k += 32) M.state[z.a + k] = 0u;
                        if (z.c) break;
                    }
                }
                cx.w1 = nxt;
                break;
            }
            case G_RUNS_END: {  // a run [w0, w1) of the inner tree is in acc, cx.L samples of it (zeros after, :309)
                const int t_len = ctl[sp - 1], w1s = ctl[sp - 2], w0s = ctl[sp - 3];
                const int end = w0s + t_len;
                const int rs = cx.w0, re = cx.w1, got = cx.L;
                float rv[C];
                if (rs > w0s) slot_load(M.slots, in.b, rv);
                UNROLL for (int j = 0; j < C; j++) {
                    const int i = l * C + j;
                    if (i >= rs && i < re) rv[j] = i < rs + got ? acc[j] : 0.0f;
                    else if (rs <= w0s) rv[j] = 0.0f;
                }
                if (re < end) {  // a restart at `re`: set_state(inner, Initial) (:311-313), then the next run
                    __syncwarp();
                    slot_store(M.slots, in.b, rv);
                    for (int q = pc;; q++) {
                        const tb_insn z = code[q];
                        for (int k = l; k < z.b; k += 32) M.state[z.a + k] = 0u;
                        if (z.c) break;
                    }
                    float ov[C];
                    slot_load(M.slots, in.a, ov);
                    int nxt = end;
                    UNROLL for (int j = 0; j < C; j++) {
                        const int i = l * C + j;
                        if (i > re && i < end && __float_as_int(ov[j]) == i) nxt = min(nxt, i);
                    }
                    nxt = __reduce_min_sync(FULL, nxt);
                    cx.w0 = re;
                    cx.w1 = nxt;
                    pc = in.c;
                } else {
                    UNROLL for (int j = 0; j < C; j++) acc[j] = rv[j];
                    sp -= 3;
                    cx.w0 = w0s;
                    cx.w1 = w1s;
                    cx.L = t_len;
                    while (!code[pc++].c) {}
                }
                break;
            }
            case G_SAVE: {
                slot_store(M.slots, in.a, acc);
                M.slot_len[in.a] = cx.L;
                break;
            }
            case G_RESTORE: {
                slot_load(M.slots, in.a, acc);
                break;
            }

            // ------------------------------ length ------------------------------
            case L_INF: cx.L = n; break;  // generator.rs:625,635
            case L_TIME: {                // generator.rs:626-633
                const u64 pos = ld_state64(M.state, in.a);
                __syncwarp();
                st_state64(M.state, in.a, pos + (u64)n);
                cx.L = n;
                break;
            }
            case L_FIXED: {  // generator.rs:636-647
                const u64 pos = ld_state64(M.state, in.a);
                const tb_fixed_tab ft = P.fixed[in.b];
                int len = 0;
                if (pos < ft.len) {
                    const u64 rem = ft.len - pos;
                    len = rem < (u64)n ? (int)rem : n;
                }
                __syncwarp();
                st_state64(M.state, in.a, pos + (u64)len);
                cx.L = len;
                break;
            }
            case L_PUSH: PUSH(cx.L); break;
            case L_MIN: { const int t = POP(); cx.L = min(t, cx.L); break; }
            case L_MAX: { const int t = POP(); cx.L = max(t, cx.L); break; }
            case L_POP: cx.L = POP(); break;
            case L_FILT_HEAD: {  // generator.rs:690-704
                uint32_t* S = M.state + in.a;
                const int was_init = (int)S[0];
                __syncwarp();
                if (!was_init) {
                    float* hx = reinterpret_cast<float*>(S + 2);
                    for (int e = l; e < (in.b - 1) + in.c; e += 32) hx[e] = 0.0f;
                    if (l == 0) { S[0] = 1u; S[1] = (uint32_t)(in.b - 1); }
                }
                PUSH(was_init);
                break;
            }
            case L_FILT_MID: {  // the Initial arm returns length(inner) without touching coefficients
                const int was_init = POP();
                PUSH(cx.L);
                if (!was_init) pc = in.c;
                break;
            }
            case L_FILT_END: cx.L = POP(); break;
            case L_APP_BEGIN: {  // generator.rs:725-738
                const int fin = (int)M.state[in.a];
                PUSH(fin);
                if (fin) { cx.L = 0; pc = in.c; }
                break;
            }
            case L_APP_MID: {
                const int fin = POP();
                const int a_len = fin ? 0 : cx.L;
                __syncwarp();
                if (!fin && a_len < n && l == 0) M.state[in.a] = 1u;
                PUSH(a_len);
                PUSH(cx.w0);
                cx.w0 += a_len;
                break;
            }
            case L_APP_END: {
                cx.w0 = POP();
                const int a_len = POP();
                cx.L = a_len + cx.L;
                break;
            }
            case L_FIN_SCAN1: {  // generator.rs:679-683: first index with i == length_len || v >= 0
                int cand = 0x7fffffff;
                UNROLL for (int j = C - 1; j >= 0; j--) {
                    const int m = l * C + j - cx.w0;
                    if (m >= 0 && m < cx.L && acc[j] >= 0.0f) cand = m;
                }
                cand = __reduce_min_sync(FULL, cand);
                sp -= 2;
                PUSH(min(cand, cx.L));
                break;
            }
            case L_FIN_SCAN2: {
                const int first = POP();
                cx.L = min(min(first, cx.L), n);
                pc = in.c;
                break;
            }
            case L_FIN_STATIC: {  // generator.rs:658-671
                const int inner_len = POP();
                const int len_g = POP();
                const int kind = POP();
                cx.L = kind == 0 ? min(len_g, inner_len) : inner_len;
                break;
            }

            // ------------------------------ segmented ------------------------------
            case S_CONST: {
                const float c = M.cval[in.a];
                UNROLL for (int j = 0; j < C; j++) acc[j] = c;
                cx.vm = 0xffu;
                break;
            }
            case S_TIME:
            case S_FIXED: {
                const u64 pos = ld_state64(M.state, in.a);
                float ov[C];
                slot_load(M.slots, cx.seg_slot, ov);
                tb_fixed_tab ft;
                const float* src = nullptr;
                if (op == S_FIXED) { ft = P.fixed[in.b]; src = P.pool + ft.off; }
                uint32_t vm = 0;
                UNROLL for (int j = 0; j < C; j++) {
                    const int i = l * C + j;
                    const int o = __float_as_int(ov[j]);
                    const u64 nloc = o >= 0 ? (u64)(i - o) : pos + (u64)(i - cx.w0);
                    if (op == S_TIME) {
                        acc[j] = __fdiv_rn(__ull2float_rn(nloc), srf);
                        vm |= 1u << j;
                    } else {
                        const bool ok = i >= cx.w0 && i < cx.w1 && nloc < ft.len;
                        acc[j] = ok ? __ldg(src + nloc) : 0.0f;
                        vm |= (ok ? 1u : 0u) << j;
                    }
                }
                __syncwarp();
                u64 np = cx.o_last >= 0 ? (u64)(cx.w1 - cx.o_last) : pos + (u64)n;
                if (op == S_FIXED && np > ft.len) np = ft.len;
                st_state64(M.state, in.a, np);
                cx.vm = vm;
                break;
            }
            case S_BINC: {
                const float c = M.cval[in.b];
                if (in.c) {
                    UNROLL for (int j = 0; j < C; j++) acc[j] = __fadd_rn(((cx.vm >> j) & 1u) ? acc[j] : 0.0f, c);
                    cx.vm = 0xffu;
                } else {
                    APPLY_OP((uint32_t)in.a, acc, acc[j], c)
                }
                break;
            }
            case S_BIN_BEGIN:
            case S_SINE_BEGIN:
            case S_ALT_BEGIN: {
                slot_store(M.slots, in.a, acc);
                M.slot_vm[in.a * 32 + l] = (uint8_t)cx.vm;
                break;
            }
            case S_BIN_END: {
                float av[C];
                slot_load(M.slots, in.a, av);
                const uint32_t va = M.slot_vm[in.a * 32 + l], vb = cx.vm;
                if (!in.c) {
                    APPLY_OP((uint32_t)in.b, acc, av[j], acc[j])
                    cx.vm = va & vb;
                } else {
                    UNROLL for (int j = 0; j < C; j++)
                        acc[j] = __fadd_rn(((va >> j) & 1u) ? av[j] : 0.0f, ((vb >> j) & 1u) ? acc[j] : 0.0f);
                    cx.vm = va | vb;
                }
                break;
            }
            case S_SINE_CC:
            case S_SINE_CA:
            case S_SINE_AC:
            case S_SINE_END: {
                // Segmented phase: restarts from 0 at every run origin (waveform.rs:341-349 via
                // set_state), continues from the carried accumulator where origin == -1.
                const bool f_const = op == S_SINE_CC || op == S_SINE_CA;
                const bool p_const = op == S_SINE_CC || op == S_SINE_AC;
                float ov[C];
                slot_load(M.slots, cx.seg_slot, ov);
                const u64 acc0 = ld_state64(M.state, in.a);
                u64 ph[C];
                uint32_t vm = 0xffu;
                if (p_const) {
                    const u64 p0 = M.aux[in.c];
                    UNROLL for (int j = 0; j < C; j++) ph[j] = p0;
                } else {
                    UNROLL for (int j = 0; j < C; j++) ph[j] = ((cx.vm >> j) & 1u) ? phase_to_fx(acc[j], sk) : 0ull;
                    vm = cx.vm;
                }
                u64 newacc;
                if (f_const) {
                    const u64 inc = M.aux[in.b];
                    UNROLL for (int j = 0; j < C; j++) {
                        const int i = l * C + j;
                        const int o = __float_as_int(ov[j]);
                        const u64 base = o >= 0 ? inc * (u64)(i - o) : acc0 + inc * (u64)(i - cx.w0);
                        acc[j] = sin_turns(base + ph[j], fast);
                    }
                    newacc = cx.o_last >= 0 ? inc * (u64)(cx.w1 - cx.o_last) : acc0 + inc * (u64)n;
                } else {
                    float f[C];
                    uint32_t fvm;
                    if (op == S_SINE_AC) {
                        UNROLL for (int j = 0; j < C; j++) f[j] = acc[j];
                        fvm = cx.vm;
                    } else {
                        slot_load(M.slots, in.b, f);
                        fvm = M.slot_vm[in.b * 32 + l];
                    }
                    vm &= fvm;
                    u64 ex[C];
                    u64 run = 0;
                    bool head = false;
                    UNROLL for (int j = 0; j < C; j++) {
                        const int i = l * C + j;
                        const int o = __float_as_int(ov[j]);
                        if (o == i) { run = 0; head = true; }
                        ex[j] = run;
                        if (i >= cx.w0 && i < cx.w1 && ((fvm >> j) & 1u)) run += freq_to_inc(f[j], sk);
                    }
                    // segmented inclusive scan of (run, head) over lanes
                    u64 sum = run;
                    int flag = head ? 1 : 0;
                    UNROLL for (int d = 1; d < 32; d <<= 1) {
                        const u64 ts = __shfl_up_sync(FULL, sum, d);
                        const int tf = __shfl_up_sync(FULL, flag, d);
                        if (l >= d) {
                            if (!flag) sum += ts;
                            flag |= tf;
                        }
                    }
                    u64 carry = __shfl_up_sync(FULL, sum, 1);
                    if (l == 0) carry = 0;
                    bool seen = false;
                    UNROLL for (int j = 0; j < C; j++) {
                        const int i = l * C + j;
                        const int o = __float_as_int(ov[j]);
                        if (o == i) seen = true;
                        u64 p = ex[j];
                        if (!seen) p += carry;
                        if (o < 0) p += acc0;
                        acc[j] = sin_turns(p + ph[j], fast);
                    }
                    const u64 tot = __shfl_sync(FULL, sum, 31);
                    const int tflag = __shfl_sync(FULL, flag, 31);
                    newacc = tflag ? tot : acc0 + tot;
                }
                __syncwarp();
                st_state64(M.state, in.a, newacc);
                cx.vm = vm;
                break;
            }
            case S_ALT_CC: {
                const float cp = M.cval[in.a], cn = M.cval[in.b];
                UNROLL for (int j = 0; j < C; j++) acc[j] = acc[j] >= 0.0f ? cp : cn;
                break;
            }
            case S_ALT_POS: {
                UNROLL for (int j = 0; j < C; j++) if (!((cx.vm >> j) & 1u)) acc[j] = 0.0f;
                slot_store(M.slots, in.a, acc);
                break;
            }
            case S_ALT_END: {
                float t[C], pv[C];
                slot_load(M.slots, in.a, t);
                const uint32_t tv = M.slot_vm[in.a * 32 + l];
                if (in.b >= 0) slot_load(M.slots, in.b, pv);
                else { const float c = M.cval[~in.b]; UNROLL for (int j = 0; j < C; j++) pv[j] = c; }
                if (in.c < 0) { const float c = M.cval[~in.c]; UNROLL for (int j = 0; j < C; j++) acc[j] = c; }
                else { UNROLL for (int j = 0; j < C; j++) if (!((cx.vm >> j) & 1u)) acc[j] = 0.0f; }
                UNROLL for (int j = 0; j < C; j++) acc[j] = t[j] >= 0.0f ? pv[j] : acc[j];
                cx.vm = tv;
                break;
            }
            case S_RESET_BEGIN: {
                PUSH(cx.seg_slot);
                PUSH(cx.o_last);
                M.slot_vm[in.b * 32 + l] = (uint8_t)cx.vm;  // trigger validity = output validity
                const int o_last = reset_origins(M, acc, cx.vm, cx.w0, n, in.a, in.b, cx.seg_slot);
                cx.seg_slot = in.b;
                cx.o_last = o_last;
                break;
            }
            case S_RESET_END: {
                UNROLL for (int j = 0; j < C; j++) if (!((cx.vm >> j) & 1u)) acc[j] = 0.0f;
                cx.vm = M.slot_vm[cx.seg_slot * 32 + l];
                cx.o_last = POP();
                cx.seg_slot = POP();
                break;
            }
            case S_APP_BEGIN: {  // Append under a Reset (lower.cpp emit_seg): local time of the live run at w0
                const u64 pos = ld_state64(M.state, P.goe[in.a].term_arg);
                PUSH((int)(pos < 0x3fffffffull ? pos : 0x3fffffffull));
                break;
            }
            case S_APP_MID: {
                // The first part is in acc with its validity.  Where does the second part start in each run?
                // Run origin o >= 0: at o + La; a run that began before this tile (o = -1) and stood at local
                // time P0 at w0: it is already running when P0 >= La (origin -1: its own carried state),
                // else it starts at w0 + La - P0.  Samples before that point get an origin beyond
                // themselves; the stateful nodes of the second part compute garbage there, masked in
                // S_APP_END.
                const int P0 = POP();
                const tb_goe g = P.goe[in.c];
                float value = 0.0f;
                for (uint32_t s = 0; s < g.n_steps; s++) {
                    const int sign = P.goe_steps[g.step_off + 2 * s];
                    const float c = M.cval[P.goe_steps[g.step_off + 2 * s + 1]];
                    value = sign > 0 ? __fadd_rn(value, c) : __fsub_rn(value, c);
                }
                const u64 target = f32_as_usize(ceilf(__fmul_rn(value, srf)));
                const int La = (int)(target < 0x3fffffffull ? target : 0x3fffffffull);
                slot_store(M.slots, in.a, acc);
                M.slot_vm[in.a * 32 + l] = (uint8_t)cx.vm;
                float ov[C];
                slot_load(M.slots, cx.seg_slot, ov);
                const int far = 0x3fffffff;
                int mine = far;
                const int last_pos = cx.w1 - 1;
                UNROLL for (int j = 0; j < C; j++) {
                    const int i = l * C + j;
                    const int o = __float_as_int(ov[j]);
                    int ob;
                    if (o >= 0) ob = La < far - o ? o + La : far;
                    else ob = P0 >= La ? -1 : (La - P0 < far - cx.w0 ? cx.w0 + (La - P0) : far);
                    ov[j] = __int_as_float(ob);
                    if (i == last_pos) mine = ob;
                }
                slot_store(M.slots, in.b, ov);
                int ob_last = __shfl_sync(FULL, mine, (last_pos >> 3) & 31);
                if (n <= 0) ob_last = -1;
                // The run live at the end of the window: second part running since before the tile (-1),
                // started inside it (its origin), or not started (w1: "zero samples in", state stays Initial).
                PUSH(cx.seg_slot);
                PUSH(cx.o_last);
                cx.seg_slot = in.b;
                cx.o_last = ob_last < 0 ? -1 : (ob_last > last_pos ? cx.w1 : ob_last);
                cx.vm = 0xffu;
                break;
            }
            case S_APP_END: {
                float ov[C], av[C];
                slot_load(M.slots, in.b, ov);
                slot_load(M.slots, in.a, av);
                const uint32_t va = M.slot_vm[in.a * 32 + l];
                uint32_t vm = 0;
                UNROLL for (int j = 0; j < C; j++) {
                    const int i = l * C + j;
                    const int ob = __float_as_int(ov[j]);
                    const bool in_a = (va >> j) & 1u;
                    const bool in_b = !in_a && (ob < 0 || i >= ob) && ((cx.vm >> j) & 1u);
                    acc[j] = in_a ? av[j] : (in_b ? acc[j] : 0.0f);
                    vm |= ((in_a || in_b) ? 1u : 0u) << j;
                }
                const bool idle = cx.o_last >= cx.w1;  // the second part has not started in the live run
                cx.o_last = POP();
                cx.seg_slot = POP();
                cx.vm = vm;
                if (idle) {  // ... so it is still in its Initial state, whatever the masked samples left behind
                    __syncwarp();
                    const int s0 = in.c & 0xffff, sn = (in.c >> 16) & 0x7fff;
                    for (int k = l; k < sn; k += 32) M.state[s0 + k] = 0u;
                    __syncwarp();
                }
                break;
            }
            case S_FIN: {  // Fin with an analytic length inside a Reset run
                const tb_goe g = P.goe[in.a];
                float value = 0.0f;
                for (uint32_t s = 0; s < g.n_steps; s++) {
                    const int sign = P.goe_steps[g.step_off + 2 * s];
                    const float c = M.cval[P.goe_steps[g.step_off + 2 * s + 1]];
                    value = sign > 0 ? __fadd_rn(value, c) : __fsub_rn(value, c);
                }
                if (g.term == GOE_TIME) {
                    const u64 pos = ld_state64(M.state, g.term_arg);
                    const u64 target = f32_as_usize(ceilf(__fmul_rn(value, srf)));
                    float ov[C];
                    slot_load(M.slots, cx.seg_slot, ov);
                    UNROLL for (int j = 0; j < C; j++) {
                        const int i = l * C + j;
                        const int o = __float_as_int(ov[j]);
                        const u64 nloc = o >= 0 ? (u64)(i - o) : pos + (u64)(i - cx.w0);
                        if (!(nloc < target)) cx.vm &= ~(1u << j);
                    }
                    __syncwarp();
                    st_state64(M.state, g.term_arg,
                               cx.o_last >= 0 ? (u64)(cx.w1 - cx.o_last) : pos + (u64)n);
                } else if (g.term == GOE_CONST) {
                    if (M.cval[g.term_arg] >= value) cx.vm = 0u;
                }
                break;
            }
            default: return;
        }
        // Post-ops: `npost` constant-operand point ops (generator.rs:541-548) fused behind the
        // instruction that produced the accumulator; each is one extra code word.
        for (uint32_t np = in.op >> 16; np > 0; np--) {
            const tb_insn po = code[pc++];
            const float c = M.cval[po.b];
            APPLY_OP((uint32_t)po.a, acc, acc[j], c)
        }
    }
#undef PUSH
#undef POP
}

#include "steady.cuh"

// ------------------------------------------------------------------------------------------
// Per-voice setup: constant table (is_const folding, generator.rs:574-612) and derived constants.
// ------------------------------------------------------------------------------------------
__device__ void setup_voice(const tb_launch& P, const WarpMem& M, const float* prow, const SineK& sk) {
    const int l = lane_id();
    if (l == 0) {
        for (uint32_t k = 0; k < P.n_cval; k++) {
            const tb_cexpr e = P.cexpr[k];
            float v;
            if (e.kind == CE_LIT) v = e.value;
            else if (e.kind == CE_PARAM) v = prow ? prow[e.a] : e.value;
            else if (e.kind == CE_NEG) v = -M.cval[e.a];
            else v = apply1(e.op, M.cval[e.a], M.cval[e.b]);
            M.cval[k] = v;
        }
    }
    __syncwarp();
    for (uint32_t t = 0; t < P.n_aux; t++) {  // warp-uniform walk; lanes share the work of one entry
        const tb_aux a = P.aux[t];
        if (a.kind == AUX_SINE_INC) {
            if (l == 0) M.aux[a.off] = turns_to_fx_slow((double)M.cval[a.a] / (TB_TAU * (double)P.sample_rate));
        } else if (a.kind == AUX_SINE_ROT) {  // steady stream: increment + the 16 rotations (cos, sin)(j * inc)
            if (P.n_samples < (u64)TB_TILE_S) continue;  // no steady tile in this launch
            const u64 inc = turns_to_fx_slow((double)M.cval[a.a] / (TB_TAU * (double)P.sample_rate));
            if (l == 0) M.aux[a.off] = inc;
            if (l < TB_CS) {
                const u64 ang = inc * (u64)l;
                double2 r;
                r.x = l == 0 ? 1.0 : sin_turns_d8(ang + 0x4000000000000000ull);
                r.y = l == 0 ? 0.0 : sin_turns_d8(ang);
                reinterpret_cast<double2*>(M.aux + a.off + 2)[l] = r;
            }
        } else if (a.kind == AUX_SINE_PHASE) {
            if (l == 0) M.aux[a.off] = turns_to_fx_slow((double)M.cval[a.a] / TB_TAU);
        } else if (a.kind == AUX_FILT_COEF) {  // coefficient values of a constant filter: b_0.., a_1..
            const tb_filter_tab* ft = &P.filt[a.b];
            float* cf = reinterpret_cast<float*>(M.aux + a.off);
            for (uint32_t e = l; e < ft->K + ft->J; e += 32) cf[e] = M.cval[~ft->coef[e]];
        } else if (P.exact_fb) {  // AUX_FILT_POW is for the scans only
            continue;
        } else if (l == 0) {  // AUX_FILT_POW: A^(8*2^k), k = 0..5, A the companion matrix of the feedback taps
            const tb_filter_tab* ft = &P.filt[a.b];
            const int J = ft->J, K = ft->K;
            double A[TB_MAX_J * TB_MAX_J], B[TB_MAX_J * TB_MAX_J];
            for (int r = 0; r < J; r++)
                for (int c = 0; c < J; c++)
                    A[r * J + c] = r == 0 ? -(double)M.cval[~ft->coef[K + c]] : (c == r - 1 ? 1.0 : 0.0);
            double* out = reinterpret_cast<double*>(M.aux + a.off);
            for (int step = 0; step < 3 + 6; step++) {
                if (step >= 3) {
                    for (int e = 0; e < J * J; e++) out[(step - 3) * J * J + e] = A[e];
                    if (step == 8) break;
                }
                for (int r = 0; r < J; r++)
                    for (int c = 0; c < J; c++) {
                        double s = 0.0;
                        for (int m = 0; m < J; m++) s = fma(A[r * J + m], A[m * J + c], s);
                        B[r * J + c] = s;
                    }
                for (int e = 0; e < J * J; e++) A[e] = B[e];
            }
        }
    }
    __syncwarp();
}

}  // namespace

// ------------------------------------------------------------------------------------------
// Kernel: grid = ceil(n_voices / warps), block = 32 * warps; warps = TB_WARPS_PER_CTA unless the
// program's per-warp shared-memory footprint forces fewer (abi.cpp).
// ------------------------------------------------------------------------------------------
#ifndef TB_MIN_BLOCKS
#define TB_MIN_BLOCKS 2
#endif
extern "C" __global__ void __launch_bounds__(32 * TB_WARPS_PER_CTA, TB_MIN_BLOCKS)
tb_render_kernel(const tb_launch P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, l = threadIdx.x & 31;
    // CTA-shared: the program.
    tb_insn* code = reinterpret_cast<tb_insn*>(smem_raw);
    for (uint32_t t = threadIdx.x; t < P.n_code; t += blockDim.x) code[t] = P.code[t];
    __syncthreads();
    const uint32_t voice = blockIdx.x * (blockDim.x >> 5) + warp;
    if (voice >= P.n_voices) return;
    if (P.done && P.done[voice]) {  // already returned short earlier in this (chunked) call
        if (l == 0 && P.out_len && !P.accumulate) P.out_len[voice] = 0ull;
        return;
    }

    size_t off = ((size_t)P.n_code * sizeof(tb_insn) + 15) & ~(size_t)15;
    // A root Fin with an analytic length over a steady tree (tb_launch::lane_fin_goe, lower.cpp) carries the
    // ST_* stream of its inner tree: whole tiles well inside the note run through the steady interpreter too.
    const bool fin = P.lane_fin_goe >= 0;
    const bool st_cap = P.steady_ok || (fin && !P.lane_clk);  // clocked words (ST_*_CLK) are for the lane kernels
    const size_t per_warp_slots = (size_t)P.n_slots * (st_cap ? TILE_S : TILE) * sizeof(float);
    const size_t aux_b = ((size_t)P.aux_words * 8 + 15) & ~(size_t)15;
    const size_t cval_b = ((size_t)P.n_cval * 4 + 15) & ~(size_t)15;
    const size_t state_b = ((size_t)P.state_words * 4 + 15) & ~(size_t)15;
    const size_t slen_b = ((size_t)P.n_slots * 4 + 15) & ~(size_t)15;
    const size_t svm_b = ((size_t)P.n_slots * 32 + 15) & ~(size_t)15;
    const size_t per_warp = per_warp_slots + aux_b + cval_b + state_b + slen_b + svm_b;
    unsigned char* base = smem_raw + off + per_warp * warp;
    WarpMem M;
    // time-axis split (program.h tb_launch::vsplit*): `voice` is a virtual voice — a segment of a real one
    const uint32_t vsplit = P.vsplit > 1u ? P.vsplit : 1u;
    const uint32_t rvoice = voice / vsplit;
    const uint32_t vseg_i = P.vseg_lo + (voice - rvoice * vsplit);  // segment of the voice
    const size_t vidx = P.vsplit_total > 1u ? (size_t)rvoice * P.vsplit_total + vseg_i : (size_t)voice;  // state block
    M.voice = rvoice;
    M.slots = reinterpret_cast<float*>(base); base += per_warp_slots;
    M.aux = reinterpret_cast<u64*>(base); base += aux_b;
    M.cval = reinterpret_cast<float*>(base); base += cval_b;
    M.state = reinterpret_cast<uint32_t*>(base); base += state_b;
    M.slot_len = reinterpret_cast<int*>(base); base += slen_b;
    M.slot_vm = reinterpret_cast<uint8_t*>(base);

    SineK sk;
    sk.kscale = 17592186044416.0 / (TB_TAU * (double)P.sample_rate);
    sk.pscale = 17592186044416.0 / TB_TAU;
    sk.inv_turn = 1.0 / (TB_TAU * (double)P.sample_rate);
    sk.flimit = (float)(100.0 * TB_TAU * (double)P.sample_rate);
    sk.plimit = 600.0f;

    uint32_t* gstate = P.state + vidx * P.state_words;
    for (uint32_t t = l; t < P.state_words; t += 32) M.state[t] = gstate[t];
    setup_voice(P, M, P.params ? P.params + (size_t)rvoice * P.n_params : nullptr, sk);

    int ctl[TB_CTL_DEPTH];
    float acc[C];
    Ctx cx;
    cx.seg_slot = -1;
    cx.o_last = -1;
    cx.vm = 0xffu;
    u64 total = 0;
    if (P.mode == 0) {
        float* row = P.out ? P.out + (size_t)rvoice * P.out_stride + (size_t)vseg_i * P.vseg : nullptr;
        const bool vec_ok = row && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
        const uint32_t code_s = (uint32_t)__cvta_generic_to_shared(code);
        // Whole tiles take the steady path once every filter holds its full history
        // (generator.rs:234-252): normally after the first general tile of a stream.
        auto filters_ready = [&]() {
            bool ready = true;
            for (uint32_t fi = 0; fi < P.n_filt; fi++) {
                const uint32_t* S = M.state + P.filt[fi].state_off;
                ready = ready && S[0] != 0u && S[1] == P.filt[fi].K - 1u;
            }
            return ready;
        };
        bool steady = st_cap && filters_ready();
        // Root Fin: where the note ends (greater_or_equals_at over the flattened chain, goe_eval) and the
        // position state of the Time node its length is measured on.
        u64 fin_target = ~0ull;
        int fin_time = -1;
        if (fin) {
            const tb_goe g = P.goe[P.lane_fin_goe];
            float value = 0.0f;
            for (uint32_t k = 0; k < g.n_steps; k++) {
                const int sign = P.goe_steps[g.step_off + 2 * k];
                const float c = M.cval[P.goe_steps[g.step_off + 2 * k + 1]];
                value = sign > 0 ? __fadd_rn(value, c) : __fsub_rn(value, c);
            }
            if (g.term == GOE_TIME) {
                fin_time = g.term_arg;
                fin_target = f32_as_usize(ceilf(__fmul_rn(value, (float)P.sample_rate)));
            } else if (M.cval[g.term_arg] >= value) {
                fin_target = 0ull;  // over before it starts: the general tile says so
            }
        }
        // The lane index held in a register for the steady loop (S2R would otherwise be re-issued,
        // with its latency, wherever the compiler rematerialises threadIdx).
        int lane_pinned;
        asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane_pinned));
        u64 tbase = 0;
        while (tbase < P.n_samples) {
            const u64 left = P.n_samples - tbase;
            bool inside = true;  // a root Fin: the whole tile lies inside the note, and it is not the tile that
                                 // applies the reference's per-call "already over?" test (generator.rs:808-809)
            if (fin) {
                const u64 pos = fin_time >= 0 ? ld_state64(M.state, fin_time) : 0ull;
                inside = (tbase > 0 || P.mid_call) && fin_target >= pos + (u64)TILE_S;
            }
            if (steady && inside && left >= (u64)TILE_S) {
                // Whole tile, every node infinite, histories complete: the steady-state interpreter.
                float sacc[CS];
                if (P.fast_mode == 2) run_steady<2>(P, code_s, M, sacc, sk, lane_pinned);
                else run_steady<1>(P, code_s, M, sacc, sk, lane_pinned);
                if (row) {
                    float* dst = row + tbase + l * CS;
                    if (vec_ok) {
                        UNROLL for (int q = 0; q < CS / 4; q++)
                            reinterpret_cast<float4*>(dst)[q] =
                                make_float4(sacc[4 * q], sacc[4 * q + 1], sacc[4 * q + 2], sacc[4 * q + 3]);
                    } else {
                        UNROLL for (int j = 0; j < CS; j++) dst[j] = sacc[j];
                    }
                }
                if (fin_time >= 0) {  // Fin advances both children to the end of the block (generator.rs:141-167)
                    __syncwarp();
                    const u64 pos = ld_state64(M.state, fin_time);
                    __syncwarp();
                    if (l == 0) st_state64(M.state, fin_time, pos + (u64)TILE_S);
                    __syncwarp();
                }
                total += (u64)TILE_S;
                tbase += (u64)TILE_S;
                continue;
            }
            cx.w0 = 0;
            cx.w1 = left < (u64)TILE ? (int)left : TILE;
            cx.L = 0;
            cx.first_tile = tbase == 0 && !P.mid_call;
            run_program(P, code, M, cx, acc, (int)P.pc_gen, sk, ctl);
            const int L = cx.L;
            if (row) {
                float* dst = row + tbase + l * C;
                if (vec_ok && l * C + C <= L) {
                    reinterpret_cast<float4*>(dst)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                    reinterpret_cast<float4*>(dst)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
                } else {
                    UNROLL for (int j = 0; j < C; j++) if (l * C + j < L) dst[j] = acc[j];
                }
            }
            total += (u64)L;
            if (L < cx.w1) break;
            tbase += (u64)TILE;
            if (st_cap && !steady) {
                __syncwarp();
                steady = filters_ready();
            }
        }
    } else {
        const u64 step = P.pure_len ? (u64)(1 << 30) : (u64)TILE;
        for (u64 tbase = 0; tbase < P.n_samples; tbase += step) {
            const u64 left = P.n_samples - tbase;
            cx.w0 = 0;
            cx.w1 = left < step ? (int)left : (int)step;
            cx.L = 0;
            cx.first_tile = tbase == 0 && !P.mid_call;
            run_program(P, code, M, cx, acc, (int)P.pc_len, sk, ctl);
            total += (u64)cx.L;
            if (cx.L < cx.w1) break;
        }
    }
    __syncwarp();
    for (uint32_t t = l; t < P.state_words; t += 32) gstate[t] = M.state[t];
    if (l == 0) {
        if (P.out_len) P.out_len[vidx] = (P.accumulate ? P.out_len[vidx] : 0ull) + total;
        if (P.done && total < P.n_samples) P.done[voice] = 1;
    }
}

extern "C" size_t tb_kernel_smem_bytes(uint32_t n_code, uint32_t n_slots, uint32_t aux_words,
                                       uint32_t n_cval, uint32_t state_words, uint32_t steady_ok,
                                       uint32_t warps) {
    size_t off = ((size_t)n_code * sizeof(tb_insn) + 15) & ~(size_t)15;
    const size_t per_warp = (size_t)n_slots * (steady_ok ? TILE_S : TILE) * sizeof(float) + (((size_t)aux_words * 8 + 15) & ~(size_t)15) +
                            (((size_t)n_cval * 4 + 15) & ~(size_t)15) + (((size_t)state_words * 4 + 15) & ~(size_t)15) +
                            (((size_t)n_slots * 4 + 15) & ~(size_t)15) + (((size_t)n_slots * 32 + 15) & ~(size_t)15);
    return off + per_warp * warps;
}

extern "C" cudaError_t tb_kernel_launch(const tb_launch* P, size_t smem, uint32_t warps, cudaStream_t stream) {
    static size_t configured_on[64] = {};  // per device; raised, never lowered
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    size_t& configured = configured_on[dev & 63];
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(tb_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const uint32_t grid = (P->n_voices + warps - 1) / warps;
    tb_render_kernel<<<grid, 32 * warps, smem, stream>>>(*P);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Mixdown: mix[i] (+)= sum over voices, in voice index order, of rows[v][i] for i < lens[v] —
// the tracker's serial `out[filled + j] += tmp[j]` (tracker.rs:617-619).  One thread per sample,
// rows read coalesced; `t0` is the time offset of this chunk inside the voices' streams.  lens == NULL:
// every row is full (the per-warp partial sums of the lane kernel's on-chip mixdown).
// ------------------------------------------------------------------------------------------
extern "C" __global__ void __launch_bounds__(256)
tb_mix_kernel(const float* __restrict__ rows, uint64_t stride, const unsigned long long* __restrict__ lens,
              uint32_t n_voices, uint64_t n_samples, uint64_t t0, float* __restrict__ mix, int accumulate) {
    // Four consecutive samples per thread (one 128-bit load per voice when the rows allow it), eight
    // voices in flight; the adds stay in voice index order, each rounded (tracker.rs:617-619).
    const uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= n_samples) return;
    const bool vec = ((stride & 3) == 0) && ((reinterpret_cast<uintptr_t>(rows) & 15) == 0) && i + 4 <= n_samples;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (accumulate) {
        UNROLL for (int k = 0; k < 4; k++) if (i + k < n_samples) s[k] = mix[i + k];
    }
    if (vec) {
        uint32_t v = 0;
        for (; v + 8 <= n_voices; v += 8) {
            float4 x[8];
            unsigned long long len[8];
            UNROLL for (int u = 0; u < 8; u++) {
                len[u] = lens ? lens[v + u] : ~0ull;
                x[u] = __ldcs(reinterpret_cast<const float4*>(rows + (size_t)(v + u) * stride + i));
            }
            UNROLL for (int u = 0; u < 8; u++) {
                if (t0 + i + 0 < len[u]) s[0] = __fadd_rn(s[0], x[u].x);
                if (t0 + i + 1 < len[u]) s[1] = __fadd_rn(s[1], x[u].y);
                if (t0 + i + 2 < len[u]) s[2] = __fadd_rn(s[2], x[u].z);
                if (t0 + i + 3 < len[u]) s[3] = __fadd_rn(s[3], x[u].w);
            }
        }
        for (; v < n_voices; v++) {
            const unsigned long long len = lens ? lens[v] : ~0ull;
            const float4 x = __ldcs(reinterpret_cast<const float4*>(rows + (size_t)v * stride + i));
            if (t0 + i + 0 < len) s[0] = __fadd_rn(s[0], x.x);
            if (t0 + i + 1 < len) s[1] = __fadd_rn(s[1], x.y);
            if (t0 + i + 2 < len) s[2] = __fadd_rn(s[2], x.z);
            if (t0 + i + 3 < len) s[3] = __fadd_rn(s[3], x.w);
        }
    } else {
        for (uint32_t v = 0; v < n_voices; v++) {
            const unsigned long long len = lens ? lens[v] : ~0ull;  // total generated so far, including this chunk
            UNROLL for (int k = 0; k < 4; k++)
                if (i + k < n_samples && t0 + i + k < len) s[k] = __fadd_rn(s[k], rows[(size_t)v * stride + i + k]);
        }
    }
    UNROLL for (int k = 0; k < 4; k++) if (i + k < n_samples) mix[i + k] = s[k];
}

extern "C" cudaError_t tb_mix_launch(const float* rows, uint64_t stride, const unsigned long long* lens,
                                     uint32_t n_voices, uint64_t n_samples, uint64_t t0, float* mix,
                                     int accumulate, cudaStream_t stream) {
    const uint32_t grid = (uint32_t)((n_samples + 1023) / 1024);
    tb_mix_kernel<<<grid, 256, 0, stream>>>(rows, stride, lens, n_voices, n_samples, t0, mix, accumulate);
    return cudaGetLastError();
}
