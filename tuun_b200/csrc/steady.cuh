// steady.cuh — the steady-state interpreter of render.cu (included there, inside its anonymous
// namespace, after the helpers it shares with the general interpreter).
//
// Most of a render is spent far from any boundary: every node of the program is an infinite
// waveform, the window is a whole tile, all filter histories are complete and nothing finishes.
// For trees made only of such nodes lower.cpp emits, next to the general byte-code, a second
// straight-line stream of ST_* words (program.h) that this interpreter runs:
//   * tiles of 32 x CS = 512 samples (lane l owns [16 l, 16 l + 16)): half the dispatches and
//     half the warp scans per sample of the general path,
//   * no window / length / validity bookkeeping, no jumps, every operand pre-resolved
//     (filter coefficients as values, K and J in the instruction word),
//   * constant-rate sines (generator.rs:206-219 with Const frequency and phase) by angle
//     addition in f64: one sin/cos pair per lane per tile, then
//     sin(a + j d) = sin a cos(j d) + cos a sin(j d) against a per-voice table of 16 rotations,
//   * constant point operators as branch-free `acc * m + a` post-op words (two roundings),
//   * FAST-class sines from the top 32 phase bits only (the 2^-32-turn carry is dropped),
//   * the next instruction word prefetched while the current one executes.
// State blocks, constants and carried semantics are those of the general interpreter, so the two
// alternate tile by tile (the first tile of a stream and the tail of a launch always run there).
#pragma once

constexpr int CS = TB_CS;
constexpr int TILE_S = TB_TILE_S;
static_assert(CS == 16, "steady slot layout and history code assume 16 samples per lane");

__device__ __forceinline__ tb_insn lds_insn(uint32_t saddr) {
    tb_insn r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.op), "=r"(r.a), "=r"(r.b), "=r"(r.c)
                 : "r"(saddr)
                 : "memory");
    return r;
}

__device__ __forceinline__ u64 warp_incl_sum_l(u64 x, int l) {
    UNROLL for (int d = 1; d < 32; d <<= 1) {
        const u64 t = __shfl_up_sync(FULL, x, d);
        if (l >= d) x += t;
    }
    return x;
}

// Steady slots: float4 #q (q = 0..3) of lane l at float4 index q*32 + l (conflict free).
__device__ __forceinline__ void sslot_store(float* slots, int s, const float (&v)[CS], int l) {
    float4* p = reinterpret_cast<float4*>(slots + (size_t)s * TILE_S) + l;
    UNROLL for (int q = 0; q < CS / 4; q++) p[32 * q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
__device__ __forceinline__ void sslot_load(const float* slots, int s, float (&v)[CS], int l) {
    const float4* p = reinterpret_cast<const float4*>(slots + (size_t)s * TILE_S) + l;
    UNROLL for (int q = 0; q < CS / 4; q++) {
        const float4 t = p[32 * q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}

// Constant frequency and phase: angle addition against the per-voice rotation table
// rot[j] = (cos, sin)(2 pi j inc / 2^64), j < CS  (setup_voice, AUX_SINE_INC).
__device__ __forceinline__ void steady_sine_cc(float (&acc)[CS], u64 inc, u64 ph0, const double2* rot,
                                               uint32_t* state, int st, int l) {
    const u64 acc0 = ld_state64(state, st);
    const u64 pb = acc0 + inc * (u64)(l * CS) + ph0;
    const double S = sin_turns_d8(pb);
    const double Cq = sin_turns_d8(pb + 0x4000000000000000ull);  // a quarter turn ahead: the cosine
    UNROLL for (int j = 0; j < CS; j++) {
        const double2 r = rot[j];
        acc[j] = (float)fma(S, r.x, Cq * r.y);
    }
    __syncwarp();
    st_state64(state, st, acc0 + inc * (u64)TILE_S);
}

// Variable frequency (and optionally variable phase): 64-bit prefix sum over the tile.
//   f    frequencies (rad/s)
//   p    phase offsets (rad), unused when UNIFORM_PH
// One warp vote up front decides whether any input is outside the exact range of the
// magic-number conversion (|f| >= 100 tau sr, |p| >= 600 rad: never for audio).
template <bool UNIFORM_PH, int MODE>
__device__ __forceinline__ void steady_sine_scan(float (&acc)[CS], const float (&f)[CS], const float (&p)[CS],
                                                 u64 ph0, uint32_t* state, int st, const SineK& sk, int l) {
    const u64 acc0 = ld_state64(state, st);
    float big = 0.0f, bigp = 0.0f;
    UNROLL for (int j = 0; j < CS; j++) {
        big = fmaxf(big, fabsf(f[j]));
        if (!UNIFORM_PH) bigp = fmaxf(bigp, fabsf(p[j]));
    }
    const bool slow = __any_sync(FULL, !(big < sk.flimit) || !(bigp < sk.plimit));
    if (MODE != 0 && !slow) {
        // FAST: only the top 32 bits of each sample's phase are formed; the running sum of the
        // lane stays 64-bit (it feeds the carried accumulator, which must stay exact).
        int h[CS];
        u64 raw = 0;
        UNROLL for (int j = 0; j < CS; j++) {
            h[j] = raw_hi(raw);
            if (!UNIFORM_PH) h[j] += raw_hi(magic_raw(p[j], sk.pscale));
            raw += magic_raw(f[j], sk.kscale);
        }
        const u64 run = raw << 20;
        const u64 incl = warp_incl_sum_l(run, l);
        const u64 base = acc0 + (incl - run) + (UNIFORM_PH ? ph0 : 0ull);
        const u64 total = __shfl_sync(FULL, incl, 31);
        const int bh = (int)(base >> 32);
        UNROLL for (int j = 0; j < CS; j++) acc[j] = sin_hi<MODE>(h[j] + bh);
        __syncwarp();
        st_state64(state, st, acc0 + total);
        return;
    }
    u64 ph[CS];
    u64 run = 0;
    if (!slow) {
        UNROLL for (int j = 0; j < CS; j++) {
            ph[j] = UNIFORM_PH ? run : run + (magic_raw(p[j], sk.pscale) << 20);
            run += magic_raw(f[j], sk.kscale) << 20;
        }
    } else {
        UNROLL for (int j = 0; j < CS; j++) {
            ph[j] = UNIFORM_PH ? run : run + phase_to_fx(p[j], sk);
            run += freq_to_inc(f[j], sk);
        }
    }
    const u64 incl = warp_incl_sum_l(run, l);
    const u64 base = acc0 + (incl - run) + (UNIFORM_PH ? ph0 : 0ull);
    const u64 total = __shfl_sync(FULL, incl, 31);
    UNROLL for (int j = 0; j < CS; j++) {
        const u64 a = ph[j] + base;
        acc[j] = MODE == 0 ? sin_turns_exact(a) : sin_hi<MODE>((int)(a >> 32));
    }
    __syncwarp();
    st_state64(state, st, acc0 + total);
}

// Constant frequency, phase offsets in p (phase modulation).
template <int MODE>
__device__ __forceinline__ void steady_sine_ca(float (&acc)[CS], u64 inc, const float (&p)[CS], uint32_t* state,
                                               int st, const SineK& sk, int l) {
    const u64 acc0 = ld_state64(state, st);
    u64 b = acc0 + inc * (u64)(l * CS);
    float bigp = 0.0f;
    UNROLL for (int j = 0; j < CS; j++) bigp = fmaxf(bigp, fabsf(p[j]));
    if (__any_sync(FULL, !(bigp < sk.plimit))) {
        UNROLL for (int j = 0; j < CS; j++) {
            const u64 a = b + phase_to_fx(p[j], sk);
            acc[j] = MODE == 0 ? sin_turns_exact(a) : sin_hi<MODE>((int)(a >> 32));
            b += inc;
        }
    } else if (MODE == 0) {
        UNROLL for (int j = 0; j < CS; j++) {
            acc[j] = sin_turns_exact(b + (magic_raw(p[j], sk.pscale) << 20));
            b += inc;
        }
    } else {
        UNROLL for (int j = 0; j < CS; j++) {
            acc[j] = sin_hi<MODE>((int)(b >> 32) + raw_hi(magic_raw(p[j], sk.pscale)));
            b += inc;
        }
    }
    __syncwarp();
    st_state64(state, st, acc0 + inc * (u64)TILE_S);
}

// ---- constant-coefficient filter over a whole tile with complete history (generator.rs:382-515) ----
// Same arithmetic, operation order and carried deques as filter_full_tile of the general path.
// Feed-forward part: u[j] = x[j] b0 + b1 x[j-1] + ... (each product and sum rounded, :496-499).
// KT > 0: K known at compile time (1, 2, 3 cover every filter of lib/v0/std.tuun); KT == 0: any K.
template <int KT>
__device__ __forceinline__ void steady_fir(float (&u)[CS], const float (&x)[CS], const float* b, int K, float* hx,
                                           int l) {
    constexpr int NP = KT > 0 ? (KT > 1 ? KT - 1 : 1) : TB_MAX_K - 1;
    // pe[m] = x[-1 - m]: the sample m + 1 places before this lane's chunk.
    float pe[NP];
    UNROLL for (int m = 0; m < NP; m++) {
        pe[m] = 0.0f;
        if (KT == 0 || m < KT - 1) {
            const float t = __shfl_up_sync(FULL, x[CS - 1 - m], 1);
            pe[m] = (l == 0) ? ((m < K - 1) ? hx[K - 2 - m] : 0.0f) : t;
        }
    }
    {
        const float b0 = b[0];
        UNROLL for (int j = 0; j < CS; j++) u[j] = __fmul_rn(x[j], b0);
    }
    UNROLL for (int k = 1; k < (KT > 0 ? KT : TB_MAX_K); k++) {
        if (KT > 0 || k < K) {
            const float bk = b[k];
            UNROLL for (int j = 0; j < CS; j++) u[j] = __fadd_rn(u[j], __fmul_rn(bk, (j >= k) ? x[j - k] : pe[k - j - 1]));
        }
    }
    __syncwarp();
    if (l == 31) {  // the input deque keeps the last K-1 inputs of the tile (oldest first)
        UNROLL for (int m = 0; m < NP; m++)
            if (KT > 0 ? (m < KT - 1) : (m < K - 1)) hx[K - 2 - m] = x[CS - 1 - m];
    }
}

// Feedback part: scan of affine state transitions (see iir_scan_const in render.cu); mpow holds
// A^(16 * 2^k), k = 0..4.
template <int J>
__device__ __forceinline__ void steady_iir(float (&acc)[CS], const float (&u)[CS], const float* af,
                                           const double* mpow, float* hy, int l) {
    float a[J];
    UNROLL for (int jj = 0; jj < J; jj++) a[jj] = af[jj];
    float s[J];
    UNROLL for (int jj = 0; jj < J; jj++) s[jj] = (l == 0) ? hy[J - 1 - jj] : 0.0f;
    UNROLL for (int j = 0; j < CS; j++) {  // pass 1: chunk response from a zero state
        float y = u[j];
        UNROLL for (int jj = 0; jj < J; jj++) y = fmaf(-a[jj], s[jj], y);
        UNROLL for (int jj = J - 1; jj > 0; jj--) s[jj] = s[jj - 1];
        s[0] = y;
    }
    double v[J];
    UNROLL for (int jj = 0; jj < J; jj++) v[jj] = (double)s[jj];
    UNROLL for (int k = 0; k < 5; k++) {
        const int d = 1 << k;
        double t[J];
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
    UNROLL for (int j = 0; j < CS; j++) {  // pass 2: the reference's f32 operation order (generator.rs:500-502)
        float y = u[j];
        UNROLL for (int jj = 0; jj < J; jj++) y = __fsub_rn(y, __fmul_rn(a[jj], s[jj]));
        UNROLL for (int jj = J - 1; jj > 0; jj--) s[jj] = s[jj - 1];
        s[0] = y;
        acc[j] = y;
    }
    __syncwarp();
    if (l == 31) {
        UNROLL for (int jj = 0; jj < J; jj++) hy[J - 1 - jj] = acc[CS - 1 - jj];
    }
}

__device__ __forceinline__ void steady_filter(float (&acc)[CS], uint32_t* S, int K, int J, const float* coef,
                                              const double* pow8, int l) {
    float* hx = reinterpret_cast<float*>(S + 2);
    float* hy = hx + (K - 1);
    float u[CS];
    const double* mpow = pow8 + J * J;  // skip A^8: the steady chunk is 16 samples
    if (K == 3 && J == 2) {
        // The biquad (every filter of lib/v0/std.tuun): feed-forward and feedback in one basic
        // block, so the independent FIR arithmetic fills the latency of the serial recurrences.
        steady_fir<3>(u, acc, coef, 3, hx, l);
        steady_iir<2>(acc, u, coef + 3, mpow, hy, l);
        return;
    }
    switch (K) {
        case 1: steady_fir<1>(u, acc, coef, K, hx, l); break;
        case 2: steady_fir<2>(u, acc, coef, K, hx, l); break;
        case 3: steady_fir<3>(u, acc, coef, K, hx, l); break;
        default: steady_fir<0>(u, acc, coef, K, hx, l); break;
    }
    switch (J) {
        case 0: { UNROLL for (int j = 0; j < CS; j++) acc[j] = u[j]; break; }
        case 1: steady_iir<1>(acc, u, coef + K, mpow, hy, l); break;
        case 2: steady_iir<2>(acc, u, coef + K, mpow, hy, l); break;
        case 3: steady_iir<3>(acc, u, coef + K, mpow, hy, l); break;
        default: steady_iir<4>(acc, u, coef + K, mpow, hy, l); break;
    }
}

#define APPLY_OP_S(OPV, DST, A, B)                                                    \
    switch (OPV) {                                                                    \
        case TB_ADD:                                                                  \
        case TB_MERGE: UNROLL for (int j = 0; j < CS; j++) DST[j] = __fadd_rn(A, B); break; \
        case TB_SUBTRACT: UNROLL for (int j = 0; j < CS; j++) DST[j] = __fsub_rn(A, B); break; \
        case TB_MULTIPLY: UNROLL for (int j = 0; j < CS; j++) DST[j] = __fmul_rn(A, B); break; \
        case TB_DIVIDE:                                                               \
            UNROLL for (int j = 0; j < CS; j++) {                                     \
                float bb_ = (B);                                                      \
                DST[j] = bb_ == 0.0f ? 0.0f : __fdiv_rn(A, bb_);                      \
            }                                                                         \
            break;                                                                    \
        default: UNROLL for (int j = 0; j < CS; j++) DST[j] = powf(A, B); break;      \
    }

// One steady tile.  `code_s` is the shared-memory address of the program.
template <int FASTMODE>
__device__ __forceinline__ void run_steady(const tb_launch& P, uint32_t code_s, const WarpMem& M, float (&acc)[CS],
                                           const SineK& sk, const int l) {
    const float srf = (float)P.sample_rate;
    uint32_t ip = code_s + P.pc_steady * (uint32_t)sizeof(tb_insn);
    tb_insn nxt = lds_insn(ip);
    for (;;) {
        const tb_insn in = nxt;
        ip += sizeof(tb_insn);
        nxt = lds_insn(ip);
        const uint32_t op = in.op & 0xffu;
        const bool fast = ((in.op >> 8) & 0xffu) == TB_SINE_FAST;
        switch (op) {
            case ST_END: return;
            case ST_CONST: {
                const float c = M.cval[in.a];
                UNROLL for (int j = 0; j < CS; j++) acc[j] = c;
                break;
            }
            case ST_TIME: {  // generator.rs:101-111
                const u64 pos0 = ld_state64(M.state, in.a);
                const u64 pos = pos0 + (u64)(l * CS);
                UNROLL for (int j = 0; j < CS; j++) acc[j] = __fdiv_rn(__ull2float_rn(pos + (u64)j), srf);
                __syncwarp();
                st_state64(M.state, in.a, pos0 + (u64)TILE_S);
                break;
            }
            case ST_NOISE: {  // generator.rs:113-118
                const u64 pos = ld_state64(M.state, in.a);
                const u64 stream = noise_stream(P, M.voice, in.b);
                UNROLL for (int j = 0; j < CS; j++) acc[j] = noise_at(stream, pos + (u64)(l * CS + j));
                __syncwarp();
                st_state64(M.state, in.a, pos + (u64)TILE_S);
                break;
            }
            case ST_SAVE: sslot_store(M.slots, in.a, acc, l); break;
            case ST_BIN: {  // generator.rs:555-567 with both sides infinite
                float av[CS];
                sslot_load(M.slots, in.a, av, l);
                APPLY_OP_S((uint32_t)in.b, acc, av[j], acc[j])
                break;
            }
            case ST_SINE_CC:
                steady_sine_cc(acc, M.aux[in.b], M.aux[in.c], reinterpret_cast<const double2*>(M.aux + in.b + 2),
                               M.state, in.a, l);
                break;
            case ST_SINE_AC: {
                float f[CS];
                UNROLL for (int j = 0; j < CS; j++) f[j] = acc[j];
                if (!fast) steady_sine_scan<true, 0>(acc, f, f, M.aux[in.c], M.state, in.a, sk, l);
                else steady_sine_scan<true, FASTMODE>(acc, f, f, M.aux[in.c], M.state, in.a, sk, l);
                break;
            }
            case ST_SINE_CA: {
                float p[CS];
                UNROLL for (int j = 0; j < CS; j++) p[j] = acc[j];
                if (!fast) steady_sine_ca<0>(acc, M.aux[in.b], p, M.state, in.a, sk, l);
                else steady_sine_ca<FASTMODE>(acc, M.aux[in.b], p, M.state, in.a, sk, l);
                break;
            }
            case ST_SINE_AA: {
                float f[CS], p[CS];
                sslot_load(M.slots, in.b, f, l);
                UNROLL for (int j = 0; j < CS; j++) p[j] = acc[j];
                if (!fast) steady_sine_scan<false, 0>(acc, f, p, 0ull, M.state, in.a, sk, l);
                else steady_sine_scan<false, FASTMODE>(acc, f, p, 0ull, M.state, in.a, sk, l);
                break;
            }
            case ST_ALT_CC: {  // generator.rs:335-341
                const float cp = M.cval[in.a], cn = M.cval[in.b];
                UNROLL for (int j = 0; j < CS; j++) acc[j] = acc[j] >= 0.0f ? cp : cn;
                break;
            }
            case ST_ALT: {
                float t[CS];
                sslot_load(M.slots, in.a, t, l);
                if (in.c < 0) { const float c = M.cval[~in.c]; UNROLL for (int j = 0; j < CS; j++) acc[j] = c; }
                if (in.b >= 0) {
                    float pv[CS];
                    sslot_load(M.slots, in.b, pv, l);
                    UNROLL for (int j = 0; j < CS; j++) acc[j] = t[j] >= 0.0f ? pv[j] : acc[j];
                } else {
                    const float c = M.cval[~in.b];
                    UNROLL for (int j = 0; j < CS; j++) acc[j] = t[j] >= 0.0f ? c : acc[j];
                }
                break;
            }
            case ST_FILT:
                steady_filter(acc, M.state + in.a, (int)((in.op >> 8) & 0xfu), (int)((in.op >> 12) & 0x7u),
                              reinterpret_cast<const float*>(M.aux + in.b),
                              reinterpret_cast<const double*>(M.aux + in.c), l);
                break;
            default: return;  // unreachable: lower.cpp emits only the words above
        }
        for (uint32_t np = in.op >> 16; np > 0; np--) {  // constant point operators (generator.rs:541-548)
            const tb_insn po = nxt;
            ip += sizeof(tb_insn);
            nxt = lds_insn(ip);
            if ((po.op & 0xffu) == ST_AFFINE) {
                const float m = M.cval[po.b], a = M.cval[po.c];
                UNROLL for (int j = 0; j < CS; j++) acc[j] = __fadd_rn(__fmul_rn(acc[j], m), a);
            } else {
                const float c = M.cval[po.b];
                APPLY_OP_S((uint32_t)po.a, acc, acc[j], c)
            }
        }
        __syncwarp();
    }
}
