// steady.cuh — the steady-state interpreter of render.cu (included there, inside its anonymous
// namespace, after the helpers it shares with the general interpreter).
//
// Most of a render is spent far from any boundary: every node of the program is an infinite
// waveform, the window is a whole tile, all filter histories are complete and nothing finishes.
// For programs whose generate code contains only such nodes (lower.cpp: `steady_ok`) the kernel
// runs the SAME byte-code through this second interpreter, which
//   * walks tiles of 32 x CS = 512 samples (lane l owns [16 l, 16 l + 16)): half the dispatches
//     and half the warp scans per sample of the general path,
//   * carries no window / length / validity bookkeeping at all,
//   * evaluates constant-rate sines (generator.rs:206-219 with Const frequency and phase) by
//     angle addition in f64: one sin/cos pair per lane per tile, then
//     sin(a + j d) = sin a cos(j d) + cos a sin(j d) against a per-voice table of the 16 rotations,
//   * prefetches the next instruction word while the current one executes.
// State blocks, constants and the carried semantics are those of the general interpreter, so the
// two can alternate tile by tile (the first tile of a launch and the tail always run there).
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

// Steady slots: float4 #q (q = 0..3) of lane l at float4 index q*32 + l (conflict free).
__device__ __forceinline__ void sslot_store(float* slots, int s, const float (&v)[CS]) {
    float4* p = reinterpret_cast<float4*>(slots + (size_t)s * TILE_S) + lane_id();
    UNROLL for (int q = 0; q < CS / 4; q++) p[32 * q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
__device__ __forceinline__ void sslot_load(const float* slots, int s, float (&v)[CS]) {
    const float4* p = reinterpret_cast<const float4*>(slots + (size_t)s * TILE_S) + lane_id();
    UNROLL for (int q = 0; q < CS / 4; q++) {
        const float4 t = p[32 * q];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}

// sin(2 pi ph / 2^64) as a double, 8 coefficients (|err| < 5e-16; tools/fit_sine.py SIN_D8 with
// the 2^-62 scaling of the integer phase folded in).
__constant__ double c_sin_exact8[8] = {0x1.921fb54442d17p-62,  -0x1.4abbce625bd83p-187, 0x1.466bc677522bdp-314,
                                       -0x1.32d2cce1ea145p-442, 0x1.5078327046959p-571,  -0x1.e30631bdf732dp-701,
                                       0x1.e89f6fe44fe7bp-831,  -0x1.62903d02bb153p-961};
__device__ __forceinline__ double sin_turns_d8(u64 ph) {
    double x = (double)(i64)ph;
    const int hi = __double2hiint(x);
    const double half = __hiloint2double((hi & 0x80000000) | 0x43e00000, 0);
    const double folded = half - x;
    x = ((hi & 0x7fffffff) > 0x43d00000) ? folded : x;
    const double z = x * x;
    double p = c_sin_exact8[7];
    UNROLL for (int k = 6; k >= 0; k--) p = fma(p, z, c_sin_exact8[k]);
    return x * p;
}

// FAST class through the special-function unit: the top 32 phase bits as radians in [-pi, pi),
// sin.approx (range reduction multiply + MUFU.SIN), |err| <= 2^-21.4 (CUDA math API, __sinf on
// [-pi, pi]).  Selected by tb_launch::fast_mode == 2.
__device__ __forceinline__ float sin_turns_mufu(u64 ph) {
    return __sinf((float)(int)(ph >> 32) * 1.4629180792671596e-09f);  // 2 pi / 2^32
}

template <int MODE>  // 0 exact, 1 f32 polynomial, 2 MUFU
__device__ __forceinline__ float sin_turns_m(u64 ph) {
    return MODE == 0 ? sin_turns_exact(ph) : (MODE == 1 ? sin_turns_fast(ph) : sin_turns_mufu(ph));
}

// Constant frequency and phase: angle addition against the per-voice rotation table
// rot[j] = (cos, sin)(2 pi j inc / 2^64), j < CS  (setup_voice, AUX_SINE_ROT).
__device__ __forceinline__ void steady_sine_cc(float (&acc)[CS], u64 inc, u64 ph0, const double2* rot,
                                               uint32_t* state, int st) {
    const u64 acc0 = ld_state64(state, st);
    const u64 pb = acc0 + inc * (u64)(lane_id() * CS) + ph0;
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
                                                 u64 ph0, uint32_t* state, int st, const SineK& sk) {
    const u64 acc0 = ld_state64(state, st);
    float big = 0.0f, bigp = 0.0f;
    UNROLL for (int j = 0; j < CS; j++) {
        big = fmaxf(big, fabsf(f[j]));
        if (!UNIFORM_PH) bigp = fmaxf(bigp, fabsf(p[j]));
    }
    const bool slow = __any_sync(FULL, !(big < sk.flimit) || !(bigp < sk.plimit));
    u64 ph[CS];
    u64 run = 0;
    if (!slow) {
        UNROLL for (int j = 0; j < CS; j++) {
            const u64 inc = magic_to_fx(fma((double)f[j], sk.kscale, 6755399441055744.0));
            ph[j] = UNIFORM_PH ? run : run + magic_to_fx(fma((double)p[j], sk.pscale, 6755399441055744.0));
            run += inc;
        }
    } else {
        UNROLL for (int j = 0; j < CS; j++) {
            ph[j] = UNIFORM_PH ? run : run + phase_to_fx(p[j], sk);
            run += freq_to_inc(f[j], sk);
        }
    }
    const u64 incl = warp_incl_sum(run);
    const u64 base = acc0 + (incl - run) + (UNIFORM_PH ? ph0 : 0ull);
    const u64 total = __shfl_sync(FULL, incl, 31);
    UNROLL for (int j = 0; j < CS; j++) acc[j] = sin_turns_m<MODE>(ph[j] + base);
    __syncwarp();
    st_state64(state, st, acc0 + total);
}

// Constant frequency, phase offsets in p (phase modulation).
template <int MODE>
__device__ __forceinline__ void steady_sine_ca(float (&acc)[CS], u64 inc, const float (&p)[CS], uint32_t* state,
                                               int st, const SineK& sk) {
    const u64 acc0 = ld_state64(state, st);
    u64 b = acc0 + inc * (u64)(lane_id() * CS);
    float bigp = 0.0f;
    UNROLL for (int j = 0; j < CS; j++) bigp = fmaxf(bigp, fabsf(p[j]));
    if (__any_sync(FULL, !(bigp < sk.plimit))) {
        UNROLL for (int j = 0; j < CS; j++) {
            acc[j] = sin_turns_m<MODE>(b + phase_to_fx(p[j], sk));
            b += inc;
        }
    } else {
        UNROLL for (int j = 0; j < CS; j++) {
            acc[j] = sin_turns_m<MODE>(b + magic_to_fx(fma((double)p[j], sk.pscale, 6755399441055744.0)));
            b += inc;
        }
    }
    __syncwarp();
    st_state64(state, st, acc0 + inc * (u64)TILE_S);
}

// Constant-coefficient filter over a whole tile with complete history (generator.rs:382-515).
// Same arithmetic, operation order and carried deques as filter_full_tile of the general path.
template <int J>
__device__ __forceinline__ void steady_filter(const WarpMem& M, const tb_filter_tab* ft, float (&acc)[CS],
                                              uint32_t* S) {
    const int l = lane_id();
    const int K = ft->K;
    float* hx = reinterpret_cast<float*>(S + 2);
    float* hy = hx + (K - 1);
    // pe[m] = x[-1 - m]: the sample m + 1 places before this lane's chunk.
    // (All TB_MAX_K - 1 are fetched unconditionally: eight shuffles per 16 samples cost less than
    // keeping conditionally defined registers alive across the tap loop.)
    float pe[TB_MAX_K - 1];
    UNROLL for (int m = 0; m < TB_MAX_K - 1; m++) {
        const float t = __shfl_up_sync(FULL, acc[CS - 1 - m], 1);
        pe[m] = (l == 0) ? ((m < K - 1) ? hx[K - 2 - m] : 0.0f) : t;
    }
    float u[CS];
    {
        const float b0 = M.cval[~ft->coef[0]];
        UNROLL for (int j = 0; j < CS; j++) u[j] = __fmul_rn(acc[j], b0);
    }
    UNROLL for (int k = 1; k < TB_MAX_K; k++) {
        if (k < K) {
            const float bk = M.cval[~ft->coef[k]];
            UNROLL for (int j = 0; j < CS; j++) u[j] = __fadd_rn(u[j], __fmul_rn(bk, (j >= k) ? acc[j - k] : pe[k - j - 1]));
        }
    }
    __syncwarp();
    if (l == 31) {  // the input deque keeps the last K-1 inputs of the tile (oldest first)
        UNROLL for (int m = 0; m < TB_MAX_K - 1; m++)
            if (m < K - 1) hx[K - 2 - m] = acc[CS - 1 - m];
    }
    if (J > 0) {
        float a[J > 0 ? J : 1];
        UNROLL for (int jj = 0; jj < J; jj++) a[jj] = M.cval[~ft->coef[K + jj]];
        const double* mpow = reinterpret_cast<const double*>(M.aux + ft->pow_aux) + J * J;  // A^(16*2^k)
        float s[J > 0 ? J : 1];
        UNROLL for (int jj = 0; jj < J; jj++) s[jj] = (l == 0) ? hy[J - 1 - jj] : 0.0f;
        UNROLL for (int j = 0; j < CS; j++) {  // pass 1: chunk response from a zero state
            float y = u[j];
            UNROLL for (int jj = 0; jj < J; jj++) y = fmaf(-a[jj], s[jj], y);
            UNROLL for (int jj = J - 1; jj > 0; jj--) s[jj] = s[jj - 1];
            s[0] = y;
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
    } else {
        UNROLL for (int j = 0; j < CS; j++) acc[j] = u[j];
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

// One steady tile.  `code_s` is the shared-memory address of the program, `pc` its entry point.
template <int FASTMODE>
__device__ __forceinline__ void run_steady(const tb_launch& P, uint32_t code_s, const WarpMem& M, float (&acc)[CS],
                                           int pc, const SineK& sk) {
    const int l = lane_id();
    const float srf = (float)P.sample_rate;
    uint32_t ip = code_s + (uint32_t)pc * (uint32_t)sizeof(tb_insn);
    tb_insn nxt = lds_insn(ip);
    for (;;) {
        const tb_insn in = nxt;
        ip += sizeof(tb_insn);
        nxt = lds_insn(ip);
        const uint32_t op = in.op & 0xffu;
        const bool fast = ((in.op >> 8) & 0xffu) == TB_SINE_FAST;
        switch (op) {
            case OP_END: return;
            case G_CONST: {
                const float c = M.cval[in.a];
                UNROLL for (int j = 0; j < CS; j++) acc[j] = c;
                break;
            }
            case G_TIME: {  // generator.rs:101-111
                const u64 pos = ld_state64(M.state, in.a) + (u64)(l * CS);
                UNROLL for (int j = 0; j < CS; j++) acc[j] = __fdiv_rn(__ull2float_rn(pos + (u64)j), srf);
                __syncwarp();
                st_state64(M.state, in.a, pos - (u64)(l * CS) + (u64)TILE_S);
                break;
            }
            case G_BINC: {  // generator.rs:538-549; every operand is infinite here, so Merge is Add
                const float c = M.cval[in.b];
                APPLY_OP_S((uint32_t)in.a, acc, acc[j], c)
                break;
            }
            case G_BIN_BEGIN:
            case G_SINE_BEGIN:
            case G_ALT_BEGIN:
            case G_ALT_POS: sslot_store(M.slots, in.a, acc); break;
            case G_BIN_END: {
                float av[CS];
                sslot_load(M.slots, in.a, av);
                APPLY_OP_S((uint32_t)in.b, acc, av[j], acc[j])
                break;
            }
            case G_SINE_CC:
                steady_sine_cc(acc, M.aux[in.b], M.aux[in.c], reinterpret_cast<const double2*>(M.aux + in.b + 2),
                               M.state, in.a);
                break;
            case G_SINE_AC: {
                float f[CS];
                UNROLL for (int j = 0; j < CS; j++) f[j] = acc[j];
                if (!fast) steady_sine_scan<true, 0>(acc, f, f, M.aux[in.c], M.state, in.a, sk);
                else steady_sine_scan<true, FASTMODE>(acc, f, f, M.aux[in.c], M.state, in.a, sk);
                break;
            }
            case G_SINE_CA: {
                float p[CS];
                UNROLL for (int j = 0; j < CS; j++) p[j] = acc[j];
                if (!fast) steady_sine_ca<0>(acc, M.aux[in.b], p, M.state, in.a, sk);
                else steady_sine_ca<FASTMODE>(acc, M.aux[in.b], p, M.state, in.a, sk);
                break;
            }
            case G_SINE_END: {
                float f[CS], p[CS];
                sslot_load(M.slots, in.b, f);
                UNROLL for (int j = 0; j < CS; j++) p[j] = acc[j];
                if (!fast) steady_sine_scan<false, 0>(acc, f, p, 0ull, M.state, in.a, sk);
                else steady_sine_scan<false, FASTMODE>(acc, f, p, 0ull, M.state, in.a, sk);
                break;
            }
            case G_ALT_CC: {  // generator.rs:335-341
                const float cp = M.cval[in.a], cn = M.cval[in.b];
                UNROLL for (int j = 0; j < CS; j++) acc[j] = acc[j] >= 0.0f ? cp : cn;
                break;
            }
            case G_ALT_END: {
                float t[CS];
                sslot_load(M.slots, in.a, t);
                if (in.c < 0) { const float c = M.cval[~in.c]; UNROLL for (int j = 0; j < CS; j++) acc[j] = c; }
                if (in.b >= 0) {
                    float pv[CS];
                    sslot_load(M.slots, in.b, pv);
                    UNROLL for (int j = 0; j < CS; j++) acc[j] = t[j] >= 0.0f ? pv[j] : acc[j];
                } else {
                    const float c = M.cval[~in.b];
                    UNROLL for (int j = 0; j < CS; j++) acc[j] = t[j] >= 0.0f ? c : acc[j];
                }
                break;
            }
            case G_FILT_PRE:  // history is complete: skip the pre-read block
                ip = code_s + (uint32_t)in.c * (uint32_t)sizeof(tb_insn);
                nxt = lds_insn(ip);
                break;
            case G_FILT_RUN: {
                const tb_filter_tab* ft = &P.filt[in.b];
                uint32_t* S = M.state + in.a;
                switch (ft->J) {
                    case 0: steady_filter<0>(M, ft, acc, S); break;
                    case 1: steady_filter<1>(M, ft, acc, S); break;
                    case 2: steady_filter<2>(M, ft, acc, S); break;
                    case 3: steady_filter<3>(M, ft, acc, S); break;
                    default: steady_filter<4>(M, ft, acc, S); break;
                }
                break;
            }
            default: return;  // unreachable: lower.cpp admits only the ops above
        }
        for (uint32_t np = in.op >> 16; np > 0; np--) {  // fused constant post-ops
            const tb_insn po = nxt;
            ip += sizeof(tb_insn);
            nxt = lds_insn(ip);
            const float c = M.cval[po.b];
            APPLY_OP_S((uint32_t)po.a, acc, acc[j], c)
        }
        __syncwarp();
    }
}
