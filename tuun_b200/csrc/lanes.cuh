// lanes.cuh — the lane-per-voice kernels (lanes.cu, lanes_queue.cu, lanes_fm.cu; lanes_split.cu and lanes_fm_split.cu for
// the segments of a time-axis split): large batches of steady-state voices.
//
// render.cu gives every voice a warp and turns the reference's sequential state into warp scans.
// With tens of thousands of voices in one call the batch itself is the parallel axis: here ONE
// THREAD owns one voice and walks its time axis in tiles of TB_LS = 16 samples, so
//   * the phase accumulator of a Sine (generator.rs:206-219) is a running 64-bit sum in registers,
//   * the feedback of a Filter (generator.rs:500-507) is the reference's own recurrence, in the
//     reference's operation order, from the carried history — bit-identical to it for identical
//     inputs, with no scan, no shuffle and no second pass,
//   * constants, state and derived constants of the voice sit in the thread's own column of shared
//     memory (program.h: W words and Q units, conflict free), and so does the running result
//     between two instructions of the program,
//   * every second tile the 32 finished samples of the 32 voices of a warp leave through a transposed
//     read of that buffer, so that each output row receives a whole 128-byte piece (eight lanes cover
//     one row with one 128-bit store each),
//   * independent f32 operations of neighbouring samples are issued in pairs (FMUL2 / FADD2 / FFMA2).
// The program is the ST_* stream of the steady-state interpreter (steady.cuh) with operands
// rewritten and common chains fused by lower.cpp (build_lane_plan, fuse_lane_fm); state blocks are
// those of the other kernels, so launches of either kind continue one stream.  The host (abi.cpp
// launch_generate_seq) renders the first general tile of a stream and the < 16 samples of a call that
// do not fill a tile with tb_render_kernel — except for a program that is one fused FM voice, whose
// loop (run_fm_voice) starts the stream and takes those samples itself.  A thread may also own ONE
// SEGMENT of a voice (TB_LANES_VSPLIT: abi.cpp split_pass): parameters, noise streams and the row
// belong to the voice, the state block to the segment.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tuun_b200.h"
#include "program.h"

#ifndef TB_LANE_MIN_BLOCKS
#define TB_LANE_MIN_BLOCKS 7
#endif
// Time-axis split (program.h tb_launch::vsplit*): the kernels that take launches of virtual voices are compiled
// apart (lanes_split.cu, lanes_fm_split.cu) so that the plain ones keep their code, registers and schedule.
#ifndef TB_LANES_VSPLIT
#define TB_LANES_VSPLIT 0
#endif

namespace {

#include "common.cuh"


constexpr int LS = TB_LS;
constexpr int LT = TB_LANE_THREADS;
static_assert(LS == 16, "the symmetric rotation table and the row transpose assume 16 samples per tile");

// ---- the thread's column of shared memory ----------------------------------------------------------
constexpr int AS = LT + 1;  // chunk stride of the accumulator tiles, in 16-byte units (see lacc_store)
struct LaneMem {
    uint32_t* W;   // word w of this thread at W[w * LT]
    float4* Q;     // unit q of this thread at Q[q * LT]; slots follow the derived constants
    float4* A;     // the accumulator tile of the current step: chunk c (samples 4c .. 4c+3) of this thread at
                   // A[c * AS].  Two tiles alternate (chunks 0-3 and 4-7 of an 8-chunk buffer).
    uint32_t slot0;
};
__device__ __forceinline__ uint32_t ldw(const LaneMem& M, int w) { return M.W[w * LT]; }
__device__ __forceinline__ float ldf(const LaneMem& M, int w) { return __uint_as_float(M.W[w * LT]); }
__device__ __forceinline__ void stw(const LaneMem& M, int w, uint32_t v) { M.W[w * LT] = v; }
__device__ __forceinline__ void stf(const LaneMem& M, int w, float v) { M.W[w * LT] = __float_as_uint(v); }
__device__ __forceinline__ u64 ld64(const LaneMem& M, int w) {
    return (u64)M.W[w * LT] | ((u64)M.W[(w + 1) * LT] << 32);
}
__device__ __forceinline__ void st64(const LaneMem& M, int w, u64 v) {
    M.W[w * LT] = (uint32_t)v;
    M.W[(w + 1) * LT] = (uint32_t)(v >> 32);
}
__device__ __forceinline__ double ldd(const LaneMem& M, int w) { return __longlong_as_double((i64)ld64(M, w)); }
__device__ __forceinline__ void std_(const LaneMem& M, int w, double v) { st64(M, w, (u64)__double_as_longlong(v)); }
__device__ __forceinline__ void lslot_store(const LaneMem& M, int s, const float (&v)[LS]) {
    float4* p = M.Q + (size_t)(M.slot0 + 4 * s) * LT;
    UNROLL for (int q = 0; q < 4; q++) p[q * LT] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
__device__ __forceinline__ void lslot_load(const LaneMem& M, int s, float (&v)[LS]) {
    const float4* p = M.Q + (size_t)(M.slot0 + 4 * s) * LT;
    UNROLL for (int q = 0; q < 4; q++) {
        const float4 t = p[q * LT];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}
// The running result of the program lives in shared memory between instructions (registers carry
// it only inside one instruction: a 16-register value live across the dispatch costs a register
// move per value and dispatch).  The buffer doubles as the staging of the row transpose: two
// consecutive tiles sit side by side as chunks 0-3 and 4-7, and every second step a row leaves as
// one whole 128-byte piece (eight lanes x 16 bytes).  With a chunk stride of LT + 1 units the
// owner's stores (8 consecutive threads, one chunk) and the transposed loads (one row x 8 chunks)
// of a quarter warp both touch 8 distinct bank groups.
__device__ __forceinline__ void lacc_store(const LaneMem& M, const float (&v)[LS]) {
    UNROLL for (int q = 0; q < 4; q++) M.A[q * AS] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}
__device__ __forceinline__ void lacc_load(const LaneMem& M, float (&v)[LS]) {
    UNROLL for (int q = 0; q < 4; q++) {
        const float4 t = M.A[q * AS];
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
    }
}

// ---- packed f32x2 arithmetic (FMUL2 / FADD2 / FFMA2: one issue slot, two IEEE operations) ----------
// Every operation rounds to nearest like its scalar form.  ptxas contracts a mul.rn.f32x2 whose only
// use is an add.rn.f32x2 into one FFMA2 (one rounding instead of the reference's two); giving the
// multiply .ftz makes the pair unfusable.  .ftz only changes products of or into subnormals
// (|x| < 1.2e-38), far below anything audible or tested.
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

// ---- FAST-class sine on the special-function unit from the top 23 phase bits -----------------------
// p = phase in 2^-32 turns.  1.m with m = p >> 9 is the float 1 + frac (frac truncated to 2^-23), and
//   sin(2 pi (frac + 2^-24)) = sin(pi - 2 pi (frac + 2^-24)) = sin(3 pi - 2 pi 2^-24 - 2 pi (1 + frac)):
// one FFMA into (-pi, pi], then sin.approx (range multiply + MUFU.SIN).  The 2^-24 centres the
// truncation.  No integer-to-float conversion: that unit is shared with MUFU and with the
// f64 <-> f32 conversions of the exact sines.  Two samples share the FFMA (FFMA2).
#define TB_SIN23_A (-6.2831853071795865f)
#define TB_SIN23_B 9.4247776f  // 3 pi - 2 pi 2^-24
__device__ __forceinline__ uint32_t mant23(uint32_t p) { return __umulhi(p, 0x00800000u) + 0x3f800000u; }
__device__ __forceinline__ float sin_p32(uint32_t p) {
    return __sinf(fmaf(__uint_as_float(mant23(p)), TB_SIN23_A, TB_SIN23_B));
}
__device__ __forceinline__ void sin_p32x2(uint32_t p0, uint32_t p1, float& s0, float& s1) {
    float x0, x1;
    unpk2(fma2(pk2(__uint_as_float(mant23(p0)), __uint_as_float(mant23(p1))), pk2(TB_SIN23_A, TB_SIN23_A),
               pk2(TB_SIN23_B, TB_SIN23_B)),
          x0, x1);
    s0 = __sinf(x0);
    s1 = __sinf(x1);
}
// ---- the running phase of a frequency-modulated FAST sine ----------------------------------------
// Pd = 1.5 * 2^52 + P with P the phase in units of 2^-44 turns: the low 44 mantissa bits are the
// fraction of a turn, whole turns collect in the 7 bits above (and fall away when pd_make rebuilds
// the double, once per tile).  One DFMA per sample adds f * kscale exactly and rounds the SUM to the
// grid, so an increment is off by at most 2^-45 turns — for a constant rate a systematic 2^-45
// turns per sample, 7.9e-8 rad after 441,000 samples and 4.7e-7 rad after a minute (the 2^-32-turn
// grid this loop used before drifted 3.2e-4 rad in 10 s on constant-rate carriers).  The warp
// kernel's scan rounds increments to the same 2^-44 turns.  The reference's own f64 accumulator
// (generator.rs:212-218) is some 1e-11 rad off the exact phase after 10 s.
// Range: a tile may add up to +-2^51 units = +-128 turns; callers bound |f| to TB_FM_TURNS turns a
// sample (16 samples x 4 turns = 64) and convert the slow way beyond.
#define TB_FM_TURNS 4.0
#define TB_P44_MASK 0x00000fffffffffffull
__device__ __forceinline__ double pd_make(u64 p44) {
    return __longlong_as_double((i64)(0x4338000000000000ull | (p44 & TB_P44_MASK)));
}
__device__ __forceinline__ u64 pd_bits(double Pd) { return (u64)__double_as_longlong(Pd); }
// The float 1.m whose mantissa is the top 23 phase bits (what mant23 makes of a 32-bit phase).
__device__ __forceinline__ uint32_t pd_m23(double Pd) {
    return (__funnelshift_l((uint32_t)__double2loint(Pd), (uint32_t)__double2hiint(Pd), 11) & 0x007fffffu) | 0x3f800000u;
}
// The same in two instructions instead of three: `(x & 0x007fffff) | 0x3f800000` with both masks as literals is two
// LOP3s (the instruction holds one immediate); with 1.0f's bits in a register (SineK::one23, made from a launch
// parameter so that neither compiler folds it back) it is one.  16 instructions of 352 a tile in the FM voice loop.
__device__ __forceinline__ uint32_t pd_m23(double Pd, uint32_t one) {
    const uint32_t x = __funnelshift_l((uint32_t)__double2loint(Pd), (uint32_t)__double2hiint(Pd), 11);
    uint32_t r;
    asm("lop3.b32 %0, %1, 0x007fffff, %2, 0xEA;" : "=r"(r) : "r"(x), "r"(one));
    return r;
}
__device__ __forceinline__ uint32_t p44_m23(u64 p44) { return ((uint32_t)(p44 >> 21) & 0x007fffffu) | 0x3f800000u; }
__device__ __forceinline__ float sin_m23(uint32_t m) { return __sinf(fmaf(__uint_as_float(m), TB_SIN23_A, TB_SIN23_B)); }
__device__ __forceinline__ void sin_m23x2(uint32_t m0, uint32_t m1, float& s0, float& s1) {
    float x0, x1;
    unpk2(fma2(pk2(__uint_as_float(m0), __uint_as_float(m1)), pk2(TB_SIN23_A, TB_SIN23_A), pk2(TB_SIN23_B, TB_SIN23_B)),
          x0, x1);
    s0 = __sinf(x0);
    s1 = __sinf(x1);
}
// f as a double, by integer instructions: sign | (exponent + 896) << 20 | mantissa >> 3, mantissa << 29.
// An alternative to F2F.F64.F32, which runs on the 16-lane conversion unit that also serves MUFU.SIN
// and the f64 -> f32 conversions of the exact sines (tools/ubench/xu.cu: 14-16 results/clk/SM each).
// Measured on the FM voice loop: 5 ALU instructions cost more issue slots than the conversion costs
// unit time (32.3 ms against 31.1 ms per launch), so TB_F2F_ALU stays 0; kept for programs whose
// conversion unit is the busier side.  Zero and subnormal inputs come out as ~2^-127 (their
// increments still round to zero); infinities and NaNs are excluded by the range test that guards
// the magic-number conversion.
__device__ __forceinline__ double f32_to_f64_alu(float f) {
    // |f| x 2^29 as a 64-bit product is (|f| >> 3, |f| << 29) in one IMAD.WIDE, whose addend rebiases the exponent:
    // three instructions with the two sign operations.
    const uint32_t b = __float_as_uint(f);
    unsigned long long w;  // (as PTX: the compiler takes a multiplication by 2^29 apart into shifts)
    asm("mad.wide.u32 %0, %1, 0x20000000, %2;" : "=l"(w) : "r"(b & 0x7fffffffu), "l"(0x3800000000000000ull));
    const uint32_t hi = (uint32_t)(w >> 32) | (b & 0x80000000u);
    return __hiloint2double((int)hi, (int)(uint32_t)w);
}
// The same for f >= 0 (sign bit clear): the multiply-add alone.
__device__ __forceinline__ double f32_to_f64_pos(float f) {
    unsigned long long w;
    asm("mad.wide.u32 %0, %1, 0x20000000, %2;" : "=l"(w) : "r"(__float_as_uint(f)), "l"(0x3800000000000000ull));
    return __longlong_as_double((long long)w);
}
// What the phase warp of lanes_fm_ws.cu hands over for two samples (fm_carrier_tile<.., RAW>): the floats 1.m, or —
// TB_WS_ARG — the sines' arguments made from them (the FFMA2 of sin_m23x2 moved to the warp with time to spare).
#ifndef TB_WS_ARG
#define TB_WS_ARG 0
#endif
// TB_WS_SINES: the first TB_WS_SINES chunks of four samples leave as finished sines.
#ifndef TB_WS_SINES
#define TB_WS_SINES 0
#endif
__device__ __forceinline__ void raw_pair(uint32_t m0, uint32_t m1, float& r0, float& r1) {
#if TB_WS_ARG
    unpk2(fma2(pk2(__uint_as_float(m0), __uint_as_float(m1)), pk2(TB_SIN23_A, TB_SIN23_A), pk2(TB_SIN23_B, TB_SIN23_B)), r0, r1);
#else
    r0 = __uint_as_float(m0);
    r1 = __uint_as_float(m1);
#endif
}
// The low word of (f * scale + 1.5 * 2^52): rint(f * scale) mod 2^32.
#ifndef TB_F2F_ALU
#define TB_F2F_ALU 0
#endif

// The carrier frequency as a double in the running-phase DFMA of fm_carrier_tile: the conversion unit, or (TB_WIDEN_ALU) integer
// instructions.
#ifndef TB_WIDEN_ALU
#define TB_WIDEN_ALU 0
#endif
#if TB_WIDEN_ALU
#define TB_WIDEN(x) f32_to_f64_alu(x)
#else
#define TB_WIDEN(x) ((double)(x))
#endif

__device__ __forceinline__ uint32_t magic_lo(float f, double scale) {
#if TB_F2F_ALU
    return (uint32_t)__double2loint(fma(f32_to_f64_alu(f), scale, 6755399441055744.0));
#else
    return (uint32_t)__double2loint(fma((double)f, scale, 6755399441055744.0));
#endif
}

// ---- Sine, constant frequency and phase (generator.rs:206-219) -------------------------------------
// The thread carries (sin, cos) of the phase at the centre of the coming tile as doubles (set from
// the exact 64-bit accumulator when the kernel starts, setup_lane).  A tile is angle addition to
// both sides of the centre against the voice's rotations (cos, sin)(k d), k = 1..8 — the products
// are shared by samples 8 + k and 8 - k — and one more rotation by 16 d moves the pair to the next
// centre.  Rounding errors of the carried pair grow by ~1e-16 per tile (a random walk; 1e-13 after
// a minute of audio), against the 6e-8 of the f32 result.  The accumulator of the state block is
// advanced once, when the kernel ends.
__device__ __forceinline__ void lane_sine_cc(const LaneMem& M, float (&acc)[LS], int w_sc, int q) {
    const double S = ldd(M, w_sc), Cq = ldd(M, w_sc + 2);
    const double2* rot = reinterpret_cast<const double2*>(M.Q + (size_t)q * LT);
    acc[LS / 2] = (float)S;
    UNROLL for (int k = 1; k < LS / 2; k++) {
        const double2 r = rot[(size_t)(k - 1) * LT];
        const double b = Cq * r.y;
        acc[LS / 2 + k] = (float)fma(S, r.x, b);
        acc[LS / 2 - k] = (float)fma(S, r.x, -b);
    }
    {
        const double2 r = rot[(size_t)(LS / 2 - 1) * LT];
        acc[0] = (float)fma(S, r.x, -(Cq * r.y));
    }
    const double2 r16 = rot[(size_t)(LS / 2) * LT];
    std_(M, w_sc, fma(S, r16.x, Cq * r16.y));
    std_(M, w_sc + 2, fma(Cq, r16.x, -(S * r16.y)));
}

// ---- Sine with a frequency waveform (and optionally a phase waveform) ------------------------------
// acc holds the phase offsets on entry when !UNIFORM_PH; f the frequencies.  The sample is taken
// before the increment (generator.rs:212-218).
//   MODE 2 (FAST class on the special-function unit): the tile's phases are running sums on the 2^-44-turn
//   grid of the other kernels, and the 64-bit accumulator of the state block advances by exactly their sum.
//   CHECK = false: the caller has bounded |f| below the range limit of the magic-number conversion.
template <bool UNIFORM_PH, int MODE, bool CHECK = true>
__device__ __forceinline__ void lane_sine_var(const LaneMem& M, float (&acc)[LS], const float (&f)[LS], u64 ph0, int st,
                                              const SineK& sk) {
    const u64 a0 = ld64(M, st);
    float big = 0.0f, bigp = 0.0f;
    if (CHECK) {
        UNROLL for (int j = 0; j < LS; j++) {
            big = fmaxf(big, fabsf(f[j]));
            if (!UNIFORM_PH) bigp = fmaxf(bigp, fabsf(acc[j]));
        }
    }
    const bool slow = CHECK && (!(big < sk.flimit) || !(bigp < sk.plimit));
    if (MODE == 2 && !slow) {
        // Running phase on the 2^-44-turn grid in the mantissa of a double (pd_make): one DFMA per sample
        // adds the increment, a second one the phase offset of a phase waveform (not accumulated).  Whole
        // turns are dropped every 4 samples: 4 x TB_FM_TURNS + plimit (96 turns) stays inside the +-128 turns
        // the field holds.
        const u64 p0 = (a0 + (UNIFORM_PH ? ph0 : 0ull)) >> 20;
        u64 p = p0;
        UNROLL for (int j = 0; j < LS; j += 4) {
            double Pd = pd_make(p);
            UNROLL for (int i = j; i < j + 4; i += 2) {
                const double T0 = UNIFORM_PH ? Pd : fma((double)acc[i], sk.pscale, Pd);
                Pd = fma((double)f[i], sk.kscale, Pd);
                const double T1 = UNIFORM_PH ? Pd : fma((double)acc[i + 1], sk.pscale, Pd);
                Pd = fma((double)f[i + 1], sk.kscale, Pd);
                sin_m23x2(pd_m23(T0, sk.one23), pd_m23(T1, sk.one23), acc[i], acc[i + 1]);
            }
            p = pd_bits(Pd);
        }
        st64(M, st, a0 + ((p - p0) << 20));
        return;
    }
    u64 ph = a0;
    UNROLL for (int j = 0; j < LS; j++) {
        const u64 a = ph + (UNIFORM_PH ? ph0 : (slow ? phase_to_fx(acc[j], sk) : magic_raw(acc[j], sk.pscale) << 20));
        ph += slow ? freq_to_inc(f[j], sk) : magic_raw(f[j], sk.kscale) << 20;
        acc[j] = MODE == 0 ? sin_turns_exact(a) : (MODE == 1 ? sin_hi<1>((int)(a >> 32)) : sin_p32((uint32_t)(a >> 32)));
    }
    st64(M, st, ph);
}

// Constant frequency, phase offsets in acc (phase modulation).
template <int MODE>
__device__ __forceinline__ void lane_sine_ca(const LaneMem& M, float (&acc)[LS], u64 inc, int st, const SineK& sk) {
    const u64 a0 = ld64(M, st);
    float bigp = 0.0f;
    UNROLL for (int j = 0; j < LS; j++) bigp = fmaxf(bigp, fabsf(acc[j]));
    const bool slow = !(bigp < sk.plimit);
    u64 b = a0;
    if (MODE == 2 && !slow) {
        const double ps = sk.pscale * (1.0 / 4096.0);
        UNROLL for (int j = 0; j < LS; j += 2) {
            const uint32_t t0 = (uint32_t)(b >> 32) + magic_lo(acc[j], ps);
            b += inc;
            const uint32_t t1 = (uint32_t)(b >> 32) + magic_lo(acc[j + 1], ps);
            b += inc;
            sin_p32x2(t0, t1, acc[j], acc[j + 1]);
        }
    } else {
        UNROLL for (int j = 0; j < LS; j++) {
            const u64 a = b + (slow ? phase_to_fx(acc[j], sk) : magic_raw(acc[j], sk.pscale) << 20);
            b += inc;
            acc[j] = MODE == 0 ? sin_turns_exact(a) : (MODE == 1 ? sin_hi<1>((int)(a >> 32)) : sin_p32((uint32_t)(a >> 32)));
        }
    }
    st64(M, st, a0 + inc * (u64)LS);
}

// ---- Filter with constant coefficients and complete history (generator.rs:382-515) -----------------
// The reference's per-sample evaluation, literally: y = x b0; y += b_i x[n-i] ...; y -= a_j y[n-j] ...,
// every product and sum rounded (no FMA).  History layout of the state block is that of the other
// kernels: S[2 ..] the last K-1 inputs oldest first, then the last J outputs oldest first.
// Feed-forward products of two neighbouring samples share an FMUL2, and so do the sums whose
// operands pair up (even tap distances); the feedback recurrence is scalar.
template <int KT, int JT>
__device__ __forceinline__ void lane_filter(const LaneMem& M, float (&acc)[LS], int st, int w_coef, int K, int J) {
    constexpr int NK = KT >= 0 ? KT : TB_MAX_K;
    constexpr int NJ = JT >= 0 ? JT : TB_MAX_J;
    if (KT >= 0) K = KT;
    if (JT >= 0) J = JT;
    float b[NK > 0 ? NK : 1], a[NJ > 0 ? NJ : 1];
    float hx[NK > 1 ? NK - 1 : 1], hy[NJ > 0 ? NJ : 1];  // hx[m] = x[-1-m], hy[m] = y[-1-m]
    UNROLL for (int k = 0; k < NK; k++) b[k] = k < K ? ldf(M, w_coef + k) : 0.0f;
    UNROLL for (int j = 0; j < NJ; j++) a[j] = j < J ? ldf(M, w_coef + K + j) : 0.0f;
    const int wx = st + 2, wy = st + 2 + (K - 1);
    UNROLL for (int m = 0; m < NK - 1; m++) hx[m] = m < K - 1 ? ldf(M, wx + (K - 2 - m)) : 0.0f;
    UNROLL for (int m = 0; m < NJ; m++) hy[m] = m < J ? ldf(M, wy + (J - 1 - m)) : 0.0f;
    float u[LS];
    if (KT >= 1 && KT <= 4) {
        // pr[k][i] = b_k x[i] for i in [-(K-1), LS): aligned sample pairs by FMUL2, history by FMUL.
        float pr[NK][LS + NK];
        UNROLL for (int k = 0; k < NK; k++) {
            const u64 bb = pk2(b[k], b[k]);
            UNROLL for (int i = 0; i < LS; i += 2) {
                if (i + k < LS) unpk2(mul2(pk2(acc[i], acc[i + 1]), bb), pr[k][NK + i], pr[k][NK + i + 1]);
            }
            UNROLL for (int m = 1; m <= k; m++) pr[k][NK - m] = __fmul_rn(b[k], hx[m - 1]);
        }
        UNROLL for (int i = 0; i < LS; i += 2) {
            u64 s = pk2(pr[0][NK + i], pr[0][NK + i + 1]);
            UNROLL for (int k = 1; k < NK; k++) {
                if (k % 2 == 0) {
                    s = add2(s, pk2(pr[k][NK + i - k], pr[k][NK + i + 1 - k]));
                } else {
                    float s0, s1;
                    unpk2(s, s0, s1);
                    s0 = __fadd_rn(s0, pr[k][NK + i - k]);
                    s1 = __fadd_rn(s1, pr[k][NK + i + 1 - k]);
                    s = pk2(s0, s1);
                }
            }
            unpk2(s, u[i], u[i + 1]);
        }
        UNROLL for (int m = 0; m < NK - 1; m++) hx[m] = acc[LS - 1 - m];
    } else {
        UNROLL for (int j = 0; j < LS; j++) {
            const float x = acc[j];
            float y = __fmul_rn(x, b[0]);
            UNROLL for (int k = 1; k < NK; k++)
                if (KT >= 0 || k < K) y = __fadd_rn(y, __fmul_rn(b[k], hx[k - 1]));
            UNROLL for (int m = NK - 2; m > 0; m--) hx[m] = hx[m - 1];
            if (NK > 1) hx[0] = x;
            u[j] = y;
        }
    }
    if (JT == 2) {
        // y[n] = (u[n] - a1 y[n-1]) - a2 y[n-2]: both products of a new output, a1 y and a2 y, leave in
        // one FMUL2; the second waits one sample for its turn.
        const u64 aa = pk2(a[0], a[1]);
        float p1, p2, q2;  // a1 y[n-1], a2 y[n-2]; a2 y[n-1] for the next sample
        p1 = __fmul_rn(a[0], hy[0]);
        q2 = __fmul_rn(a[1], hy[0]);
        p2 = __fmul_rn(a[1], hy[1]);
        UNROLL for (int j = 0; j < LS; j++) {
            const float y = __fsub_rn(__fsub_rn(u[j], p1), p2);
            p2 = q2;
            unpk2(mul2(aa, pk2(y, y)), p1, q2);
            acc[j] = y;
        }
        hy[0] = acc[LS - 1];
        hy[1] = acc[LS - 2];
    } else {
        UNROLL for (int j = 0; j < LS; j++) {
            float y = u[j];
            UNROLL for (int m = 0; m < NJ; m++)
                if (JT >= 0 || m < J) y = __fsub_rn(y, __fmul_rn(a[m], hy[m]));
            UNROLL for (int m = NJ - 1; m > 0; m--) hy[m] = hy[m - 1];
            if (NJ > 0) hy[0] = y;
            acc[j] = y;
        }
    }
    UNROLL for (int m = 0; m < NK - 1; m++)
        if (m < K - 1) stf(M, wx + (K - 2 - m), hx[m]);
    UNROLL for (int m = 0; m < NJ; m++)
        if (m < J) stf(M, wy + (J - 1 - m), hy[m]);
}

// (n as f32) / (sample_rate as f32), correctly rounded (generator.rs:105-110).  For n < 2^24 — every clock
// within 6 minutes of its origin at 44.1 kHz — the quotient comes from the correctly rounded reciprocal and
// one exact residual (Markstein's sequence: q0 = RN(n r), e = n - q0 b exactly by FMA, q = RN(q0 + e r);
// correctly rounded for every b whose mantissa is not all ones; checked exhaustively over n < 2^24 for the
// common sample rates): 4 instructions instead of the ~12 of an IEEE division.
struct TimeDiv {
    float b, r;
    bool ok;
};
__device__ __forceinline__ TimeDiv time_div_setup(uint32_t sample_rate) {
    TimeDiv t;
    t.b = (float)sample_rate;
    t.r = __frcp_rn(t.b);
    t.ok = (__float_as_uint(t.b) & 0x007fffffu) != 0x007fffffu;
    return t;
}
__device__ __forceinline__ float time_div(u64 n, const TimeDiv& t) {
    if (t.ok && n < 16777216ull) {
        const float a = (float)(uint32_t)n;
        const float q0 = __fmul_rn(a, t.r);
        const float e = fmaf(-q0, t.b, a);
        return fmaf(e, t.r, q0);
    }
    return __fdiv_rn(__ull2float_rn(n), t.b);
}

#define APPLY_OP_L(OPV, DST, A, B)                                                    \
    switch (OPV) {                                                                    \
        case TB_ADD:                                                                  \
        case TB_MERGE: UNROLL for (int j = 0; j < LS; j++) DST[j] = __fadd_rn(A, B); break; \
        case TB_SUBTRACT: UNROLL for (int j = 0; j < LS; j++) DST[j] = __fsub_rn(A, B); break; \
        case TB_MULTIPLY: UNROLL for (int j = 0; j < LS; j++) DST[j] = __fmul_rn(A, B); break; \
        case TB_DIVIDE:                                                               \
            UNROLL for (int j = 0; j < LS; j++) {                                     \
                float bb_ = (B);                                                      \
                DST[j] = bb_ == 0.0f ? 0.0f : __fdiv_rn(A, bb_);                      \
            }                                                                         \
            break;                                                                    \
        default: UNROLL for (int j = 0; j < LS; j++) DST[j] = powf(A, B); break;      \
    }

__device__ __forceinline__ tb_insn lds_insn(uint32_t saddr) {
    tb_insn r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.op), "=r"(r.a), "=r"(r.b), "=r"(r.c)
                 : "r"(saddr)
                 : "memory");
    return r;
}

// One tile of one voice; leaves the result in the accumulator tile M.A.  Control flow is uniform
// over the CTA (every voice runs the same program).
template <int FASTMODE>
__device__ __forceinline__ void run_lane_tile(const tb_launch& P, uint32_t code_s, const LaneMem& M, uint32_t voice,
                                              const SineK& sk) {
    uint32_t ip = code_s;
    tb_insn nxt = lds_insn(ip);
    for (;;) {
        const tb_insn in = nxt;
        ip += sizeof(tb_insn);
        nxt = lds_insn(ip);
        const uint32_t op = in.op & 0xffu;
        const bool fast = ((in.op >> 8) & 0xffu) == TB_SINE_FAST;
        float acc[LS];
        switch (op) {
            case ST_END: return;
            case ST_CONST: {
                const float c = ldf(M, in.a);
                UNROLL for (int j = 0; j < LS; j++) acc[j] = c;
                break;
            }
            case ST_TIME: {  // generator.rs:101-111
                const TimeDiv td = time_div_setup(P.sample_rate);
                const u64 pos = ld64(M, in.a);
                UNROLL for (int j = 0; j < LS; j++) acc[j] = time_div(pos + (u64)j, td);
                st64(M, in.a, pos + (u64)LS);
                break;
            }
            case ST_NOISE: {  // generator.rs:113-118
                const u64 pos = ld64(M, in.a);
                const u64 stream = noise_stream(P, voice, in.b);
                UNROLL for (int j = 0; j < LS; j++) acc[j] = noise_at(stream, pos + (u64)j);
                st64(M, in.a, pos + (u64)LS);
                break;
            }
            case ST_SAVE:
                lacc_load(M, acc);
                lslot_store(M, in.a, acc);
                continue;  // the accumulator tile is unchanged; ST_SAVE never carries post-ops
            case ST_RESET_CLK: {
                // Reset (generator.rs:273-318): the running result is the trigger.  A restart happens at the
                // first non-negative sample after a negative one (the state word remembers the class of the
                // last sample; -0.0 counts as non-negative but re-arms).  What the inner tree needs is each
                // sample's local time: j - o behind a restart at o, else "j samples into a run that began
                // before this tile", encoded -1 - j, which every clocked node adds to its own carried position.
                // Nested in another Reset (in.c = its clock slot): where that one restarts, this one is Initial again —
                // "previously negative" (:276-279) — and so is everything under it (set_state, :311-313).
                lacc_load(M, acc);
                bool neg = ldw(M, in.a) == 0u;
                // second state word: the class of the first sample seen since it was last cleared (split.cu SP_CLK)
                if (ldw(M, in.a + 1) == 0u) stw(M, in.a + 1, acc[0] >= 0.0f ? 2u : 1u);
                int o = -1;
                float clk[LS];
                float oclk[LS];
                const bool nested = in.c >= 0;
                if (nested) lslot_load(M, in.c, oclk);
                // (selects, not branches: the state machine runs once a sample; a NaN changes nothing)
                if (nested) {
                    UNROLL for (int j = 0; j < LS; j++) {
                        const float x = acc[j];
                        const bool r = __float_as_int(oclk[j]) == 0;
                        neg = neg || r;
                        o = r ? j : o;
                        const bool fire = neg && x >= 0.0f;
                        o = fire ? j : o;
                        neg = fire ? (__float_as_int(x) < 0) : (neg || x < 0.0f);
                        clk[j] = __int_as_float(o >= 0 ? j - o : -1 - j);
                    }
                } else {
                    UNROLL for (int j = 0; j < LS; j++) {
                        const float x = acc[j];
                        const bool fire = neg && x >= 0.0f;
                        o = fire ? j : o;
                        neg = fire ? (__float_as_int(x) < 0) : (neg || x < 0.0f);
                        clk[j] = __int_as_float(o >= 0 ? j - o : -1 - j);
                    }
                }
                lslot_store(M, in.b, clk);
                stw(M, in.a, neg ? 0u : 1u);
                continue;
            }
            case ST_SEG_CLK: {  // a piece of a timeline: samples since it began (program.h)
                const tb_insn ex = nxt;  // second word: a = where the piece ends, b = words to jump (lower.cpp)
                ip += sizeof(tb_insn);
                const u64 p0 = ld64(M, in.a);
                // A tile that lies wholly in front of the piece or behind it: nothing of the piece is kept (the running
                // result stays what the pieces before it made it, or a later piece takes over), so its words are not
                // run — nor its ST_SEG_SEL, unless it is the last one.  (Positions are the same for every voice of a
                // launch: no divergence.)
                if (p0 + (u64)LS <= (u64)(uint32_t)in.c || p0 >= (u64)(uint32_t)ex.a) {
                    ip += (uint32_t)ex.b * (uint32_t)sizeof(tb_insn);
                    nxt = lds_insn(ip);
                    continue;
                }
                nxt = lds_insn(ip);
                float clk[LS];
                if (p0 < 0x7fff0000ull) {  // (pieces begin below 2^31: lower.cpp literal_target) 32-bit arithmetic
                    const int d0 = (int)(uint32_t)p0 - in.c;
                    UNROLL for (int j = 0; j < LS; j++) clk[j] = __int_as_float(max(d0 + j, 0));
                } else {
                    UNROLL for (int j = 0; j < LS; j++) {
                        const u64 at = p0 + (u64)j, c0 = (u64)(uint32_t)in.c;
                        const u64 d = at > c0 ? at - c0 : 0ull;
                        clk[j] = __int_as_float((int)(d < 0x7fffffffull ? d : 0x7fffffffull));
                    }
                }
                lslot_store(M, in.b, clk);
                continue;
            }
            case ST_SEG_SEL: {  // the running result is this piece; before its first sample the pieces before it hold
                const u64 p0 = ld64(M, in.a);
                const u64 c0 = (u64)(uint32_t)in.c;
                lacc_load(M, acc);
                if (p0 < c0) {  // (a tile wholly inside the piece keeps the running result as it is)
                    float before[LS];
                    lslot_load(M, in.b, before);
                    const int k = (int)(c0 - p0 < (u64)LS ? c0 - p0 : (u64)LS);  // samples of this tile in front of the piece
                    UNROLL for (int j = 0; j < LS; j++) acc[j] = j >= k ? acc[j] : before[j];
                }
                if (in.op & 0x100u) st64(M, in.a, p0 + (u64)LS);
                break;
            }
            case ST_TIME_CLK: {  // Time under a Reset: (local time) as f32 / sample_rate (generator.rs:101-111)
                const TimeDiv td = time_div_setup(P.sample_rate);
                float clk[LS];
                lslot_load(M, in.b, clk);
                const u64 pos = ld64(M, in.a);
                // Every local time of the tile below 2^24 (any run younger than 6 minutes): the short division of
                // time_div for all sixteen, the range looked at once instead of sample by sample.
                bool done = false;
                if (td.ok && pos < 16777216ull - (u64)LS) {
                    const uint32_t p32 = (uint32_t)pos;
                    uint32_t any = 0u;
                    UNROLL for (int j = 0; j < LS; j++) {
                        const int c = __float_as_int(clk[j]);
                        const uint32_t nloc = c >= 0 ? (uint32_t)c : p32 + (uint32_t)(-1 - c);
                        any |= nloc;
                        const float a = (float)nloc;
                        const float q0 = __fmul_rn(a, td.r);
                        acc[j] = fmaf(fmaf(-q0, td.b, a), td.r, q0);
                    }
                    done = any < 16777216u;
                }
                if (!done) {
                    UNROLL for (int j = 0; j < LS; j++) {
                        const int c = __float_as_int(clk[j]);
                        const u64 nloc = c >= 0 ? (u64)c : pos + (u64)(-1 - c);
                        acc[j] = time_div(nloc, td);
                    }
                }
                const int cl = __float_as_int(clk[LS - 1]);
                st64(M, in.a, cl >= 0 ? (u64)cl + 1ull : pos + (u64)LS);
                break;
            }
            case ST_SINE_CLK: {  // constant rate and phase under a Reset: the accumulator restarts with the clock
                float clk[LS];
                lslot_load(M, (int)(in.op >> 24), clk);
                const u64 a0 = ld64(M, in.a), inc = ld64(M, in.b), ph0 = ld64(M, in.c);
                const u64 a1 = a0 + ph0;
                UNROLL for (int j = 0; j < LS; j++) {
                    const int c = __float_as_int(clk[j]);
                    const u64 ph = inc * (u64)(uint32_t)(c >= 0 ? c : -1 - c) + (c >= 0 ? ph0 : a1);
                    acc[j] = !fast ? sin_turns_exact(ph) : (FASTMODE == 2 ? sin_p32((uint32_t)(ph >> 32)) : sin_hi<1>((int)(ph >> 32)));
                }
                const int cl = __float_as_int(clk[LS - 1]);
                st64(M, in.a, cl >= 0 ? inc * ((u64)cl + 1ull) : a0 + inc * (u64)LS);
                break;
            }
            case ST_BIN: {  // generator.rs:555-567 with both sides infinite
                float av[LS];
                lslot_load(M, in.a, av);
                lacc_load(M, acc);
                APPLY_OP_L((uint32_t)in.b, acc, av[j], acc[j])
                break;
            }
            case ST_SINE_CC: lane_sine_cc(M, acc, in.b + 2, (int)((in.op >> 8) & 0xffu)); break;
            case ST_SINE_AC: {
                float f[LS];
                lacc_load(M, f);
                const u64 ph0 = ld64(M, in.c);
                if (!fast) lane_sine_var<true, 0>(M, acc, f, ph0, in.a, sk);
                else lane_sine_var<true, FASTMODE>(M, acc, f, ph0, in.a, sk);
                break;
            }
            case ST_SINE_CA: {
                const u64 inc = ld64(M, in.b);
                lacc_load(M, acc);
                if (!fast) lane_sine_ca<0>(M, acc, inc, in.a, sk);
                else lane_sine_ca<FASTMODE>(M, acc, inc, in.a, sk);
                break;
            }
            case ST_SINE_AA: {
                float f[LS];
                lslot_load(M, in.b, f);
                lacc_load(M, acc);
                if (!fast) lane_sine_var<false, 0>(M, acc, f, 0ull, in.a, sk);
                else lane_sine_var<false, FASTMODE>(M, acc, f, 0ull, in.a, sk);
                break;
            }
            case ST_ALT_CC: {  // generator.rs:335-341
                const float cp = ldf(M, in.a), cn = ldf(M, in.b);
                lacc_load(M, acc);
                UNROLL for (int j = 0; j < LS; j++) acc[j] = acc[j] >= 0.0f ? cp : cn;
                break;
            }
            case ST_ALT: {
                float t[LS];
                lslot_load(M, in.a, t);
                if (in.c < 0) { const float c = ldf(M, ~in.c); UNROLL for (int j = 0; j < LS; j++) acc[j] = c; }
                else lacc_load(M, acc);
                if (in.b >= 0) {
                    float pv[LS];
                    lslot_load(M, in.b, pv);
                    UNROLL for (int j = 0; j < LS; j++) acc[j] = t[j] >= 0.0f ? pv[j] : acc[j];
                } else {
                    const float c = ldf(M, ~in.b);
                    UNROLL for (int j = 0; j < LS; j++) acc[j] = t[j] >= 0.0f ? c : acc[j];
                }
                break;
            }
            case LN_FM: {
                // Fused by lower.cpp: Sine(const) * m + c  ->  frequency of a Sine with constant phase
                // [ -> biquad ].  One dispatch, one basic block: the f64 angle additions, the phase
                // sums, the MUFU sines and the filter recurrence overlap in the pipeline.
                const tb_insn ex = nxt;
                ip += sizeof(tb_insn);
                nxt = lds_insn(ip);
                const float m = ldf(M, ex.a), c = ldf(M, ex.b);
                const u64 ph0 = ld64(M, in.c);  // phase offset of the carrier (TB_LN_FM_PHASE: its increment)
                float f[LS];
                lane_sine_cc(M, f, in.a, (int)((in.op >> 8) & 0xffu));
                {
                    const u64 mm = pk2(m, m), cc = pk2(c, c);
                    UNROLL for (int j = 0; j < LS; j += 2) unpk2(add2(mul2(pk2(f[j], f[j + 1]), mm), cc), f[j], f[j + 1]);
                }
                const bool fastfm = ((in.op >> 24) & 0x3fu) == TB_SINE_FAST;
                if ((in.op >> 24) & TB_LN_FM_PHASE) {  // phase modulation: constant increment, f holds the phases
                    UNROLL for (int j = 0; j < LS; j++) acc[j] = f[j];
                    if (!fastfm) lane_sine_ca<0>(M, acc, ph0, in.b, sk);
                    else lane_sine_ca<FASTMODE>(M, acc, ph0, in.b, sk);
                } else if (fabsf(m) + fabsf(c) < sk.flimit) {  // |f| <= |m| + |c|: no per-sample range test
                    if (!fastfm) lane_sine_var<true, 0, false>(M, acc, f, ph0, in.b, sk);
                    else lane_sine_var<true, FASTMODE, false>(M, acc, f, ph0, in.b, sk);
                } else {
                    if (!fastfm) lane_sine_var<true, 0>(M, acc, f, ph0, in.b, sk);
                    else lane_sine_var<true, FASTMODE>(M, acc, f, ph0, in.b, sk);
                }
                if (ex.c >= 0) lane_filter<3, 2>(M, acc, ex.c, (int)ex.op, 3, 2);
                break;
            }
            case ST_FILT: {
                const int K = (int)((in.op >> 8) & 0xfu), J = (int)((in.op >> 12) & 0x7u);
                lacc_load(M, acc);
                if (K == 3 && J == 2) lane_filter<3, 2>(M, acc, in.a, in.b, K, J);  // every filter of lib/v0/std.tuun
                else if (K == 1 && J == 1) lane_filter<1, 1>(M, acc, in.a, in.b, K, J);
                else if (K == 2 && J == 1) lane_filter<2, 1>(M, acc, in.a, in.b, K, J);
                else lane_filter<-1, -1>(M, acc, in.a, in.b, K, J);
                break;
            }
            default: return;  // unreachable: lower.cpp emits only the words above
        }
        // Post-op words: constant point operators (generator.rs:541-548) and what lower.cpp folded behind the
        // producer.  Runs of ST_AFFINE — most of them — have a loop of their own: with one path through its body the
        // sixteen values stay where they are (in one loop over all kinds every word cost sixteen register moves).
        for (uint32_t np = (in.op >> 16) & 0xffu; np > 0; np--) {
            while ((nxt.op & 0xffu) == ST_AFFINE) {  // (acc * m) + a, both rounded
                const float m = ldf(M, nxt.b), a = ldf(M, nxt.c);
                ip += sizeof(tb_insn);
                nxt = lds_insn(ip);
                const u64 mm = pk2(m, m), aa = pk2(a, a);
                UNROLL for (int j = 0; j < LS; j += 2) unpk2(add2(mul2(pk2(acc[j], acc[j + 1]), mm), aa), acc[j], acc[j + 1]);
                if (--np == 0) break;
            }
            if (np == 0) break;
            const tb_insn po = nxt;
            ip += sizeof(tb_insn);
            nxt = lds_insn(ip);
            const uint32_t pk = po.op & 0xffu;
            if (pk == ST_ALT_CC) {  // folded by lower.cpp fold_lane_postops (generator.rs:335-341)
                const float cp = ldf(M, po.a), cn = ldf(M, po.b);
                UNROLL for (int j = 0; j < LS; j++) acc[j] = acc[j] >= 0.0f ? cp : cn;
            } else if (pk == ST_FILT) {    // folded biquad
                lane_filter<3, 2>(M, acc, po.a, po.b, 3, 2);
            } else if (pk == ST_SAVE) {    // folded: the running result into a slot
                lslot_store(M, po.a, acc);
            } else if (pk == ST_BIN) {     // folded: slot (operator) running result
                float av[LS];
                lslot_load(M, po.a, av);
                APPLY_OP_L((uint32_t)po.b, acc, av[j], acc[j])
            } else {
                const float c = ldf(M, po.b);
                APPLY_OP_L((uint32_t)po.a, acc, acc[j], c)
            }
        }
        lacc_store(M, acc);
    }
}

// Per-voice setup by the owning thread: constant table (is_const folding, generator.rs:574-612),
// then the derived constants of the lane plan.  The state block is already in W.
__device__ void setup_lane(const tb_launch& P, const LaneMem& M, const float* prow) {
    for (uint32_t k = 0; k < P.n_cval; k++) {
        const tb_cexpr e = P.cexpr[k];
        float v;
        if (e.kind == CE_LIT) v = e.value;
        else if (e.kind == CE_PARAM) v = prow ? prow[e.a] : e.value;
        else if (e.kind == CE_NEG) v = -ldf(M, e.a);
        else v = apply1(e.op, ldf(M, e.a), ldf(M, e.b));
        stf(M, (int)k, v);
    }
    for (uint32_t t = 0; t < P.n_lane_aux; t++) {
        const tb_lane_aux a = P.lane_aux[t];
        if (a.kind == LA_INC || a.kind == LA_ROT) {
            const u64 inc = turns_to_fx_slow((double)ldf(M, a.a) / (TB_TAU * (double)P.sample_rate));
            st64(M, (int)a.w_off, inc);
            if (a.kind == LA_ROT) {
                double2* rot = reinterpret_cast<double2*>(M.Q + (size_t)a.q_off * LT);
                for (int k = 1; k <= LS / 2 + 1; k++) {  // k = 1..8, then the tile step 16
                    const u64 ang = inc * (u64)(k <= LS / 2 ? k : LS);
                    double2 r;
                    r.x = sin_turns_d8(ang + 0x4000000000000000ull);
                    r.y = sin_turns_d8(ang);
                    rot[(size_t)(k - 1) * LT] = r;
                }
                // (sin, cos) at the centre of the first tile, from the exact accumulator and phase.
                const u64 pc = ld64(M, a.b) + turns_to_fx_slow((double)ldf(M, a.c) / TB_TAU) + inc * (u64)(LS / 2);
                std_(M, (int)a.w_off + 2, sin_turns_d8(pc));
                std_(M, (int)a.w_off + 4, sin_turns_d8(pc + 0x4000000000000000ull));
            }
        } else if (a.kind == LA_PHASE) {
            st64(M, (int)a.w_off, turns_to_fx_slow((double)ldf(M, a.a) / TB_TAU));
        } else if (a.kind == LA_TL_POS) {  // a timeline starts with the root Fin: its position is that clock's
            st64(M, (int)a.w_off, ld64(M, a.a));
        } else if (a.kind == LA_TL_PIECE || a.kind == LA_TL_ZERO) {
        } else {  // LA_COEF
            const tb_filter_tab* ft = &P.filt[a.a];
            for (uint32_t e = 0; e < ft->K + ft->J; e++) stw(M, (int)(a.w_off + e), ldw(M, ~ft->coef[e]));
        }
    }
}
// When the kernel ends: the accumulators of the constant sines advance by the samples rendered.
__device__ void finish_lane(const tb_launch& P, const LaneMem& M, u64 n_samples) {
    for (uint32_t t = 0; t < P.n_lane_aux; t++) {
        const tb_lane_aux a = P.lane_aux[t];
        if (a.kind == LA_ROT) st64(M, a.b, ld64(M, a.b) + ld64(M, (int)a.w_off) * n_samples);
        // A timeline (ST_SEG_*): the state the general interpreter would hold at this position (generator.rs:133-188).
        // A piece that has begun: its Fin's clock counts from its first sample (the clocked words of the piece left
        // theirs), the Append in front of the next piece says whether that one has begun; a piece that has rendered
        // no sample yet is Initial (its clocked words ran on a clock held at 0).
        if (a.kind == LA_TL_PIECE) {
            const u64 pos = ld64(M, (int)a.w_off), c0 = (u64)(uint32_t)a.a;
            if (pos >= c0) {
                if (a.b >= 0) st64(M, a.b, pos - c0);
                if (a.c >= 0) stw(M, a.c, pos >= (u64)a.q_off ? 1u : 0u);
            }
        } else if (a.kind == LA_TL_ZERO) {
            if (ld64(M, (int)a.w_off) <= (u64)(uint32_t)a.a)
                for (int k = 0; k < a.c; k++) stw(M, a.b + k, 0u);
        }
    }
}


// ---- the row stores --------------------------------------------------------------------------------
// After every second tile the 32 samples x 32 voices of the warp leave: lane l takes chunk (l & 7) of
// rows (l >> 3) + 4 i, i < 8, so eight lanes write the 128 contiguous bytes a row has gathered
// (half the requests and address translations per byte of a store per tile: 65,536 rows are open
// at once).  Both __syncwarp()s belong to the protocol: owners have written before the first,
// readers have read before the owners write again.  An odd last tile leaves as 64 bytes per row.
struct RowStore {
    float* out;            // first sample of this launch, row 0
    size_t stride;         // floats between rows
    const float4* tbase;   // the warp's first row in the accumulator buffer
    uint32_t v0, n_voices; // first voice of the warp
    size_t off;            // samples already stored
    bool fast, vec_ok;
    float* mix;            // mixdown without rows (tb_launch::mix_partial): this warp's row of partial sums
    float* dfast;          // fast path: this lane's 16 bytes of its first row at the current offset, and 4 rows further on:
    size_t step4;          // kept in registers so that the loop neither re-derives them nor reloads the stride (LDC)
    const unsigned long long* rowoff;  // time-axis split (tb_launch::vsplit*): rows are segments of the real voices' rows,
                                       // whose first samples (offsets from `out`, in floats) sit in a table, one per row of the warp
};
// First sample (of this launch) of virtual voice vv's row.
__device__ __forceinline__ float* row_of(const RowStore& R, uint32_t vv) {
    if (!TB_LANES_VSPLIT) return R.out + (size_t)vv * R.stride;
    return R.out + R.rowoff[vv - R.v0];
}
#ifndef TB_ST
#define TB_ST 2
#endif
__device__ __forceinline__ void st_row(float4* d, const float4& v) {
#if TB_ST == 0
    __stcs(d, v);
#elif TB_ST == 1
    *d = v;
#elif TB_ST == 2
    __stcg(d, v);
#else
    __stwt(d, v);
#endif
}
__device__ __forceinline__ void put4(float* d, const float4& v, bool vec_ok) {
    if (vec_ok) __stcs(reinterpret_cast<float4*>(d), v);
    else { d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
}
__device__ __forceinline__ void store_pair(RowStore& R, int l) {
    __syncwarp();
    float4 v[8];
    const float4* src = R.tbase + (l & 7) * AS + (l >> 3);
    UNROLL for (int i = 0; i < 8; i++) v[i] = src[4 * i];
    __syncwarp();
    float* d = R.out + (size_t)(R.v0 + (l >> 3)) * R.stride + R.off + (size_t)(l & 7) * 4;
    const size_t step = 4 * R.stride;
    if (TB_LANES_VSPLIT) {  // segments of real voices' rows (abi.cpp split_pass)
        if (R.fast) {  // all 32 rows of the warp exist and are 16-byte aligned
            const unsigned long long* ro = R.rowoff + (l >> 3);
            float* base = R.out + R.off + (size_t)(l & 7) * 4;
            UNROLL for (int i = 0; i < 8; i++) st_row(reinterpret_cast<float4*>(base + ro[4 * i]), v[i]);
        } else if (R.out) {
            UNROLL for (int i = 0; i < 8; i++) {
                const uint32_t r = (uint32_t)(l >> 3) + 4u * i;
                if (R.v0 + r < R.n_voices) {
                    float* q = R.out + R.rowoff[r] + R.off + (size_t)(l & 7) * 4;
                    if (R.vec_ok) st_row(reinterpret_cast<float4*>(q), v[i]);
                    else put4(q, v[i], false);
                }
            }
        }
    } else if (R.fast) {  // all 32 rows of the warp exist and are 16-byte aligned
        float* q = R.dfast;
        UNROLL for (int i = 0; i < 8; i++) {
            st_row(reinterpret_cast<float4*>(q), v[i]);
            q += R.step4;
        }
        R.dfast += 2 * LS;
    } else if (R.out) {
        UNROLL for (int i = 0; i < 8; i++)
            if (R.v0 + (uint32_t)(l >> 3) + 4u * i < R.n_voices) put4(d + i * step, v[i], R.vec_ok);
    }
    R.off += 2 * LS;
}
__device__ __forceinline__ void store_single(RowStore& R, int l, int half) {
    __syncwarp();
    float4 v[4];
    const float4* src = R.tbase + (4 * half + (l & 3)) * AS + (l >> 2);
    UNROLL for (int i = 0; i < 4; i++) v[i] = src[8 * i];
    __syncwarp();
    if (R.out) {
        UNROLL for (int i = 0; i < 4; i++) {
            const uint32_t vv = R.v0 + (uint32_t)(l >> 2) + 8u * i;
            if (vv < R.n_voices) put4(row_of(R, vv) + R.off + (size_t)(l & 3) * 4, v[i], R.vec_ok);
        }
    }
    R.off += LS;
}
// Mixdown without rows (tb_render_mix with TB_NO_VOICE_OUT): instead of leaving, every pair of tiles is summed over
// the warp's 32 voices on the chip — lane l adds sample l of the pair (32 samples) over voices 0 .. 31 IN VOICE
// ORDER (with the chunk stride of AS units the 32 lanes of a load touch 32 different banks) — and 128 bytes of
// partial sums go to the warp's row of tb_launch::mix_partial.  tb_mix_kernel then adds the rows of all warps in
// warp order: the tracker's serial `out[j] += tmp[j]` (tracker.rs:617-619) re-associated in blocks of 32 voices.
__device__ __forceinline__ void mix_pair(RowStore& R, int l) {
    __syncwarp();
    const int half = l >> 4, smp = l & 15;
    const float* a = reinterpret_cast<const float*>(R.tbase + (4 * half + (smp >> 2)) * AS) + (smp & 3);
    float s = a[0];
    UNROLL for (int v = 1; v < 32; v++) s = __fadd_rn(s, a[4 * v]);
    __syncwarp();
    R.mix[R.off + l] = s;
    R.off += 2 * LS;
}
// An odd last tile (in half `half` of the buffer): the same sums, sixteen lanes at work.
__device__ __forceinline__ void mix_tile(RowStore& R, int l, int half) {
    __syncwarp();
    const int smp = l & 15;
    const float* a = reinterpret_cast<const float*>(R.tbase + (4 * half + (smp >> 2)) * AS) + (smp & 3);
    float s = a[0];
    UNROLL for (int v = 1; v < 32; v++) s = __fadd_rn(s, a[4 * v]);
    __syncwarp();
    if (l < 16) R.mix[R.off + l] = s;
    R.off += LS;
}
// Tile number t (from 0) of the launch has just been written to its half of the buffer.
template <bool MIX>
__device__ __forceinline__ void tile_done(RowStore& R, int l, u64 t, u64 n_tiles) {
    if (t & 1) {
        if (MIX) mix_pair(R, l);
        else store_pair(R, l);
    } else if (t + 1 == n_tiles) {
        if (MIX) mix_tile(R, l, 0);
        else store_single(R, l, 0);
    }
}

// ---- a program that is ONE fused FM voice ----------------------------------------------------------
// When the whole lane program is a single LN_FM (program.h) with a FAST carrier — the shape of the
// 65,536-voice FM + low-pass batch — there is nothing to dispatch: the thread keeps the voice's
// state in registers for the whole launch (carried sin/cos pair, 32-bit carrier phase, filter
// history and its pending products) and runs tiles in a software-pipelined loop: the carrier tile
// of step t is computed in the same basic block as the filter recurrence over the carrier tile of
// step t-1, so the serial y[n] chain overlaps the f64 angle additions, conversions and MUFU sines.
#ifndef TB_ABL
#define TB_ABL 0
#endif
#if TB_ABL == 3
#define TB_D2F(x) __int_as_float(__double2hiint(x))
#else
#define TB_D2F(x) ((float)(x))
#endif
// SLOW: some voice of the warp has |m| + |c| beyond TB_FM_TURNS turns a sample (nothing audible: the
// Nyquist rate is half a turn): every increment goes through the integer conversion instead.
//   p: the carrier phase in 2^-44 turns (bits above 43 are ignored and may hold anything).
//   CAP: also report in p_cap the phase before sample `cap` (the phase after `cap` samples of the tile).
// The modulator's rotations held in REGISTERS for the whole launch (TB_FM_ROT_REGS, the default): the tile
// loop's loads of the 9-entry table from shared memory queue behind the conversions and MUFU sines in the
// memory-I/O pipe, and the products that consume them were the largest single stall of the loop (26 % of its
// samples, ncu source view, profiles/r2_*).  Five entries suffice when the tile is built in two levels: the
// pair (sin, cos) at samples 4 and 12 by a rotation of -+4 steps from the centre, then every other sample as
// (sub-centre) +- k steps, k = 1..3, sample 0 as sample 4 - 4 steps: 30 f64 operations a tile against 27,
// no loads.  Rounding errors of the extra level are ~1e-16, as before far below the f32 result's 6e-8.
#ifndef TB_FM_ROT_REGS
#define TB_FM_ROT_REGS 1
#endif
#ifndef TB_FM_CHEB
#define TB_FM_CHEB 1
#endif
struct FmRot {
    double c1, s1, c2, s2, c3, s3, c4, s4, c16, s16;
};
__device__ __forceinline__ FmRot fm_rot_load(const double2* rot) {
    FmRot r;
    double2 t;
    t = rot[(size_t)0 * LT]; r.c1 = t.x; r.s1 = t.y;
    t = rot[(size_t)1 * LT]; r.c2 = t.x; r.s2 = t.y;
    t = rot[(size_t)2 * LT]; r.c3 = t.x; r.s3 = t.y;
    t = rot[(size_t)3 * LT]; r.c4 = t.x; r.s4 = t.y;
    t = rot[(size_t)(LS / 2) * LT]; r.c16 = t.x; r.s16 = t.y;
    return r;
}
//   PHASE_ONLY: the summary pass of a time-axis split (run_fm_sums) wants the phase after the tile, not its sines.
//   RAW: car[] receives the bit patterns 1.m of the sines' arguments (pd_m23) instead of the sines — the producer
//   warp of the warp-specialised kernel (lanes_fm_ws.cu) hands those to the warp that owns MUFU and the filter.
//   FMODE (what the whole warp's voices allow, lanes_fm_ws.cuh): 1 = the carrier frequency is never negative
//   (c >= |m|: any FM patch of moderate index), so its widening to f64 by integer instructions drops the two sign
//   operations; 2 = no modulation at all (m == 0: the frequency IS c, whatever the modulator does), so the tile needs
//   neither the modulator nor the affine map.  Same bits as FMODE 0 in both cases.
template <bool SLOW, bool CAP = false, bool PHASE_ONLY = false, bool RAW = false, int FMODE = 0>
__device__ __forceinline__ void fm_carrier_tile(float (&car)[LS], double& S, double& Cq, const double2* rot, const FmRot& rr,
                                                u64 mm, u64 cc, u64& p, const SineK& sk, int cap = 0, u64* p_cap = nullptr) {
    float f[LS];
    if (FMODE == 2) {
        float c0, c1;
        unpk2(cc, c0, c1);
        UNROLL for (int j = 0; j < LS; j++) f[j] = c0;  // fl(fl(s * 0) + c) = c for every finite s
    } else {
#if TB_ABL == 4
    UNROLL for (int j = 0; j < LS; j++) f[j] = TB_D2F(S) + j;
#elif TB_FM_ROT_REGS && TB_FM_CHEB
    (void)rot;
    {
        // The tile from its centre outwards by the three-term recurrence s[k + 1] = 2 cos(d) s[k] - s[k - 1], both ways:
        // sin at +-1 by angle addition (3 operations), then one DFMA a sample — 16 f64 operations a tile against 30 for
        // the two-level angle additions, and 3 entries of the rotation table in registers against 5.  Rounding errors
        // grow like k^2 ulp over the k <= 8 steps of a chain (the recurrence is a double integrator at worst, for a
        // modulator near 0 Hz): 1e-14, against the 6e-8 of the f32 the value is rounded to; the centre itself moves by
        // an angle addition as before, so nothing carries from tile to tile.
        f[8] = TB_D2F(S);
        const double k2 = rr.c1 + rr.c1;
        const double b = Cq * rr.s1;
        double up1 = fma(S, rr.c1, b), dn1 = fma(S, rr.c1, -b);   // sin at samples 9 and 7
        f[9] = TB_D2F(up1);
        f[7] = TB_D2F(dn1);
        double up0 = S, dn0 = S;
        UNROLL for (int k = 2; k <= 7; k++) {
            const double u = fma(k2, up1, -up0);
            up0 = up1; up1 = u;
            f[8 + k] = TB_D2F(u);
        }
        UNROLL for (int k = 2; k <= 8; k++) {
            const double v = fma(k2, dn1, -dn0);
            dn0 = dn1; dn1 = v;
            f[8 - k] = TB_D2F(v);
        }
    }
#elif TB_FM_ROT_REGS
    (void)rot;
    {
        f[8] = TB_D2F(S);
        const double cs4 = Cq * rr.s4, ss4 = S * rr.s4;
        const double Sa = fma(S, rr.c4, -cs4), Sb = fma(S, rr.c4, cs4);     // sin at samples 4 and 12
        const double Ca = fma(Cq, rr.c4, ss4), Cb = fma(Cq, rr.c4, -ss4);   // cos there
        f[4] = TB_D2F(Sa);
        f[12] = TB_D2F(Sb);
        double b;
        b = Ca * rr.s1; f[5] = TB_D2F(fma(Sa, rr.c1, b)); f[3] = TB_D2F(fma(Sa, rr.c1, -b));
        b = Cb * rr.s1; f[13] = TB_D2F(fma(Sb, rr.c1, b)); f[11] = TB_D2F(fma(Sb, rr.c1, -b));
        b = Ca * rr.s2; f[6] = TB_D2F(fma(Sa, rr.c2, b)); f[2] = TB_D2F(fma(Sa, rr.c2, -b));
        b = Cb * rr.s2; f[14] = TB_D2F(fma(Sb, rr.c2, b)); f[10] = TB_D2F(fma(Sb, rr.c2, -b));
        b = Ca * rr.s3; f[7] = TB_D2F(fma(Sa, rr.c3, b)); f[1] = TB_D2F(fma(Sa, rr.c3, -b));
        b = Cb * rr.s3; f[15] = TB_D2F(fma(Sb, rr.c3, b)); f[9] = TB_D2F(fma(Sb, rr.c3, -b));
        f[0] = TB_D2F(fma(Sa, rr.c4, -(Ca * rr.s4)));
    }
#else
    (void)rr;
    f[LS / 2] = TB_D2F(S);
    UNROLL for (int k = 1; k < LS / 2; k++) {
        const double2 r = rot[(size_t)(k - 1) * LT];
        const double b = Cq * r.y;
        f[LS / 2 + k] = TB_D2F(fma(S, r.x, b));
        f[LS / 2 - k] = TB_D2F(fma(S, r.x, -b));
    }
    {
        const double2 r = rot[(size_t)(LS / 2 - 1) * LT];
        f[0] = TB_D2F(fma(S, r.x, -(Cq * r.y)));
    }
#endif
#if TB_FM_ROT_REGS
    const double S2 = fma(S, rr.c16, Cq * rr.s16);
    Cq = fma(Cq, rr.c16, -(S * rr.s16));
#else
    const double2 r16 = rot[(size_t)(LS / 2) * LT];
    const double S2 = fma(S, r16.x, Cq * r16.y);
    Cq = fma(Cq, r16.x, -(S * r16.y));
#endif
    S = S2;
    UNROLL for (int j = 0; j < LS; j += 2) unpk2(add2(mul2(pk2(f[j], f[j + 1]), mm), cc), f[j], f[j + 1]);
    }
    if (SLOW) {
        UNROLL for (int j = 0; j < LS; j += 2) {
            const u64 t0 = p;
            p += freq_to_inc(f[j], sk) >> 20;
            const u64 t1 = p;
            p += freq_to_inc(f[j + 1], sk) >> 20;
            if (CAP && j == cap) *p_cap = t0;
            if (CAP && j + 1 == cap) *p_cap = t1;
            if (RAW && j >= 4 * TB_WS_SINES) raw_pair(p44_m23(t0), p44_m23(t1), car[j], car[j + 1]);
            else if (!PHASE_ONLY) sin_m23x2(p44_m23(t0), p44_m23(t1), car[j], car[j + 1]);
        }
        return;
    }
    if (PHASE_ONLY) {
        // Only the phase after the tile is wanted: the 16 increments are rounded to the grid independently (the same
        // integers the running DFMA below adds: the sum of an integer and f * kscale rounds to that integer plus
        // rint(f * kscale); a tie would need f * kscale to be a half-integer exactly, which its 77-bit product with
        // 2^44 / (2 pi sample_rate) never is) and added as integers — no serial chain of 16 DFMAs.
        // Four running sums in the mantissas of doubles at 1.5 * 2^52 (pd_make), four increments each: a DFMA adds
        // rint(f * kscale) to an integer-valued sum whatever the order, so the tile costs 16 DFMAs and a handful of
        // integer adds instead of 16 DFMAs and 16 64-bit adds.
        // (f as a double by integer instructions: this pass is bound by the conversion unit, the ALU is idle)
        double a0 = pd_make(0), a1 = a0, a2 = a0, a3 = a0;
        UNROLL for (int j = 0; j < LS; j += 4) {
            a0 = fma(f32_to_f64_alu(f[j]), sk.kscale, a0);
            a1 = fma(f32_to_f64_alu(f[j + 1]), sk.kscale, a1);
            a2 = fma(f32_to_f64_alu(f[j + 2]), sk.kscale, a2);
            a3 = fma(f32_to_f64_alu(f[j + 3]), sk.kscale, a3);
        }
        const u64 sum = (pd_bits(a0) + pd_bits(a1)) + (pd_bits(a2) + pd_bits(a3));
        p += sum;  // the magic bits above the 44 phase bits are dropped by whoever reads p
        return;
    }
    // One DFMA per sample adds f * kscale to the running phase and rounds the sum to 2^-44 turns (pd_make).
    double Pd = pd_make(p);
    UNROLL for (int j = 0; j < LS; j += 2) {
        const double T0 = Pd;
        Pd = fma(FMODE == 1 ? f32_to_f64_pos(f[j]) : TB_WIDEN(f[j]), sk.kscale, Pd);
        const double T1 = Pd;
        Pd = fma(FMODE == 1 ? f32_to_f64_pos(f[j + 1]) : TB_WIDEN(f[j + 1]), sk.kscale, Pd);
        if (CAP && j == cap) *p_cap = pd_bits(T0);
        if (CAP && j + 1 == cap) *p_cap = pd_bits(T1);
#if TB_ABL == 2
        car[j] = __uint_as_float(pd_m23(T0)); car[j + 1] = __uint_as_float(pd_m23(T1));
#else
        if (RAW && j >= 4 * TB_WS_SINES) raw_pair(pd_m23(T0, sk.one23), pd_m23(T1, sk.one23), car[j], car[j + 1]);
        else if (!PHASE_ONLY) sin_m23x2(pd_m23(T0, sk.one23), pd_m23(T1, sk.one23), car[j], car[j + 1]);
#endif
    }
    p = pd_bits(Pd);
}
struct BiquadRegs {
    float b0, b1, b2, a1, a2;
    float x1, x2;      // x[-1], x[-2]
    float p1, p2, q2;  // a1 y[-1], a2 y[-2], a2 y[-1]
    float y1, y2;      // y[-1], y[-2] (for the state block)
};
// generator.rs:496-507 for K = 3, J = 2, operation order and roundings of lane_filter<3, 2>.
__device__ __forceinline__ void biquad_tile(float (&y)[LS], const float (&x)[LS], BiquadRegs& F) {
    const u64 bb0 = pk2(F.b0, F.b0), bb1 = pk2(F.b1, F.b1), bb2 = pk2(F.b2, F.b2), aa = pk2(F.a1, F.a2);
    float pr1[LS + 1], pr2[LS + 2];  // pr1[1 + i] = b1 x[i], pr2[2 + i] = b2 x[i]
    pr1[0] = __fmul_rn(F.b1, F.x1);
    pr2[0] = __fmul_rn(F.b2, F.x2);
    pr2[1] = __fmul_rn(F.b2, F.x1);
    UNROLL for (int i = 0; i < LS; i += 2) {
        const u64 xx = pk2(x[i], x[i + 1]);
        unpk2(mul2(xx, bb1), pr1[1 + i], pr1[2 + i]);
        if (i + 2 < LS) unpk2(mul2(xx, bb2), pr2[2 + i], pr2[3 + i]);
    }
    UNROLL for (int i = 0; i < LS; i += 2) {
        float s0, s1;
        unpk2(mul2(pk2(x[i], x[i + 1]), bb0), s0, s1);
        s0 = __fadd_rn(s0, pr1[i]);
        s1 = __fadd_rn(s1, pr1[i + 1]);
        unpk2(add2(pk2(s0, s1), pk2(pr2[i], pr2[i + 1])), s0, s1);
        const float y0 = __fsub_rn(__fsub_rn(s0, F.p1), F.p2);
        float n1, n2;
        unpk2(mul2(aa, pk2(y0, y0)), n1, n2);        // a1 y0, a2 y0
        const float y1 = __fsub_rn(__fsub_rn(s1, n1), F.q2);
        F.p2 = n2;
        unpk2(mul2(aa, pk2(y1, y1)), F.p1, F.q2);    // a1 y1, a2 y1
        y[i] = y0;
        y[i + 1] = y1;
    }
    F.x1 = x[LS - 1];
    F.x2 = x[LS - 2];
    F.y1 = y[LS - 1];
    F.y2 = y[LS - 2];
}

// What run_fm_voice needs to know about the modulator beyond the LN_FM words (from its LA_ROT entry).
struct FmPrime {
    int acc_w;     // W index of the modulator's 64-bit accumulator
    int ph_cval;   // cval of its phase offset
    bool fresh;    // this voice's stream starts here: the filter has never run (state block all zero)
};
// The last `rem` < 16 samples of a launch: the tile is in the accumulator half `half`; rows get only
// their first `rem` samples (the next row starts right behind).
__device__ __forceinline__ void store_partial(RowStore& R, int l, int half, int rem) {
    __syncwarp();
    float4 v[4];
    const float4* src = R.tbase + (4 * half + (l & 3)) * AS + (l >> 2);
    UNROLL for (int i = 0; i < 4; i++) v[i] = src[8 * i];
    __syncwarp();
    if (R.out) {
        const int e0 = (l & 3) * 4;
        UNROLL for (int i = 0; i < 4; i++) {
            const uint32_t vv = R.v0 + (uint32_t)(l >> 2) + 8u * i;
            if (vv < R.n_voices) {
                float* q = row_of(R, vv) + R.off + (size_t)(l & 3) * 4;
                if (e0 + 0 < rem) q[0] = v[i].x;
                if (e0 + 1 < rem) q[1] = v[i].y;
                if (e0 + 2 < rem) q[2] = v[i].z;
                if (e0 + 3 < rem) q[3] = v[i].w;
            }
        }
    }
    R.off += rem;
}

// n_tiles whole tiles, then `rem` < 16 more samples (a last tile of which only the first `rem` samples
// count: its state is taken where they end).  A voice whose stream starts here (FmPrime::fresh) first does
// what the filter's first call does in the reference — reads K - 1 = 2 carrier samples ahead and starts
// from zero outputs (generator.rs:234-252) — so a call needs no launch of the general kernel at all.
template <bool TAIL, bool MIX, bool SLOW>
__device__ __forceinline__ void run_fm_voice(const tb_insn* code, LaneMem& M, const SineK& sk, RowStore& R,
                                             bool active, int l, u64 n_tiles, int rem, const FmPrime& prime) {
    float4* const abase = M.A;
    const tb_insn w0 = code[0], w1 = code[1];
    double S = 0.0, Cq = 1.0;
    const double2* rot = reinterpret_cast<const double2*>(M.Q + (size_t)((w0.op >> 8) & 0xffu) * LT);
    u64 mm = 0, cc = 0;
    u64 p = 0, p_start = 0;  // carrier phase, 2^-44 turns
    BiquadRegs F = {};
    const double ks = sk.kscale;
    if (active) {
        S = ldd(M, w0.a);
        Cq = ldd(M, w0.a + 2);
        const float m = ldf(M, w1.a), c = ldf(M, w1.b);
        mm = pk2(m, m);
        cc = pk2(c, c);
        p_start = p = (ld64(M, w0.b) + ld64(M, w0.c)) >> 20;
        if (TAIL) {
            const int wc = (int)w1.op, st = w1.c;
            F.b0 = ldf(M, wc); F.b1 = ldf(M, wc + 1); F.b2 = ldf(M, wc + 2);
            F.a1 = ldf(M, wc + 3); F.a2 = ldf(M, wc + 4);
            if (prime.fresh) {
                // The two carrier samples the filter reads ahead: the modulator by the exact polynomial,
                // the affine map and the phase steps as in a tile.
                const u64 inc = ld64(M, w0.a - 2);
                const u64 am = ld64(M, prime.acc_w);
                const u64 phm = turns_to_fx_slow((double)ldf(M, prime.ph_cval) / TB_TAU);
                const float f0 = __fadd_rn(__fmul_rn((float)sin_turns_d8(am + phm), m), c);
                const float f1 = __fadd_rn(__fmul_rn((float)sin_turns_d8(am + phm + inc), m), c);
                F.x2 = sin_m23(p44_m23(p));
                p += SLOW ? freq_to_inc(f0, sk) >> 20 : magic_raw(f0, ks);
                F.x1 = sin_m23(p44_m23(p));
                p += SLOW ? freq_to_inc(f1, sk) >> 20 : magic_raw(f1, ks);
                F.y1 = F.y2 = 0.0f;
                const u64 am2 = am + 2ull * inc;
                st64(M, prime.acc_w, am2);
                const u64 pc = am2 + phm + inc * (u64)(LS / 2);
                S = sin_turns_d8(pc);
                Cq = sin_turns_d8(pc + 0x4000000000000000ull);
                stw(M, st, 1u);      // initialised,
                stw(M, st + 1, 2u);  // K - 1 inputs held
            } else {
                F.x2 = ldf(M, st + 2); F.x1 = ldf(M, st + 3);
                F.y2 = ldf(M, st + 4); F.y1 = ldf(M, st + 5);
            }
            F.p1 = __fmul_rn(F.a1, F.y1);
            F.q2 = __fmul_rn(F.a2, F.y1);
            F.p2 = __fmul_rn(F.a2, F.y2);
        }
    }
    FmRot rr = {};
    if (active) rr = fm_rot_load(rot);
    float car[LS];
    UNROLL for (int j = 0; j < LS; j++) car[j] = 0.0f;
    if (!TAIL) {
        for (u64 t = 0; t < n_tiles; t++) {
            if (active) {
                fm_carrier_tile<SLOW>(car, S, Cq, rot, rr, mm, cc, p, sk);
                M.A = abase + (t & 1) * 4 * AS;
                lacc_store(M, car);
            }
            tile_done<MIX>(R, l, t, n_tiles);
        }
    } else if (n_tiles > 0) {
        if (active) fm_carrier_tile<SLOW>(car, S, Cq, rot, rr, mm, cc, p, sk);
        // (a 32-bit down-counter and a parity bit: with `t < n_tiles` as the condition the compiler re-derives n_tiles
        // from the launch parameters in constant memory every iteration — an LDC whose latency showed as 5 % of the
        // loop's stall samples)
        uint32_t odd = 0;  // parity of the tile that leaves
        for (uint32_t left = (uint32_t)(n_tiles - 1); left != 0; left--) {
            if (active) {
                float y[LS], nxt[LS];
#if TB_ABL == 1
                UNROLL for (int j = 0; j < LS; j++) y[j] = car[j];
#else
                biquad_tile(y, car, F);                                   // tile t-1 leaves ...
#endif
                fm_carrier_tile<SLOW>(nxt, S, Cq, rot, rr, mm, cc, p, sk);    // ... while tile t is made
                M.A = abase + odd * 4 * AS;
                lacc_store(M, y);
                UNROLL for (int j = 0; j < LS; j++) car[j] = nxt[j];
            }
#if TB_ABL != 5
            if (odd) {
                if (MIX) mix_pair(R, l);
                else store_pair(R, l);
            }
#endif
            odd ^= 1u;
        }
        if (active) {
            float y[LS];
            biquad_tile(y, car, F);
            M.A = abase + ((n_tiles - 1) & 1) * 4 * AS;
            lacc_store(M, y);
        }
        tile_done<MIX>(R, l, n_tiles - 1, n_tiles);
    }
    if (rem > 0) {  // the samples that do not fill a tile
        const int half = (int)(n_tiles & 1);
        if (active) {
            u64 p_rem = p;
            fm_carrier_tile<SLOW, true>(car, S, Cq, rot, rr, mm, cc, p, sk, rem, &p_rem);
            p = p_rem;
            float y[LS];
            if (TAIL) {
                float ex[LS + 2], ey[LS + 2];  // history ++ tile, to take the state where the `rem` samples end
                ex[0] = F.x2; ex[1] = F.x1;
                ey[0] = F.y2; ey[1] = F.y1;
                biquad_tile(y, car, F);
                UNROLL for (int j = 0; j < LS; j++) { ex[2 + j] = car[j]; ey[2 + j] = y[j]; }
                UNROLL for (int j = 0; j < LS; j++) {
                    if (j == rem) { F.x2 = ex[j]; F.x1 = ex[j + 1]; F.y2 = ey[j]; F.y1 = ey[j + 1]; }
                }
            } else {
                UNROLL for (int j = 0; j < LS; j++) y[j] = car[j];
            }
            M.A = abase + half * 4 * AS;
            lacc_store(M, y);
        }
        // (mixdown: a whole tile of sums is written, of which the caller reads the first `rem`: the partial rows are
        // padded to whole tiles)
        if (MIX) mix_tile(R, l, half);
        else store_partial(R, l, half, rem);
    }
    M.A = abase;
    if (active) {  // registers -> state block; finish_lane advances the modulator's accumulator
        // the magic bits above the phase cancel in the difference or leave at the top of the shift
        st64(M, w0.b, ld64(M, w0.b) + ((p - p_start) << 20));
        if (TAIL) {
            const int st = w1.c;
            stf(M, st + 2, F.x2); stf(M, st + 3, F.x1);
            stf(M, st + 4, F.y2); stf(M, st + 5, F.y1);
        }
    }
}

// The summary pass of a time-axis split of a fused FM voice (abi.cpp render_split_fm): only the carrier's phase
// sum matters — the modulator's tile, the affine map and the phase steps; no sines, no filter, no rows (a third of
// the full tile's instructions).  After `snap_tile` tiles the accumulator so far is recorded in *snap: the phase
// at the point where the NEXT segment's filter warm-up starts — and, when the voice has a filter, in the first two
// words of its history in the state block as well (nothing reads a filter's history after a summary pass), so that
// the snapshot travels with the state blocks the ranks of a time-sharded render exchange.
template <bool SLOW>
__device__ __forceinline__ void run_fm_sums(const tb_insn* code, LaneMem& M, const SineK& sk, bool active, u64 n_tiles,
                                            u64 snap_tile, unsigned long long* snap) {
    if (!active) return;
    const tb_insn w0 = code[0], w1 = code[1];
    const double2* rot = reinterpret_cast<const double2*>(M.Q + (size_t)((w0.op >> 8) & 0xffu) * LT);
    double S = ldd(M, w0.a), Cq = ldd(M, w0.a + 2);
    const float m = ldf(M, w1.a), c = ldf(M, w1.b);
    const u64 mm = pk2(m, m), cc = pk2(c, c);
    const u64 acc0 = ld64(M, w0.b);
    const u64 p_start = (acc0 + ld64(M, w0.c)) >> 20;
    u64 p = p_start;
    const FmRot rr = fm_rot_load(rot);
    float car[LS];
    u64 snapped = acc0;
    for (u64 t = 0; t < n_tiles; t++) {
        if (t == snap_tile) snapped = acc0 + ((p - p_start) << 20);
        fm_carrier_tile<SLOW, false, true>(car, S, Cq, rot, rr, mm, cc, p, sk);
    }
    if (n_tiles == snap_tile) snapped = acc0 + ((p - p_start) << 20);
    if (snap) *snap = snapped;
    if (w1.c >= 0) st64(M, w1.c + 2, snapped);
    st64(M, w0.b, acc0 + ((p - p_start) << 20));
}

// One unit of work: the 64 voices of `group`, samples [s0, s0 + ns) of the launch (multiples of TB_LS).
//   SUMS: the summary kernel of lanes_fm_split.cu (run_fm_sums).
//   FM_ONLY: the kernels of lanes_fm.cu, launched by the host only for a program that is one fused FM
//   voice (run_fm_voice; no interpreter in the kernel).
template <bool MIX, bool FM_ONLY, bool SUMS = false>
__device__ __forceinline__ void lanes_body(const tb_launch& P, uint32_t group, u64 s0, u64 ns, bool accumulate) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int t = threadIdx.x, l = t & 31, warp = t >> 5;
    tb_insn* code = reinterpret_cast<tb_insn*>(smem_raw);
    for (uint32_t k = t; k < P.n_lane_code; k += LT) code[k] = P.lane_code[k];
    const size_t q_bytes = (size_t)(P.lane_q_units + 4 * P.lane_slots) * LT * 16;
    unsigned char* base = smem_raw + (size_t)P.n_lane_code * sizeof(tb_insn);
    LaneMem M;
    M.Q = reinterpret_cast<float4*>(base) + t;
    M.slot0 = P.lane_q_units;
    float4* atile = reinterpret_cast<float4*>(base + q_bytes);
    M.A = atile + t;
    M.W = reinterpret_cast<uint32_t*>(base + q_bytes + (size_t)8 * AS * 16) + t;
    __syncthreads();

    const uint32_t voice = group * LT + t;
    const uint32_t v0 = group * LT + warp * 32;
    bool active = voice < P.n_voices;
    // Single fused FM voice with a FAST carrier on the special-function unit (run_fm_voice).
    const bool fm_program = P.n_lane_code == 3 && (code[0].op & 0xffu) == LN_FM && ((code[0].op >> 16) & 0xffu) == 0 &&
                            (code[0].op >> 24) == TB_SINE_FAST && P.fast_mode == 2;
    FmPrime prime = {0, 0, false};
    if (fm_program) {
        for (uint32_t k = 0; k < P.n_lane_aux; k++) {
            const tb_lane_aux a = P.lane_aux[k];
            if (a.kind == LA_ROT && (int)a.w_off + 2 == code[0].a) {
                prime.acc_w = a.b;
                prime.ph_cval = a.c;
            }
        }
    }
    SineK sk;
    sk.kscale = 17592186044416.0 / (TB_TAU * (double)P.sample_rate);
    sk.pscale = 17592186044416.0 / TB_TAU;
    sk.inv_turn = 1.0 / (TB_TAU * (double)P.sample_rate);
    sk.flimit = (float)(TB_FM_TURNS * TB_TAU * (double)P.sample_rate);  // tighter than render.cu's 100: see pd_make
    sk.plimit = 600.0f;
    sk.one23 = 0x3f800000u | (P.sample_rate >> 31);  // (sample rates stay below 2^31: tb_program_create)

    // the voice's state block; with the time-axis split (tb_launch::vsplit*) that of its segment
    const uint32_t rvoice = !TB_LANES_VSPLIT ? voice : voice / (P.vsplit ? P.vsplit : 1u);  // parameters, noise streams
    const uint32_t vseg_i = !TB_LANES_VSPLIT ? 0u : P.vseg_lo + (voice - rvoice * P.vsplit);
    const size_t vidx = TB_LANES_VSPLIT ? (size_t)rvoice * P.vsplit_total + vseg_i : (size_t)voice;
    uint32_t* gstate = P.state + vidx * P.state_words;
    if (active) {
        // ld.cg: with the work queue the block was last written by another CTA, possibly on another SM
        for (uint32_t k = 0; k < P.state_words; k++) stw(M, (int)(P.n_cval + k), __ldcg(gstate + k));
        setup_lane(P, M, P.params ? P.params + (size_t)rvoice * P.n_params : nullptr);
        // Every filter must hold its full history (generator.rs:234-252): the host renders the
        // first tile of a stream with the general interpreter before it comes here.
        bool ready = true;
        for (uint32_t fi = 0; fi < P.n_filt; fi++) {
            const int so = (int)(P.n_cval + P.filt[fi].state_off);
            ready = ready && ldw(M, so) != 0u && ldw(M, so + 1) == P.filt[fi].K - 1u;
        }
        if (!ready) {
            // The fused FM voice starts its own stream (run_fm_voice): its one filter has never run.
            if (fm_program && P.n_filt == 1 && code[1].c >= 0 && ldw(M, code[1].c) == 0u) prime.fresh = true;
            else {
                if (P.fault) atomicAdd(P.fault, 1u);
                active = false;
            }
        }
    }
    if (!active) {
        UNROLL for (int q = 0; q < 8; q++) M.A[q * AS] = make_float4(0.f, 0.f, 0.f, 0.f);
    }

    RowStore R;
    R.out = P.out ? P.out + s0 : nullptr;
    R.stride = P.out_stride;
    R.tbase = atile + warp * 32;
    R.v0 = v0;
    R.n_voices = P.n_voices;
    R.off = 0;
    // Rows that are not 16-byte aligned (odd strides) take four scalar stores per lane instead.
    R.vec_ok = (reinterpret_cast<uintptr_t>(P.out) & 15) == 0 && (P.out_stride & 3) == 0;
    R.fast = R.vec_ok && P.out != nullptr && v0 + 32u <= P.n_voices && (!TB_LANES_VSPLIT || (P.vseg & 3) == 0);
    R.mix = MIX ? P.mix_partial + (size_t)(v0 >> 5) * P.mix_stride + s0 : nullptr;
    R.rowoff = nullptr;
    R.step4 = 4 * (size_t)P.out_stride;
    R.dfast = R.fast ? R.out + (size_t)(v0 + (uint32_t)(l >> 3)) * R.stride + (size_t)(l & 7) * 4 : nullptr;
#if TB_LANES_VSPLIT
    __shared__ unsigned long long rowoff_s[LT];  // the CTA's rows: offset of each one's first sample of this launch
    rowoff_s[t] = (unsigned long long)rvoice * P.out_stride + (unsigned long long)vseg_i * P.vseg;
    __syncthreads();
    R.rowoff = rowoff_s + warp * 32;
#endif
    const bool warp_live = __any_sync(FULL, active);
    const uint32_t code_s = (uint32_t)__cvta_generic_to_shared(code);
    const u64 n_tiles = ns / (u64)LS;

    // The magic-number conversion of the carrier frequency is exact while |m| + |c| stays below its range
    // limit; a warp holding a voice beyond it converts the slow way.
    bool fm_slow = false;
    if (fm_program) {
        bool ok = true;
        if (active) ok = fabsf(ldf(M, code[1].a)) + fabsf(ldf(M, code[1].b)) < sk.flimit;
        fm_slow = !__all_sync(FULL, ok);
    }
    const int rem = (int)(ns % (u64)LS);  // only the fused FM voice is launched with samples beyond whole tiles
    if (warp_live || (MIX && v0 < P.n_voices)) {  // a warp without a live voice still owes its (zero) partial sums
        if (fm_program && SUMS) {
            unsigned long long* snap = P.vsnap ? P.vsnap + vidx : nullptr;
            if (!fm_slow) run_fm_sums<false>(code, M, sk, active, n_tiles, P.vsnap_at / (u64)LS, snap);
            else run_fm_sums<true>(code, M, sk, active, n_tiles, P.vsnap_at / (u64)LS, snap);
        } else if (fm_program) {
            const bool tail = code[1].c >= 0;
            if (!fm_slow) {
                if (tail) run_fm_voice<true, MIX, false>(code, M, sk, R, active, l, n_tiles, rem, prime);
                else run_fm_voice<false, MIX, false>(code, M, sk, R, active, l, n_tiles, rem, prime);
            } else {
                if (tail) run_fm_voice<true, MIX, true>(code, M, sk, R, active, l, n_tiles, rem, prime);
                else run_fm_voice<false, MIX, true>(code, M, sk, R, active, l, n_tiles, rem, prime);
            }
        } else if constexpr (!FM_ONLY) {
            float4* const abase = M.A;
            for (u64 t = 0; t < n_tiles; t++) {
                if (active) {
                    M.A = abase + (t & 1) * 4 * AS;
                    if (P.fast_mode == 2) run_lane_tile<2>(P, code_s, M, rvoice, sk);
                    else run_lane_tile<1>(P, code_s, M, rvoice, sk);
                }
                tile_done<MIX>(R, l, t, n_tiles);
            }
            M.A = abase;
        }
    }
    if (active) {
        finish_lane(P, M, ns);
        u64 mine = ns;  // samples of this unit that belong to the voice
        const u64 before = (accumulate && P.out_len) ? __ldcg(P.out_len + vidx) : 0ull;
        if (P.lane_fin_goe >= 0) {
            // Root Fin (generator.rs:133-168): the inner tree was rendered for the whole unit (the tail of a
            // finished voice's row is undefined by contract, generator.rs:76-95); what counts is how far the
            // analytic length reaches (greater_or_equals_at, :787-862, the arithmetic of goe_eval in render.cu).
            const tb_goe g = P.goe[P.lane_fin_goe];
            float value = 0.0f;
            for (uint32_t k = 0; k < g.n_steps; k++) {
                const int sign = P.goe_steps[g.step_off + 2 * k];
                const float c = ldf(M, P.goe_steps[g.step_off + 2 * k + 1]);
                value = sign > 0 ? __fadd_rn(value, c) : __fsub_rn(value, c);
            }
            u64 rem = ~0ull;
            if (g.term == GOE_TIME) {
                const int wt = (int)P.n_cval + g.term_arg;
                const u64 pos = ld64(M, wt);
                const u64 target = f32_as_usize(ceilf(__fmul_rn(value, (float)P.sample_rate)));
                rem = target > pos ? target - pos : 0ull;
                st64(M, wt, pos + ns);  // Fin advances both children to the end of the block (:141-167)
            } else if (ldf(M, g.term_arg) >= value) {
                rem = 0ull;
            }
            if (before < P.call_pos + s0) rem = 0ull;  // returned short earlier in this call
            mine = rem < ns ? rem : ns;
        }
        for (uint32_t k = 0; k < P.state_words; k++) gstate[k] = ldw(M, (int)(P.n_cval + k));
        if (P.out_len) P.out_len[vidx] = before + mine;
    }
}

}  // namespace
