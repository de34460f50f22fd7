// common.cuh — device helpers shared by the kernels of render.cu (one warp per voice) and
// lanes.cuh (one lane per voice): fixed-point sine phase, sine cores, noise streams, state words.
// Included inside each translation unit's anonymous namespace.
#pragma once

constexpr int C = TB_C;
constexpr int TILE = TB_TILE;
constexpr unsigned FULL = 0xffffffffu;
typedef unsigned long long u64;
typedef long long i64;

static_assert(C == 8, "slot layout and unrolled loops assume 8 samples per lane");

#define TB_TAU 6.283185307179586476925286766559
#define UNROLL _Pragma("unroll")

// ------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// Rust `f as usize` (saturating, NaN -> 0), used on ceil(value * sr)  (generator.rs:813).
__device__ __forceinline__ u64 f32_as_usize(float f) {
    if (!(f > 0.f)) return 0ull;
    if (f >= 18446744073709551616.0f) return ~0ull;
    return (u64)f;
}

// Shared-memory slot: float4 #q (q = 0,1) of lane l sits at float4 index q*32 + l, so the two
// 128-bit accesses of a warp are bank-conflict free.  Only the owning lane touches its samples,
// except the serial feedback fallback which goes through slot_index().
__device__ __forceinline__ void slot_store(float* slots, int s, const float (&v)[C]) {
    float4* p = reinterpret_cast<float4*>(slots + (size_t)s * TILE);
    const int l = lane_id();
    p[l] = make_float4(v[0], v[1], v[2], v[3]);
    p[32 + l] = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void slot_load(const float* slots, int s, float (&v)[C]) {
    const float4* p = reinterpret_cast<const float4*>(slots + (size_t)s * TILE);
    const int l = lane_id();
    float4 a = p[l], b = p[32 + l];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ int slot_index(int i) {
    const int l = i >> 3, j = i & 7;
    return (((j >> 2) * 32 + l) << 2) + (j & 3);
}

// Inclusive warp prefix sum of 64-bit integers (exact and associative: the phase is
// independent of tile size and launch size).
__device__ __forceinline__ u64 warp_incl_sum(u64 x) {
    const int l = lane_id();
    UNROLL for (int d = 1; d < 32; d <<= 1) {
        u64 t = __shfl_up_sync(FULL, x, d);
        if (l >= d) x += t;
    }
    return x;
}

// ------------------------------------------------------------------------------------------
// Sine: 64-bit fixed-point phase in units of 2^-64 turns (generator.rs:206-219).
// ------------------------------------------------------------------------------------------
struct SineK {
    double kscale;    // 2^44 / (TAU * sample_rate): rad/s -> 2^-44 turns per sample
    double pscale;    // 2^44 / TAU:                 rad   -> 2^-44 turns
    double inv_turn;  // 1 / (TAU * sample_rate)
    float flimit;     // |f| below which the magic-number conversion is exact
    float plimit;
    uint32_t one23;   // lanes.cuh pd_m23: the bits of 1.0f in a register the compiler cannot see through
};

// rint(x) for |x| < 2^51 through the 1.5*2^52 trick; result shifted to 2^-64-turn units, so whole
// turns wrap away exactly like rem_euclid(TAU) (generator.rs:218).
__device__ __forceinline__ u64 magic_to_fx(double scaled_plus_magic) {
    i64 q = __double_as_longlong(scaled_plus_magic) - 0x4338000000000000LL;
    return (u64)q << 20;
}
// Full-precision conversion of a turn count of any magnitude (setup time and out-of-range inputs).
__device__ __noinline__ u64 turns_to_fx_slow(double turns) {
    if (!(fabs(turns) < 1e300)) return 0ull;  // inf / nan: the reference's phase is NaN too
    double fr = turns - floor(turns);         // [0,1]
    u64 v = __double2ull_rn(fr * 9223372036854775808.0);  // 2^63
    return v << 1;
}
__device__ __forceinline__ u64 freq_to_inc(float f, const SineK& k) {
    if (fabsf(f) < k.flimit) return magic_to_fx(fma((double)f, k.kscale, 6755399441055744.0));
    return turns_to_fx_slow((double)f * k.inv_turn);
}
__device__ __forceinline__ u64 phase_to_fx(float p, const SineK& k) {
    if (fabsf(p) < k.plimit) return magic_to_fx(fma((double)p, k.pscale, 6755399441055744.0));
    return turns_to_fx_slow((double)p * (1.0 / TB_TAU));
}
// Vector forms: the magic-number conversion for all C samples, then ONE warp vote decides whether
// any input was outside its exact range (|f| >= 100*TAU*sr: never for audio) and redoes those.
__device__ __forceinline__ void freq_to_inc_vec(u64 (&inc)[C], const float (&f)[C], const SineK& k) {
    float big = 0.0f;
    UNROLL for (int j = 0; j < C; j++) {
        inc[j] = magic_to_fx(fma((double)f[j], k.kscale, 6755399441055744.0));
        big = fmaxf(big, fabsf(f[j]));
    }
    if (__any_sync(FULL, !(big < k.flimit))) {
        UNROLL for (int j = 0; j < C; j++)
            if (!(fabsf(f[j]) < k.flimit)) inc[j] = turns_to_fx_slow((double)f[j] * k.inv_turn);
    }
}
__device__ __forceinline__ void phase_to_fx_vec(u64 (&ph)[C], const float (&p)[C], const SineK& k) {
    float big = 0.0f;
    UNROLL for (int j = 0; j < C; j++) {
        ph[j] = magic_to_fx(fma((double)p[j], k.pscale, 6755399441055744.0));
        big = fmaxf(big, fabsf(p[j]));
    }
    if (__any_sync(FULL, !(big < k.plimit))) {
        UNROLL for (int j = 0; j < C; j++)
            if (!(fabsf(p[j]) < k.plimit)) ph[j] = turns_to_fx_slow((double)p[j] * (1.0 / TB_TAU));
    }
}

// sin(2*pi * ph / 2^64).  Fold to [-1/4, 1/4] turn with integer ops (exact, branch-free), then
// sin(pi/2 x) = x P(x^2) on x in [-1, 1]; coefficients from tools/fit_sine.py.
__device__ __forceinline__ u64 fold_quarter(u64 ph) {
    const u64 m = (u64)((i64)(ph ^ (ph << 1)) >> 63);  // all ones in the 2nd and 3rd quarter turn
    return ((ph ^ m) - m) ^ (m & 0x8000000000000000ull);  // there: 2^63 - ph
}
// EXACT: f64, 7 coefficients, |err| < 8e-14, rounded once to f32: the same f32 as the
// reference's `(acc + ph).sin() as f32` except where the f64 values straddle an f32 rounding
// boundary (measured: 0.02 % of the samples, by one ulp, sign-symmetric).  The coefficients carry
// the 2^-62 scaling of the integer phase (c_k * 2^(-62 - 124 k)), so no extra multiply is needed.
// The fold is done on the converted double: |x| > 2^62  ->  x = copysign(2^63, x) - x (exact).
__constant__ double c_sin_exact[7] = {0x1.921fb54442bb4p-62,  -0x1.4abbce624ad99p-187, 0x1.466bc66ed3d1cp-314,
                                      -0x1.32d2c9b2d1df5p-442, 0x1.50770f6a5a66bp-571,  -0x1.e29b82ab98ea9p-701,
                                      0x1.d53abdeb199c1p-831};
__device__ __forceinline__ float sin_turns_exact(u64 ph) {
    double x = (double)(i64)ph;  // signed turns * 2^64, in [-2^63, 2^63]
    const int hi = __double2hiint(x);
    const double half = __hiloint2double((hi & 0x80000000) | 0x43e00000, 0);  // copysign(2^63, x)
    const double folded = half - x;
    x = ((hi & 0x7fffffff) > 0x43d00000) ? folded : x;  // |x| > 2^62 (the = case folds to itself)
    const double z = x * x;
    double p = c_sin_exact[6];
    p = fma(p, z, c_sin_exact[5]);
    p = fma(p, z, c_sin_exact[4]);
    p = fma(p, z, c_sin_exact[3]);
    p = fma(p, z, c_sin_exact[2]);
    p = fma(p, z, c_sin_exact[1]);
    p = fma(p, z, c_sin_exact[0]);
    return (float)(x * p);
}
// FAST: f32, 5 coefficients, |err| < 2e-7 — for sines whose output reaches only the sample
// stream (never a frequency, phase, trigger, length or filter coefficient).  Only the top 32
// phase bits matter here, so the fold is 32-bit.
__device__ __forceinline__ float sin_turns_fast(u64 ph) {
    const int h = (int)(ph >> 32);
    const int m = (h ^ (h << 1)) >> 31;
    const int f = ((h ^ m) - m) ^ (m & (int)0x80000000);
    const float x = (float)f * 9.31322574615478515625e-10f;  // 2^-30
    const float z = x * x;
    float p = 0.00015167170204222202f;
    p = fmaf(p, z, -0.004674143623560667f);
    p = fmaf(p, z, 0.07968991994857788f);
    p = fmaf(p, z, -0.6459637880325317f);
    p = fmaf(p, z, 1.5707963705062866f);
    return x * p;
}
// FAST through the special-function unit (tb_launch::fast_mode == 2): the top 32 phase bits as
// radians in [-pi, pi), sin.approx = range-reduction multiply + MUFU.SIN, |err| <= 2^-21.4.
__device__ __forceinline__ float sin_turns_mufu(u64 ph) {
    return __sinf((float)(int)(ph >> 32) * 1.4629180792671596e-09f);  // 2 pi / 2^32
}
// `fast`: 0 = EXACT, 1 = f32 polynomial, 2 = MUFU.
__device__ __forceinline__ float sin_turns(u64 ph, int fast) {
    return fast == 0 ? sin_turns_exact(ph) : (fast == 1 ? sin_turns_fast(ph) : sin_turns_mufu(ph));
}
__device__ __forceinline__ void sin_turns_vec(float (&out)[C], const u64 (&ph)[C], int fast) {
    if (fast == 2) { UNROLL for (int j = 0; j < C; j++) out[j] = sin_turns_mufu(ph[j]); }
    else if (fast == 1) { UNROLL for (int j = 0; j < C; j++) out[j] = sin_turns_fast(ph[j]); }
    else      { UNROLL for (int j = 0; j < C; j++) out[j] = sin_turns_exact(ph[j]); }
}

// ------------------------------------------------------------------------------------------
// Noise (generator.rs:113-118): `fastrand::f32() * 2 - 1`.  fastrand 2.3.0's generator is wyrand:
// state += C0; t = state * (state ^ C1) as u128; out = lo(t) ^ hi(t); f32 = from_bits(0x3F800000 |
// (u32 >> 9)) - 1.  The state advances by a constant, so sample k of a stream is a pure function of
// (seed, k): every Noise node of every voice owns the stream
//     state_k = seed + NODE_K (node + 1) + VOICE_K voice + C0 (k + 1)
// (the reference draws from one UNSEEDED thread-local instance: no sequence of it is reproducible,
// so parity for Noise is pinned against the oracle's restatement of the same streams only).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ u64 noise_stream(const tb_launch& P, uint32_t voice, int node) {
    return P.noise_seed + 0x9e3779b97f4a7c15ull * (u64)(node + 1) + 0xd6e8feb86659fd93ull * (P.voice_base + (u64)voice);
}
__device__ __forceinline__ float noise_at(u64 stream, u64 k) {
    const u64 s = stream + 0x2d358dccaa6c78a5ull * (k + 1ull);
    const u64 m = s ^ 0x8bb84b93962eacc9ull;
    const u64 r = (s * m) ^ __umul64hi(s, m);
    const float f = __uint_as_float(0x3F800000u | ((uint32_t)r >> 9)) - 1.0f;
    return __fsub_rn(__fmul_rn(f, 2.0f), 1.0f);
}

__device__ __forceinline__ float apply1(uint32_t op, float a, float b) {
    switch (op) {
        case TB_ADD:
        case TB_MERGE: return __fadd_rn(a, b);
        case TB_SUBTRACT: return __fsub_rn(a, b);
        case TB_MULTIPLY: return __fmul_rn(a, b);
        case TB_DIVIDE: return b == 0.0f ? 0.0f : __fdiv_rn(a, b);
        default: return powf(a, b);
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

// FAST class from the top 32 phase bits h (2^-32 turns, two's complement = [-1/2, 1/2) turn).
//   MODE 1: the f32 polynomial of the general path (|err| < 2e-7);
//   MODE 2: the special-function unit — radians in [-pi, pi), sin.approx = range-reduction
//           multiply + MUFU.SIN, |err| <= 2^-21.4 (CUDA math API, __sinf on [-pi, pi]).
template <int MODE>
__device__ __forceinline__ float sin_hi(int h) {
    if (MODE == 2) return __sinf((float)h * 1.4629180792671596e-09f);  // 2 pi / 2^32
    const int m = (h ^ (h << 1)) >> 31;
    const int f = ((h ^ m) - m) ^ (m & (int)0x80000000);
    const float x = (float)f * 9.31322574615478515625e-10f;  // 2^-30
    const float z = x * x;
    float p = 0.00015167170204222202f;
    p = fmaf(p, z, -0.004674143623560667f);
    p = fmaf(p, z, 0.07968991994857788f);
    p = fmaf(p, z, -0.6459637880325317f);
    p = fmaf(p, z, 1.5707963705062866f);
    return x * p;
}

// Bits of (f * scale + 1.5 * 2^52): the low mantissa bits hold rint(f * scale) in 2^-44 turns.
// Shifted left by 20 they are 2^-64 turns and the exponent / magic bits fall off the top, so
// sums of raw words can be shifted once at the end:  (sum raw) << 20 == sum (q << 20) mod 2^64.
__device__ __forceinline__ u64 magic_raw(float f, double scale) {
    return (u64)__double_as_longlong(fma((double)f, scale, 6755399441055744.0));
}
__device__ __forceinline__ int raw_hi(u64 raw) {  // top 32 bits of raw << 20
    return (int)__funnelshift_l((uint32_t)raw, (uint32_t)(raw >> 32), 20);
}

