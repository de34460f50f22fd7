// lanes_fm_ws.cuh — the fused FM voice (LN_FM + biquad, program.h) with TWO WARPS PER 32 VOICES: the device code of
// lanes_fm_ws.cu (real voices) and lanes_fm_ws_split.cu (TB_LANES_VSPLIT: the segments of a time-axis split).
//
// lanes_fm.cu gives a voice one thread for everything; 65,536 voices are then 13.8 warps an SM, all the
// parallelism the batch has, and the loop is bound by latency (issue slots 54 % used, the conversion /
// special-function unit 54 %: profiles/r2_*).  Here the voice's work is cut where its data flow is
// narrowest and the halves run as a producer and a consumer warp, lane l of both owning voice l:
//   * the PHASE warp: modulator tile (f64, three-term recurrence from the tile's centre), the affine map, the running
//     carrier phase (one DFMA per sample on the 2^-44-turn grid, lanes.cuh pd_make) — it hands over 16 words a tile
//     (fm_carrier_tile<.., RAW>): the arguments of the carrier's sines, made from the top 23 phase bits, and for the
//     tile's first four samples the sines themselves;
//   * the TONE warp: the other sines on the special-function unit, the biquad in the reference's operation order
//     (generator.rs:496-507, lanes.cuh biquad_tile), the row transpose and the stores.
// Twice the warps with the same instructions per sample (+ 8 shared-memory instructions a tile for the
// hand-over) and about half the registers each: 15 CTAs of 64 threads an SM at <= 64 registers.
// The hand-over is a two-buffer ring in shared memory guarded by named barriers (bar.sync / bar.arrive
// with 64 participants: PTX's producer / consumer idiom); the ring lies over the modulator's rotation
// table, which the phase warp holds in registers by then.  Columns of shared memory are 32 wide here
// (TB_LANE_THREADS below), so lanes.cuh's setup / finish / row-store code is used as it is.
// State blocks, stream position, rows and partial mix rows are those of lanes_fm.cu: results are
// bit-identical (tests/test_gpu_lanes.py::test_fm_ws_matches_single_thread_kernel).
#pragma once
#define TB_LANE_THREADS 32
// Measured on 65,536 voices x 2 s (this kernel 6.00 ms as first written): the conversion / special-function unit is the
// busy one (56 % of its instruction rate, in bursts), so the carrier frequency becomes a double by integer
// instructions here (lanes.cuh f32_to_f64_alu: four instructions instead of one F2F.F64.F32; 5.90 ms) — a trade the
// one-thread form loses (32.3 against 31.1 ms there) — and with that the phase warp also makes the sines' arguments
// (the FFMA2 of sin_m23x2: 5.85 ms; without the integer widening that move costs 0.05 ms).
#ifndef TB_WIDEN_ALU
#define TB_WIDEN_ALU 1
#endif
#ifndef TB_WS_ARG
#define TB_WS_ARG 1
#endif
// ... and the sines themselves of a tile's first TB_WS_SINES chunks of four samples: with one, the tone warp's filter
// starts on a tile without waiting for a sine (5.29 -> 5.20 ms; two: 5.30; while the phase warp was the slower one —
// before the integer widening, the recurrence and the loops by voice class — one cost 0.10 ms).
#ifndef TB_WS_SINES
#define TB_WS_SINES 1
#endif
#include "lanes.cuh"

namespace {

constexpr uint32_t WS_THREADS = 64;
// Named barriers.  Four, not five: barrier 0 (__syncthreads before the loops) doubles as "buffer 1 is full", and the
// rendezvous after the loops is one more round of "buffer 0 is empty" (every arrival of the tone warp on it has been
// consumed by then: the tone warp announces a buffer as empty only when the phase warp has a tile left for it).
// Measured: a B200 SM holds 12 of these CTAs with 5 barriers each, whatever their registers and shared memory.
constexpr uint32_t BAR_FULL0 = 1, BAR_FULL1 = 0, BAR_EMPTY0 = 2, BAR_EMPTY1 = 3;
template <uint32_t B> struct Bar {
    static constexpr uint32_t full = B ? BAR_FULL1 : BAR_FULL0;
    static constexpr uint32_t empty = B ? BAR_EMPTY1 : BAR_EMPTY0;
};

// (the __syncwarp()s in front of the barriers: a CTA with fewer than 32 live voices has diverged around the tile's arithmetic)
__device__ __forceinline__ void ring_put(uint4* ring, const float (&v)[LS]) {
    UNROLL for (int q = 0; q < 4; q++)
        ring[q * LT] = make_uint4(__float_as_uint(v[4 * q]), __float_as_uint(v[4 * q + 1]), __float_as_uint(v[4 * q + 2]),
                                  __float_as_uint(v[4 * q + 3]));
}

// Barrier numbers are immediates (a register operand makes ptxas reserve all 16 barriers for the CTA, which costs
// resident CTAs), so the loops below take tiles in pairs: buffer 0, then buffer 1.
//   CONV: the warp is converged by construction (the loops compiled for "every lane live" have no branch around the
//   tile's arithmetic), so the __syncwarp() in front of the barrier is left out.
template <uint32_t ID, bool CONV = false>
__device__ __forceinline__ void nb_sync() {
    if (!CONV) __syncwarp();
    asm volatile("bar.sync %0, 64;" ::"n"(ID) : "memory");
}
template <uint32_t ID, bool CONV = false>
__device__ __forceinline__ void nb_arrive() {
    if (!CONV) __syncwarp();
    asm volatile("bar.arrive %0, 64;" ::"n"(ID) : "memory");
}

struct PhaseRegs {
    double S, Cq;
    u64 mm, cc, p;
    FmRot rr;
};
// One tile of the phase warp into buffer B.  CAP: the tile of which only the first `rem` samples count.
//   ALL: every lane has a live voice (the flag would otherwise sit in a register this warp does not have: measured,
//   18 % of the kernel's stall samples were the two warps waiting for it to come back from local memory).
//   FMODE: what the warp's voices allow (lanes.cuh fm_carrier_tile): 1 = no negative carrier frequency, 2 = no modulation.
template <bool SLOW, uint32_t B, bool CAP, bool ALL = false, int FMODE = 0>
__device__ __forceinline__ void phase_step(PhaseRegs& G, const double2* rot, const SineK& sk, uint4* ring, bool active, int rem) {
    // (Making the tile in registers first and asking for the buffer only then — a tile further ahead of the tone warp —
    // was tried: sixteen values alive across the barrier spill, 5.59 -> 5.84 ms.)
    nb_sync<Bar<B>::empty, ALL>();
    if (ALL || active) {
        float raw[LS];
        if (CAP) {
            u64 p_rem = G.p;
            fm_carrier_tile<SLOW, true, false, true>(raw, G.S, G.Cq, rot, G.rr, G.mm, G.cc, G.p, sk, rem, &p_rem);
            G.p = p_rem;
        } else {
            fm_carrier_tile<SLOW, false, false, true, FMODE>(raw, G.S, G.Cq, rot, G.rr, G.mm, G.cc, G.p, sk);
        }
        ring_put(ring + B * 4 * LT, raw);
    }
    nb_arrive<Bar<B>::full, ALL>();
}

// The phase warp: n_tiles tiles into the ring (tile i into buffer i & 1), then — when `rem` > 0 — the tile of which
// only the first `rem` samples count: the phase is taken where they end.
template <bool SLOW>
__device__ __forceinline__ void ws_phase(const tb_insn* code, LaneMem& M, const SineK& sk, uint4* ring, bool active,
                                         u64 n_tiles, int rem) {
    const tb_insn w0 = code[0], w1 = code[1];
    const double2* rot = reinterpret_cast<const double2*>(M.Q + (size_t)((w0.op >> 8) & 0xffu) * LT);
    PhaseRegs G = {};
    G.Cq = 1.0;
    u64 p_start = 0;
    if (active) {
        G.S = ldd(M, w0.a);
        G.Cq = ldd(M, w0.a + 2);
        const float m = ldf(M, w1.a), c = ldf(M, w1.b);
        G.mm = pk2(m, m);
        G.cc = pk2(c, c);
        p_start = G.p = (ld64(M, w0.b) + ld64(M, w0.c)) >> 20;
        G.rr = fm_rot_load(rot);
    }
    __syncwarp();  // the ring lies over the rotation table: every lane has read its entries
    if (__all_sync(FULL, active)) {
        // What the 32 voices have in common decides how much of the tile is needed (same bits either way): no
        // modulation (m == 0: a plain filtered sine), or a carrier frequency that never turns negative (c >= |m|).
        float m = 0.0f, c = 0.0f, dummy;
        unpk2(G.mm, m, dummy);
        unpk2(G.cc, c, dummy);
        const bool flat = !SLOW && __all_sync(FULL, m == 0.0f);
        const bool pos = !SLOW && __all_sync(FULL, c >= fabsf(m));
        if (flat) {
            for (uint32_t left = (uint32_t)(n_tiles >> 1); left != 0; left--) {
                phase_step<SLOW, 0, false, true, 2>(G, rot, sk, ring, true, 0);
                phase_step<SLOW, 1, false, true, 2>(G, rot, sk, ring, true, 0);
            }
        } else if (pos) {
            for (uint32_t left = (uint32_t)(n_tiles >> 1); left != 0; left--) {
                phase_step<SLOW, 0, false, true, 1>(G, rot, sk, ring, true, 0);
                phase_step<SLOW, 1, false, true, 1>(G, rot, sk, ring, true, 0);
            }
        } else {
            for (uint32_t left = (uint32_t)(n_tiles >> 1); left != 0; left--) {
                phase_step<SLOW, 0, false, true>(G, rot, sk, ring, true, 0);
                phase_step<SLOW, 1, false, true>(G, rot, sk, ring, true, 0);
            }
        }
    } else {
        for (uint32_t left = (uint32_t)(n_tiles >> 1); left != 0; left--) {
            phase_step<SLOW, 0, false>(G, rot, sk, ring, active, 0);
            phase_step<SLOW, 1, false>(G, rot, sk, ring, active, 0);
        }
    }
    if (n_tiles & 1) {
        phase_step<SLOW, 0, false>(G, rot, sk, ring, active, 0);
        if (rem > 0) phase_step<SLOW, 1, true>(G, rot, sk, ring, active, rem);
    } else if (rem > 0) {
        phase_step<SLOW, 0, true>(G, rot, sk, ring, active, rem);
    }
    // the magic bits above the phase cancel in the difference or leave at the top of the shift
    if (active) st64(M, w0.b, ld64(M, w0.b) + ((G.p - p_start) << 20));
}

// The tone warp's filter state: lanes.cuh BiquadRegs plus the feed-forward products that reach into the next tile.
struct ToneRegs {
    float b0, b1, b2, a1, a2;
    float x1, x2;      // x[-1], x[-2]
    float y1, y2;      // y[-1], y[-2]
    float p1, p2, q2;  // a1 y[-1], a2 y[-2], a2 y[-1]
    float b1p;         // b1 x[-1]
    u64 b2p;           // (b2 x[-2], b2 x[-1]): the feed-forward products that reach into the next tile
};
// The four sines of chunk Q from what the phase warp handed over: the sines themselves (Q < TB_WS_SINES), their arguments
// (TB_WS_ARG), or the floats 1.m.
__device__ __forceinline__ void tone_sines(const uint4 mv, float (&x)[4], int Q = 3) {
    if (Q < TB_WS_SINES) {  // (Q is a literal after unrolling)
        x[0] = __uint_as_float(mv.x); x[1] = __uint_as_float(mv.y); x[2] = __uint_as_float(mv.z); x[3] = __uint_as_float(mv.w);
        return;
    }
#if TB_WS_ARG
    x[0] = __sinf(__uint_as_float(mv.x));
    x[1] = __sinf(__uint_as_float(mv.y));
    x[2] = __sinf(__uint_as_float(mv.z));
    x[3] = __sinf(__uint_as_float(mv.w));
#else
    sin_m23x2(mv.x, mv.y, x[0], x[1]);
    sin_m23x2(mv.z, mv.w, x[2], x[3]);
#endif
}
// Four samples through the filter: generator.rs:496-507 for K = 3, J = 2, the operations and roundings of lanes.cuh
// biquad_tile (bit-identical results) in an order that keeps few values alive — this warp lives on 64 registers.
//   b1p: b1 x[i - 1]; b2p: (b2 x[i - 2], b2 x[i - 1]) on entry, the same one chunk on when it returns.
__device__ __forceinline__ void tone_chunk(const float (&xq)[4], float (&yq)[4], ToneRegs& F, float& b1p, u64& b2p) {
    const u64 bb0 = pk2(F.b0, F.b0), bb1 = pk2(F.b1, F.b1), bb2 = pk2(F.b2, F.b2), aa = pk2(F.a1, F.a2);
    UNROLL for (int h = 0; h < 2; h++) {
        const u64 xx = pk2(xq[2 * h], xq[2 * h + 1]);
        float s0, s1, q0, q1;
        unpk2(mul2(xx, bb0), s0, s1);
        unpk2(mul2(xx, bb1), q0, q1);
        s0 = __fadd_rn(s0, b1p);
        s1 = __fadd_rn(s1, q0);
        b1p = q1;
        unpk2(add2(pk2(s0, s1), b2p), s0, s1);
        b2p = mul2(xx, bb2);
        const float y0 = __fsub_rn(__fsub_rn(s0, F.p1), F.p2);
        float n1, n2;
        unpk2(mul2(aa, pk2(y0, y0)), n1, n2);        // a1 y0, a2 y0
        const float y1 = __fsub_rn(__fsub_rn(s1, n1), F.q2);
        F.p2 = n2;
        unpk2(mul2(aa, pk2(y1, y1)), F.p1, F.q2);    // a1 y1, a2 y1
        yq[2 * h] = y0;
        yq[2 * h + 1] = y1;
    }
}
// One tile, chunk by chunk of four samples: what the phase warp handed over comes out of the ring, the sines, the
// filter, the four outputs into the row buffer.  The sines of chunk q + 1 are under way (ring load, range multiply, MUFU)
// before the serial part of chunk q starts: written out, because the compiler will not move a shared-memory load above
// the store of the chunk before it (5.85 -> 5.59 ms on 65,536 voices x 2 s).
//   xn: the sines of the tile's first chunk when PRIMED (the caller, or the tile before, started them), and on return
//   whatever `next` left there: next() runs where chunk 3 would start its successor — the tile loop of ws_tone waits for
//   the other buffer there and starts the first sines of the next tile.
// x_out / y_out (the tile's inputs and outputs in registers) only for the tile a call ends in.
template <bool KEEP, bool PRIMED, typename Next>
__device__ __forceinline__ void tone_tile(const uint4* ring, float4* dst, ToneRegs& F, float (&xn)[4], Next next, float* x_out,
                                          float* y_out) {
    float b1p = F.b1p;  // b1 x[i - 1]
    u64 b2p = F.b2p;    // b2 x[i - 2], b2 x[i - 1]
    if (!PRIMED) tone_sines(ring[0], xn, 0);
    float xq[4], yq[4];
    UNROLL for (int q = 0; q < 4; q++) {
        UNROLL for (int k = 0; k < 4; k++) xq[k] = xn[k];
        if (q + 1 < 4) tone_sines(ring[(q + 1) * LT], xn, q + 1);
        else next();
        tone_chunk(xq, yq, F, b1p, b2p);
        if (KEEP) {
            UNROLL for (int k = 0; k < 4; k++) { x_out[4 * q + k] = xq[k]; y_out[4 * q + k] = yq[k]; }
        }
        dst[q * AS] = make_float4(yq[0], yq[1], yq[2], yq[3]);
    }
    F.x1 = xq[3];
    F.x2 = xq[2];
    F.y1 = yq[3];
    F.y2 = yq[2];
    F.b1p = b1p;
    F.b2p = b2p;
}

// One tile of the tone warp out of buffer B: sines, filter, into half B of the row buffer.
//   refill: the phase warp has a tile to put into this buffer again.
template <uint32_t B, bool ALL = false>
__device__ __forceinline__ void tone_step(ToneRegs& F, float4* abase, const uint4* ring, bool active, bool refill) {
    nb_sync<Bar<B>::full, ALL>();
    float xn[4];
    if (ALL || active) tone_tile<false, false>(ring + B * 4 * LT, abase + B * 4 * AS, F, xn, [] {}, nullptr, nullptr);
    // (handed back after the tile, not after its loads: sixteen words waiting in registers do not fit next to the filter,
    // and the phase warp is the one with time to spare)
    if (refill) nb_arrive<Bar<B>::empty, ALL>();  // (the whole warp, once)
}

// The tone warp: sines, filter, rows.
template <bool MIX>
__device__ __forceinline__ void ws_tone(const tb_insn* code, LaneMem& M, RowStore& R, const uint4* ring, bool active, int l,
                                        u64 n_tiles, int rem) {
    float4* const abase = M.A;
    const tb_insn w1 = code[1];
    const int wc = (int)w1.op, st = w1.c;
    ToneRegs F = {};
    if (active) {
        F.b0 = ldf(M, wc); F.b1 = ldf(M, wc + 1); F.b2 = ldf(M, wc + 2);
        F.a1 = ldf(M, wc + 3); F.a2 = ldf(M, wc + 4);
        F.x2 = ldf(M, st + 2); F.x1 = ldf(M, st + 3);
        F.y2 = ldf(M, st + 4); F.y1 = ldf(M, st + 5);
        F.p1 = __fmul_rn(F.a1, F.y1);
        F.q2 = __fmul_rn(F.a2, F.y1);
        F.p2 = __fmul_rn(F.a2, F.y2);
        F.b1p = __fmul_rn(F.b1, F.x1);
        F.b2p = pk2(__fmul_rn(F.b2, F.x2), __fmul_rn(F.b2, F.x1));
    }
    const u64 slots = n_tiles + (rem > 0 ? 1u : 0u);  // tiles the phase warp makes
    // both buffers start empty
    if (slots > 0) nb_arrive<BAR_EMPTY0>();
    if (slots > 1) nb_arrive<BAR_EMPTY1>();
    const uint32_t pairs = (uint32_t)(n_tiles >> 1);
    // (every pair but the last is followed by at least two more tiles)
    if (!MIX && R.fast && pairs > 1 && __all_sync(FULL, active)) {
        // All 32 rows exist and are 16-byte aligned: lanes.cuh store_pair with nothing but a pointer and the row step
        // alive across the tiles (the general form's bookkeeping does not fit this warp's 64 registers next to the
        // filter), and the 32 x 128 bytes leaving as two batches of four rows per lane.
#if TB_LANES_VSPLIT
        float* d = R.out + (size_t)(l & 7) * 4;                  // rows are segments of the real voices' rows:
        const unsigned long long* ro = R.rowoff + (l >> 3);      // their first samples sit in the CTA's table
#else
        float* d = R.dfast;
        const size_t step4 = R.step4;
#endif
        const float4* src = R.tbase + (l & 7) * AS + (l >> 3);
        // (Tried on this loop, 65,536 voices x 2 s, 5.59 ms as it stands: the buffer handed back right after its last chunk
        // is read instead of after the tile, 5.62 ms; tiles chained — the next tile's buffer waited for and its first sines
        // started while chunk 3 is in the filter — 5.64 ms, 5.79 ms when the buffer is handed back after that wait.)
        auto leave = [&] {
            __syncwarp();
#if TB_LANES_VSPLIT
            UNROLL for (int h = 0; h < 2; h++) {
                float4 v[4];
                unsigned long long o[4];
                UNROLL for (int i = 0; i < 4; i++) { v[i] = src[4 * (4 * h + i)]; o[i] = ro[4 * (4 * h + i)]; }
                UNROLL for (int i = 0; i < 4; i++) st_row(reinterpret_cast<float4*>(d + o[i]), v[i]);
            }
#else
            float* q = d;
            UNROLL for (int h = 0; h < 2; h++) {
                float4 v[4];
                UNROLL for (int i = 0; i < 4; i++) v[i] = src[4 * (4 * h + i)];
                UNROLL for (int i = 0; i < 4; i++) {
                    st_row(reinterpret_cast<float4*>(q), v[i]);
                    q += step4;
                }
            }
#endif
            __syncwarp();
            d += 2 * LS;
        };
        for (uint32_t left = pairs; left > 1; left--) {
            tone_step<0, true>(F, abase, ring, true, true);
            tone_step<1, true>(F, abase, ring, true, true);
            leave();
        }
#if !TB_LANES_VSPLIT
        R.dfast = d;
#endif
        R.off = (size_t)(pairs - 1) * 2 * LS;
    } else {
        for (uint32_t left = pairs; left > 1; left--) {
            tone_step<0>(F, abase, ring, active, true);
            tone_step<1>(F, abase, ring, active, true);
            if (MIX) mix_pair(R, l);
            else store_pair(R, l);
        }
    }
    if (pairs > 0) {
        const u64 after = slots - 2ull * pairs;  // tiles behind the last pair: 0, 1 or 2
        tone_step<0>(F, abase, ring, active, after >= 1);
        tone_step<1>(F, abase, ring, active, after >= 2);
        if (MIX) mix_pair(R, l);
        else store_pair(R, l);
    }
    if (n_tiles & 1) {  // an odd last whole tile leaves alone
        tone_step<0>(F, abase, ring, active, false);
        if (MIX) mix_tile(R, l, 0);
        else store_single(R, l, 0);
    }
    if (rem > 0) {  // the samples that do not fill a tile: the state is taken where they end
        const int half = (int)(n_tiles & 1);
        if (half) nb_sync<BAR_FULL1>();
        else nb_sync<BAR_FULL0>();
        if (active) {
            float car[LS], y[LS];
            float ex[LS + 2], ey[LS + 2];  // history ++ tile
            ex[0] = F.x2; ex[1] = F.x1;
            ey[0] = F.y2; ey[1] = F.y1;
            float xn[4];
            tone_tile<true, false>(ring + half * 4 * LT, abase + half * 4 * AS, F, xn, [] {}, car, y);
            UNROLL for (int j = 0; j < LS; j++) { ex[2 + j] = car[j]; ey[2 + j] = y[j]; }
            UNROLL for (int j = 0; j < LS; j++) {
                if (j == rem) { F.x2 = ex[j]; F.x1 = ex[j + 1]; F.y2 = ey[j]; F.y1 = ey[j + 1]; }
            }
        }
        if (MIX) mix_tile(R, l, half);
        else store_partial(R, l, half, rem);
    }
    M.A = abase;
    if (active) {
        stf(M, st + 2, F.x2); stf(M, st + 3, F.x1);
        stf(M, st + 4, F.y2); stf(M, st + 5, F.y1);
    }
}

// One CTA: voices [32 blockIdx.x, +32), the whole launch.  The host (lanes.cu tb_lanes_launch) sends only
// programs that are one LN_FM with a filter tail (under a root Fin of analytic length or not).
template <bool MIX>
__device__ __forceinline__ void fm_ws_body(const tb_launch& P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t flags_s[3];  // ballot of live voices, "some voice converts the slow way", the phase warp
    const int t = threadIdx.x, l = t & 31, warp = t >> 5;
    tb_insn* code = reinterpret_cast<tb_insn*>(smem_raw);
    for (uint32_t k = t; k < P.n_lane_code; k += WS_THREADS) code[k] = P.lane_code[k];
    const size_t q_bytes = (size_t)(P.lane_q_units + 4 * P.lane_slots) * LT * 16;
    unsigned char* base = smem_raw + (size_t)P.n_lane_code * sizeof(tb_insn);
    LaneMem M;
    M.Q = reinterpret_cast<float4*>(base) + l;
    M.slot0 = P.lane_q_units;
    float4* atile = reinterpret_cast<float4*>(base + q_bytes);
    M.A = atile + l;
    M.W = reinterpret_cast<uint32_t*>(base + q_bytes + (size_t)8 * AS * 16) + l;
    uint4* ring = reinterpret_cast<uint4*>(base) + l;  // over Q units 0 .. 7
    __syncthreads();

    const uint32_t v0 = blockIdx.x * 32u, voice = v0 + (uint32_t)l;
    const u64 ns = P.n_samples;
    SineK sk;
    sk.kscale = P.lane_kscale;
    sk.pscale = 17592186044416.0 / TB_TAU;
    sk.inv_turn = 1.0 / (TB_TAU * (double)P.sample_rate);
    sk.flimit = (float)(TB_FM_TURNS * TB_TAU * (double)P.sample_rate);
    sk.plimit = 600.0f;
    sk.one23 = 0x3f800000u | (P.sample_rate >> 31);  // lanes.cuh pd_m23
    // the voice's state block; with the time-axis split (tb_launch::vsplit*) that of its segment (lanes.cuh lanes_body)
    const uint32_t rvoice = !TB_LANES_VSPLIT ? voice : voice / (P.vsplit ? P.vsplit : 1u);  // parameters
    const uint32_t vseg_i = !TB_LANES_VSPLIT ? 0u : P.vseg_lo + (voice - rvoice * P.vsplit);
    const size_t vidx = TB_LANES_VSPLIT ? (size_t)rvoice * P.vsplit_total + vseg_i : (size_t)voice;
    uint32_t* gstate = P.state + vidx * P.state_words;
#if TB_LANES_VSPLIT
    __shared__ unsigned long long rowoff_s[32];  // the CTA's rows: offset of each one's first sample of this launch
    if (warp == 0) rowoff_s[l] = (unsigned long long)rvoice * P.out_stride + (unsigned long long)vseg_i * P.vseg;
#endif

    if (warp == 0) {  // per-voice setup, as lanes_body does it
        bool active = voice < P.n_voices;
        int acc_w = 0, ph_cval = 0;
        for (uint32_t k = 0; k < P.n_lane_aux; k++) {
            const tb_lane_aux a = P.lane_aux[k];
            if (a.kind == LA_ROT && (int)a.w_off + 2 == code[0].a) {
                acc_w = a.b;
                ph_cval = a.c;
            }
        }
        bool ok = true;
        if (active) {
            for (uint32_t k = 0; k < P.state_words; k++) stw(M, (int)(P.n_cval + k), __ldcg(gstate + k));
            setup_lane(P, M, P.params ? P.params + (size_t)rvoice * P.n_params : nullptr);
            const tb_insn w0 = code[0], w1 = code[1];
            const int st = w1.c;
            const bool ready = ldw(M, st) != 0u && ldw(M, st + 1) == 2u;
            ok = fabsf(ldf(M, w1.a)) + fabsf(ldf(M, w1.b)) < sk.flimit;
            if (!ready && ldw(M, st) == 0u) {
                // The stream starts here: what the filter's first call does in the reference — it reads K - 1 = 2
                // carrier samples ahead and starts from zero outputs (generator.rs:234-252; run_fm_voice's prologue).
                const float m = ldf(M, w1.a), c = ldf(M, w1.b);
                u64 p = (ld64(M, w0.b) + ld64(M, w0.c)) >> 20;
                const u64 p_start = p;
                const u64 inc = ld64(M, w0.a - 2);
                const u64 am = ld64(M, acc_w);
                const u64 phm = turns_to_fx_slow((double)ldf(M, ph_cval) / TB_TAU);
                const float f0 = __fadd_rn(__fmul_rn((float)sin_turns_d8(am + phm), m), c);
                const float f1 = __fadd_rn(__fmul_rn((float)sin_turns_d8(am + phm + inc), m), c);
                const float x2 = sin_m23(p44_m23(p));
                p += !ok ? freq_to_inc(f0, sk) >> 20 : magic_raw(f0, sk.kscale);
                const float x1 = sin_m23(p44_m23(p));
                p += !ok ? freq_to_inc(f1, sk) >> 20 : magic_raw(f1, sk.kscale);
                const u64 am2 = am + 2ull * inc;
                st64(M, acc_w, am2);
                const u64 pc = am2 + phm + inc * (u64)(LS / 2);
                std_(M, w0.a, sin_turns_d8(pc));
                std_(M, w0.a + 2, sin_turns_d8(pc + 0x4000000000000000ull));
                st64(M, w0.b, ld64(M, w0.b) + ((p - p_start) << 20));
                stw(M, st, 1u);      // initialised,
                stw(M, st + 1, 2u);  // K - 1 inputs held
                stf(M, st + 2, x2); stf(M, st + 3, x1);
                stf(M, st + 4, 0.0f); stf(M, st + 5, 0.0f);
            } else if (!ready) {
                if (P.fault) atomicAdd(P.fault, 1u);
                active = false;
            }
        }
        if (!active) {
            UNROLL for (int q = 0; q < 8; q++) M.A[q * AS] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const uint32_t live = __ballot_sync(FULL, active);
        // (a voice that has a slow neighbour must have taken the same branch in the prologue above: the fast form
        // is exact below flimit, so both forms give the same increments there)
        const bool slow = !__all_sync(FULL, ok);
        if (l == 0) {
            flags_s[0] = live;
            flags_s[1] = slow ? 1u : 0u;
            // Warp slots alternate between the SM's four schedulers: pairs of CTAs swap roles so that every
            // scheduler (and its share of the conversion unit) sees both kinds of warp.
            uint32_t slot;
            asm("mov.u32 %0, %%warpid;" : "=r"(slot));
            // (The two warps of a CTA swapping roles every 256 or 2,048 tiles — state through the voice's column of shared
            // memory, the ring's second buffer moved off the rotation table so that the new phase warp can load it —
            // was tried: the swap itself buys 1 %, the layout and the outer loop it needs cost 6 %.)
            // (Measured on the 65,536-voice batch, 5.38 ms with this rule: roles fixed by warp number 6.08 ms — every
            // phase warp on schedulers 0 and 2; other rules that give every scheduler four warps of one kind and three
            // of the other 5.28 - 5.46 ms, depending on which voices meet on a scheduler.  This one balances from four
            // resident CTAs up.)
            flags_s[2] = (slot >> 2) & 1u;
        }
    }
    __syncthreads();
    const bool active = (flags_s[0] >> l) & 1u;
    const bool slow = flags_s[1] != 0u;
    const bool phase_warp = (uint32_t)warp == flags_s[2];
    const u64 n_tiles = ns / (u64)LS;
    const int rem = (int)(ns % (u64)LS);
    const bool any_live = flags_s[0] != 0u;

    if (phase_warp) {
        if (any_live || MIX) {
            if (!slow) ws_phase<false>(code, M, sk, ring, active, n_tiles, rem);
            else ws_phase<true>(code, M, sk, ring, active, n_tiles, rem);
        }
    } else {
        RowStore R;
        R.out = P.out;
        R.stride = P.out_stride;
        R.tbase = atile;
        R.v0 = v0;
        R.n_voices = P.n_voices;
        R.off = 0;
        R.vec_ok = (reinterpret_cast<uintptr_t>(P.out) & 15) == 0 && (P.out_stride & 3) == 0;
        R.fast = R.vec_ok && P.out != nullptr && v0 + 32u <= P.n_voices && (!TB_LANES_VSPLIT || (P.vseg & 3) == 0);
        R.mix = MIX ? P.mix_partial + (size_t)(v0 >> 5) * P.mix_stride : nullptr;
        R.rowoff = nullptr;
#if TB_LANES_VSPLIT
        R.rowoff = rowoff_s;
#endif
        R.step4 = 4 * (size_t)P.out_stride;
        R.dfast = R.fast ? R.out + (size_t)(v0 + (uint32_t)(l >> 3)) * R.stride + (size_t)(l & 7) * 4 : nullptr;
        if (any_live || MIX) ws_tone<MIX>(code, M, R, ring, active, l, n_tiles, rem);
    }
    nb_sync<BAR_EMPTY0>();  // both warps are through
    if (warp == 0 && active) {
        finish_lane(P, M, ns);
        const bool accumulate = P.accumulate != 0;
        const u64 before = (accumulate && P.out_len) ? __ldcg(P.out_len + vidx) : 0ull;
        u64 mine = ns;  // samples of this launch that belong to the voice
        if (P.lane_fin_goe >= 0) {
            // Root Fin (generator.rs:133-168), as in lanes.cuh lanes_body: the inner tree was rendered for the whole launch
            // (the tail of a finished voice's row is undefined by contract, generator.rs:76-95); what counts is how far
            // the analytic length reaches (greater_or_equals_at, :787-862, the arithmetic of goe_eval in render.cu).
            const tb_goe g = P.goe[P.lane_fin_goe];
            float value = 0.0f;
            for (uint32_t k = 0; k < g.n_steps; k++) {
                const int sign = P.goe_steps[g.step_off + 2 * k];
                const float c = ldf(M, P.goe_steps[g.step_off + 2 * k + 1]);
                value = sign > 0 ? __fadd_rn(value, c) : __fsub_rn(value, c);
            }
            u64 left = ~0ull;
            if (g.term == GOE_TIME) {
                const int wt = (int)P.n_cval + g.term_arg;
                const u64 pos = ld64(M, wt);
                const u64 target = f32_as_usize(ceilf(__fmul_rn(value, (float)P.sample_rate)));
                left = target > pos ? target - pos : 0ull;
                st64(M, wt, pos + ns);  // Fin advances both children to the end of the block (:141-167)
            } else if (ldf(M, g.term_arg) >= value) {
                left = 0ull;
            }
            if (before < P.call_pos) left = 0ull;  // returned short earlier in this call
            mine = left < ns ? left : ns;
        }
        for (uint32_t k = 0; k < P.state_words; k++) gstate[k] = ldw(M, (int)(P.n_cval + k));
        if (P.out_len) P.out_len[vidx] = before + mine;
    }
}

}  // namespace

