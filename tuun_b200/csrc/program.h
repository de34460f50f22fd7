// program.h — device byte-code shared by the host lowering (lower.cpp) and the kernels
// (render.cu).
//
// A Waveform tree (reference src/lib/waveform.rs:23-100) is lowered to a flat list of
// fixed-width instructions that ONE WARP executes per voice, with warp-uniform control flow.
// The warp renders a tile of TB_TILE = 32 * TB_C samples per pass: lane l owns the C
// consecutive tile positions [l*C, l*C + C) of every intermediate, the "accumulator" being C
// registers per lane and the named temporaries living in per-warp shared-memory slots.
//
// Three instruction families mirror the three ways the reference walks a tree:
//   G_*  Generator::generate            (generator.rs:86-380)   samples + state
//   L_*  Generator::length              (generator.rs:620-782)  lengths + position state only
//   S_*  generate under a Reset         (generator.rs:281-318)  segmented: every sample knows the
//        tile position at which its run (re)started, so all runs of a tile evaluate at once
#pragma once
#include <stdint.h>

#define TB_C 8               // samples per lane per tile
#define TB_TILE (32 * TB_C)  // samples per tile
#ifndef TB_WARPS_PER_CTA
#define TB_WARPS_PER_CTA 8
#endif
#define TB_CS 16                 // samples per lane per tile of the steady-state interpreter (steady.cuh)
#define TB_TILE_S (32 * TB_CS)
#define TB_LS 16                 // samples per lane per tile of the lane-per-voice kernels (lanes.cuh)
#ifndef TB_LANE_THREADS
#define TB_LANE_THREADS 64       // voices per CTA of the lane-per-voice kernel
#endif
#define TB_MAX_K 9   // feed-forward taps of the register / shuffle filter paths (K-1 <= C), steady and lane kernels
#define TB_MAX_K_GEN 33  // feed-forward taps the general interpreter takes (K-1 <= 32: one history word per lane)
#define TB_MAX_J 4   // feedback taps supported by the scan path
#define TB_MAX_J_GEN 8  // feedback taps the general interpreter takes (serial recurrence past TB_MAX_J)
#define TB_CTL_DEPTH 256  // control-stack words per warp (lower.cpp ctl_need checks a tree against it)

struct tb_insn {
    uint32_t op;
    int32_t a, b, c;
};

enum tb_op : uint32_t {
    OP_END = 0,
    // ---- generate ----
    G_CONST,       // a = cval index
    G_TIME,        // a = state offset
    G_FIXED,       // a = state offset, b = fixed table index (off,len)
    G_NOISE,       // a = state (samples drawn so far), b = node index
    G_BINC,        // a = operator, b = cval index, c = merge
    G_BIN_BEGIN,   // a = slot, b = merge, c = jump target (the matching G_BIN_END)
    G_BIN_END,     // a = slot, b = operator, c = merge
    G_SINE_CC,     // a = state, b = aux index of increment, c = aux index of phase; flags in op>>8
    G_SINE_AC,     // a = state, c = aux index of phase          (frequency in acc)
    G_SINE_CA,     // a = state, b = aux index of increment      (phase in acc)
    G_SINE_BEGIN,  // a = slot, c = jump target (G_SINE_END)
    G_SINE_END,    // a = state, b = slot
    G_ALT_CC,      // a = cval (positive), b = cval (negative)
    G_ALT_BEGIN,   // a = trigger slot, c = jump target (G_ALT_END)
    G_ALT_POS,     // a = slot for the positive branch
    G_ALT_END,     // a = trigger slot, b = positive operand, c = negative operand
    G_FILT_PRE,    // a = state, b = K, c = jump target (after G_FILT_PRE_END)
    G_FILT_PRE_END,  // a = state, b = K, c = J
    G_FILT_BEGIN,  // a = state, b = filter table index, c = jump target (G_FILT_RUN)
    G_FILT_COEF,   // a = slot
    G_FILT_RUN,    // a = state, b = filter table index, c = 1 when G_FILT_BEGIN preceded
    G_FIN_HEAD,    // a = goe table index, c = jump target (static path)
    G_FIN_SCAN,    // c = jump target (join)
    G_FIN_STATIC,  //
    G_FIN_INNER,   // c = jump target (G_FIN_ADV)
    G_FIN_ADV,     // a = 1: an empty advance may be skipped, c = jump target (past G_FIN_END)
    G_FIN_END,     //
    G_APP_BEGIN,   // a = state, c = jump target (G_APP_MID)
    G_APP_MID,     // a = state, b = slot, c = jump target (just past G_APP_END)
    G_APP_END,     // b = slot
    G_RESET_BEGIN, // a = state, b = origin slot, c = jump target (G_RESET_END)
    G_RESET_END,   // a = state
    G_RUNS_BEGIN,  // Reset run by run: a = state, b = origin slot, c = its G_RUNS_END
    G_RUNS_END,    // a = origin slot, b = result slot, c = first instruction of the inner tree; followed by
                   // data words {a = first state word, b = words, c = 1 on the last}: what a restart clears
    G_SAVE,        // a = slot        acc -> slot (with its length)
    G_RESTORE,     // a = slot
    // ---- length ----
    L_INF,
    L_TIME,        // a = state
    L_FIXED,       // a = state, b = fixed table index
    L_PUSH,
    L_MIN,
    L_MAX,
    L_POP,
    L_FILT_HEAD,   // a = state, b = K, c = J
    L_FILT_MID,    // c = jump target (L_FILT_END)
    L_FILT_END,
    L_APP_BEGIN,   // a = state, c = jump target (L_APP_MID)
    L_APP_MID,     // a = state
    L_APP_END,
    L_FIN_SCAN1,
    L_FIN_SCAN2,   // c = jump target (join)
    L_FIN_STATIC,
    // ---- segmented (under Reset) ----
    S_CONST,
    S_TIME,
    S_FIXED,
    S_BINC,
    S_BIN_BEGIN,   // a = slot
    S_BIN_END,     // a = slot, b = operator, c = merge
    S_SINE_CC,
    S_SINE_AC,
    S_SINE_CA,
    S_SINE_BEGIN,  // a = slot
    S_SINE_END,    // a = state, b = slot
    S_ALT_CC,
    S_ALT_BEGIN,   // a = trigger slot
    S_ALT_POS,     // a = slot
    S_ALT_END,     // a = trigger slot, b = positive operand, c = negative operand
    S_RESET_BEGIN, // a = state, b = origin slot
    S_RESET_END,   // a = state
    S_FIN,         // a = goe table index   (static Time / const forms only)
    S_NOISE,       // a = state, b = node index (noise is not restarted by a Reset)
    S_APP_BEGIN,   // a = goe table index of the first part's Fin: remembers its local time at the window start
    S_APP_MID,     // a = slot (first part), b = origin slot of the second part, c = goe table index
    S_APP_END,     // a = slot (first part), b = origin slot, c = first state word | (state words << 16) of the second part
    // ---- steady-state stream (steady.cuh): straight-line, every operand infinite ----
    ST_END,
    ST_CONST,      // a = cval index
    ST_TIME,       // a = state
    ST_NOISE,      // a = state, b = node index
    ST_SAVE,       // a = slot                      acc -> slot
    ST_BIN,        // a = slot, b = operator        acc = slot (op) acc
    ST_SINE_CC,    // a = state, b = aux of the AUX_SINE_ROT block, c = aux of phase
    ST_SINE_AC,    // a = state, c = aux of phase           (frequency in acc); class in op>>8
    ST_SINE_CA,    // a = state, b = aux of increment       (phase in acc)
    ST_SINE_AA,    // a = state, b = slot of the frequency  (phase in acc)
    ST_ALT_CC,     // a = cval (positive), b = cval (negative)
    ST_ALT,        // a = trigger slot, b = positive operand, c = negative operand (>= 0: in acc)
    ST_FILT,       // a = state, b = aux of the coefficient values, c = aux of the matrix powers;
                   // K in op bits 8-11, J in bits 12-14
    ST_AFFINE,     // post-op word: acc = (acc * cval[b]) + cval[c], both operations rounded
    ST_OPC,        // post-op word: acc = acc (operator a) cval[b]
    // ---- steady words only the lane kernels run (a Reset over a tree whose nodes are closed-form in the run's
    //      own time: the sawtooth / pulse / triangle oscillators of lib/v0/std.tuun) ----
    ST_RESET_CLK,  // a = state (sign word), b = slot: acc = trigger -> slot = per-sample local clock, see lanes.cuh;
                   // c = clock slot of the Reset it is nested in (restarts with it), or -1
    ST_TIME_CLK,   // a = state, b = clock slot
    ST_SINE_CLK,   // a = state, b = aux of increment, c = aux of phase; clock slot in op bits 24-31; class in op >> 8
    // ---- a timeline in the steady stream (lane kernels only): Append(Fin{c0, e0}, Append(Fin{c1, e1}, ..)) under a
    //      root Fin, every length a literal and every e_k closed-form in its own clock — the envelopes of
    //      lib/v0/std.tuun (ADSR: four linear ramps).  All pieces are evaluated, the one a sample lies in is kept ----
    ST_SEG_CLK,    // a = W of the timeline's position, b = clock slot, c = first sample of the piece:
                   // slot = samples since the piece began (0 before it begins); two words: the second (op bit 9) holds
                   // a = first sample behind the piece, b = words to jump for a tile outside the piece (past its ST_SEG_SEL; to it for the last piece)
    ST_SEG_SEL,    // a = W of the position, b = slot (the pieces before), c = first sample of the piece:
                   // acc = at or after c ? acc : slot; op bit 8: the last piece (the position advances by a tile)
    // ---- lane program only (lanes.cuh; fused by lower.cpp build_lane_plan) ----
    LN_FM,         // ST_SINE_CC + one ST_AFFINE + ST_SINE_AC (or ST_SINE_CA) [+ ST_FILT K=3 J=2], two words:
                   //   word 0: a = W of the carried (sin, cos), b = W of the carrier's accumulator,
                   //           c = W of its phase offset (TB_LN_FM_PHASE: of its increment); op bits 8-15
                   //           rotation Q units, 24-31 carrier class | TB_LN_FM_PHASE
                   //   word 1: a, b = cval of the affine map; c = W of the filter state or -1, op = W of its coefficients
    OP_COUNT
};

// Operand encoding for instructions that take "a waveform that may be constant":
//   >= 0  shared-memory slot index;   < 0  ~cval index.
#define TB_OPERAND_CONST(k) (~(int32_t)(k))

#define TB_LN_FM_PHASE 0x40u  // LN_FM: the scaled sine is the carrier's phase (ST_SINE_CA), not its frequency

// Sine precision classes (op >> 8 of the G_SINE_* / S_SINE_* instructions).
#define TB_SINE_EXACT 0u  // f64 polynomial, rounds like the reference's (f64 sin) as f32
#define TB_SINE_FAST 1u   // f32 polynomial (or MUFU, tb_launch::fast_mode == 2); only for sines that
                          // feed no phase, trigger or length

// Constant-table construction, evaluated per voice at kernel start (generator.rs:574-612 is_const
// folding done once instead of once per block).
enum tb_cexpr_kind : uint32_t { CE_LIT = 0, CE_PARAM = 1, CE_BIN = 2, CE_NEG = 3 /* -cval[a] */ };
struct tb_cexpr {
    uint32_t kind;
    uint32_t op;   // CE_BIN: tb_operator
    int32_t a, b;  // CE_PARAM: a = column; CE_BIN: cval indices
    float value;   // CE_LIT, and CE_PARAM default
};

// Per-voice derived 64-bit constants ("aux"), evaluated after the constant table.
enum tb_aux_kind : uint32_t {
    AUX_SINE_INC = 0,    // cval[a] rad/s  -> phase increment, 2^-64 turns per sample
    AUX_SINE_PHASE = 1,  // cval[a] rad    -> phase offset, 2^-64 turns
    AUX_FILT_POW = 2,    // feedback coefficients of filter table b -> 6 JxJ f64 matrices A^(8 * 2^k)
    AUX_FILT_COEF = 3,   // constant coefficients of filter table b as K + J floats (steady stream)
    AUX_SINE_ROT = 4     // cval[a] rad/s  -> [0] phase increment; [2 .. 2 + 2*TB_CS) the rotations
                         // (cos, sin)(2 pi j inc / 2^64), j < TB_CS, as doubles (steady stream)
};
struct tb_aux {
    uint32_t kind;
    int32_t a, b;
    uint32_t off;  // offset in 64-bit words inside the per-warp aux area
};

// Lane-per-voice rendering of the steady stream (lanes.cuh): one THREAD owns one voice, so each
// thread keeps its voice's constants, state and derived constants in its own column of shared
// memory: 32-bit word w of thread t at W[w * TB_LANE_THREADS + t], 16-byte unit q at
// Q[q * TB_LANE_THREADS + t] (conflict free).  W = [cval | state | derived].  The lane code is the
// ST_* stream with every operand rewritten to a W word index (lower.cpp build_lane_plan).
enum tb_lane_aux_kind : uint32_t {
    LA_INC = 0,    // cval[a] rad/s -> phase increment, 2 W words
    LA_PHASE = 1,  // cval[a] rad   -> phase offset, 2 W words
    LA_ROT = 2,    // cval[a] rad/s -> increment (2 W words), then the carried (sin, cos) of the phase at the
                   // centre of the coming tile as two doubles (4 W words; b = W index of the node's
                   // accumulator, c = cval of its phase offset), + the rotations (cos, sin)(2 pi k inc / 2^64),
                   // k = 1..TB_LS/2 and k = TB_LS, as double2 in TB_LS/2 + 1 Q units
    LA_COEF = 3,   // filter table a: K feed-forward then J feedback coefficient values, K + J W words
    // a timeline (ST_SEG_*): its position is that of the root Fin's clock when a launch begins; when it ends the
    // nodes of the pieces get the state the general interpreter would have left (they share state blocks)
    LA_TL_POS = 4,   // a = W of the root Fin's Time: position -> 2 W words at w_off
    LA_TL_PIECE = 5, // a = first sample of the piece, b = W of its Fin's Time or -1, c = W of its Append's word or
                     // -1, q_off = first sample of the next piece, w_off = W of the position
    LA_TL_ZERO = 6   // a = first sample of the piece, b = W of a state range of it, c = words: cleared while the
                     // piece has not begun; w_off = W of the position
};
struct tb_lane_aux {
    uint32_t kind;
    int32_t a, b, c;
    uint32_t w_off, q_off;
};

// greater_or_equals_at chain (generator.rs:787-862), flattened.
enum tb_goe_term : uint32_t { GOE_TIME = 0, GOE_CONST = 1, GOE_MAYBE = 2 };
struct tb_goe {
    uint32_t term;        // tb_goe_term
    int32_t term_arg;     // GOE_TIME: state offset; GOE_CONST: cval index
    uint32_t through_append;  // a `None` answer degrades to `Maybe` (generator.rs:821-836)
    uint32_t n_steps;
    uint32_t step_off;    // into the goe step table: pairs (sign, cval index): value -= / += cval
};

// Time-axis split (split.cu, abi.cpp render_split_round): a steady program's carried state at ANY sample offset
// follows from per-segment summaries, so one voice can be rendered as S independent segments ("virtual
// voices") once every segment knows its initial state.  One entry per stateful node of the steady stream:
//   SP_POS         Time / Noise position: + 1 per sample                                   (analytic)
//   SP_SINE_CONST  Sine with a constant rate: accumulator + inc * n, exact in 2^-64 turns  (analytic)
//   SP_SINE_VAR    Sine with a rate waveform: the accumulator is a sum of increments — a segment's
//                  (final - initial) accumulator is its summary, exclusive prefix sums give the starts
//                  (u64, exact and associative: bit-identical to the unsplit render)
//   SP_FILTER      constant-coefficient Filter: the state (K-1 inputs, J outputs) obeys
//                  X' = M^L X + z, so z = final - M^L initial is a segment's summary and a scan of affine
//                  maps over the segments gives the starts (f64; differs from the serial f32 recurrence
//                  by its round-off noise)
// `level` = the render pass after which the entry's summary is right (its inputs were rendered from
// right states in that pass); analytic entries have level 0.  A program needs `split_passes` passes, the
// last of which writes the samples.
//   SP_RESET_SIGN  Reset (generator.rs:273-318): the class of the trigger's last sample — a segment starts from the
//                  class its predecessor ended in (a copy, right once the trigger is)
//   SP_CLK         a clocked node under a Reset (Time, or a Sine with constant rate and phase, whose state is the
//                  run's local clock x its rate, lanes.cuh ST_TIME_CLK / ST_SINE_CLK): a segment either restarts
//                  the clock — its final value is then absolute — or advances it by rate x L.  Which of the two
//                  happened shows in (final - initial); a "last set" scan over the segments gives the starts.
//                  (A restart on a segment's very first sample depends on the class the segment STARTED from, which
//                  was a guess in that pass: the Reset's second state word records the class of the first sample a
//                  launch saw, which lets the scan undo a spurious restart there or supply a missed one.)
enum tb_split_kind : uint32_t { SP_POS = 0, SP_SINE_CONST = 1, SP_SINE_VAR = 2, SP_FILTER = 3, SP_RESET_SIGN = 4, SP_CLK = 5 };
struct tb_split_entry {
    uint32_t kind;
    uint32_t state_off;  // first word of the node's state block
    uint32_t level;
    int32_t a;           // SP_SINE_CONST: cval of the rate; SP_FILTER: filter table index; SP_CLK: cval of the rate, -1 = Time
    int32_t b;           // SP_CLK: state block of the Reset whose clock it is
};

struct tb_filter_tab {
    uint32_t K, J;
    uint32_t all_const;    // every coefficient literally Const (generator.rs:428-440)
    uint32_t fb_const;     // every feedback coefficient is a constant operand -> scan path
    uint32_t pow_aux;      // aux offset of the matrix powers (fb_const && J > 0)
    uint32_t state_off;    // the node's state block
    int32_t x_slot;        // slot holding the zero-extended input when not all-const
    int32_t u_slot;        // scratch slot for the serial feedback fallback
    int32_t coef[TB_MAX_K_GEN + TB_MAX_J_GEN + 3];  // K feed-forward then J feedback operands
};

struct tb_fixed_tab {
    uint64_t off, len;
};

// Everything the kernel needs, passed by value.
struct tb_launch {
    const tb_insn* code;
    uint32_t n_code;
    uint32_t pc_gen, pc_len;  // entry points of the root's G_* and L_* programs
    uint32_t pc_steady;       // entry point of the ST_* stream (valid when steady_ok)
    const tb_cexpr* cexpr;
    uint32_t n_cval;
    const tb_aux* aux;
    uint32_t n_aux, aux_words;
    const tb_goe* goe;
    const int32_t* goe_steps;
    const tb_filter_tab* filt;
    const tb_fixed_tab* fixed;
    const float* pool;
    uint32_t n_slots, state_words;
    uint32_t sample_rate;
    uint32_t n_filt;
    uint32_t steady_ok;    // the generate program may run through the steady-state interpreter
    // lane-per-voice plan (valid when lane_ok)
    const tb_insn* lane_code;
    const tb_lane_aux* lane_aux;
    uint32_t n_lane_code, n_lane_aux;
    uint32_t lane_w_words, lane_q_units, lane_slots;
    uint32_t* fault;       // device counter: voices the lane kernel found without complete filter history
    uint32_t lane_clk;     // the steady stream holds ST_*_CLK words: the warp-per-voice steady interpreter must not run it
    int32_t lane_fin_goe;  // lane kernel: the program is Fin{analytic length, steady tree}; goe entry of the length, else -1
    uint32_t mid_call;     // this launch continues a generate call (Fin keeps the cut the call's first tile made)
    uint64_t call_pos;     // samples of the current call rendered before this launch (lane kernel: voices whose
                           // out_len is below it returned short earlier in the call)
    uint32_t* lane_queue;  // lane kernel work queue ([0] unit counter, [1 + g] segments finished by group g) or NULL
    uint32_t lane_groups, lane_segs, lane_grid;  // voice groups, time segments, persistent CTAs
    uint64_t lane_seg_samples;                   // samples per segment (a multiple of 2 * TB_LS)
    float* mix_partial;    // lane kernel, mixdown without rows: [ceil(n_voices / 32)][mix_stride] sums of 32 voices
    uint64_t mix_stride;
    uint32_t fast_mode;    // FAST-class sine evaluation: 1 = f32 polynomial, 2 = MUFU
    unsigned long long noise_seed;   // tb_seed_noise
    unsigned long long voice_base;   // index of voice 0 of this launch inside the caller's batch
    // per call
    const float* params;
    uint32_t n_params, n_voices;
    uint64_t n_samples;
    float* out;
    uint64_t out_stride;
    unsigned long long* out_len;
    uint32_t* state;       // [n_voices][state_words]
    uint32_t mode;         // 0 = generate, 1 = length only
    uint32_t pure_len;     // length program contains no G_* code
    uint32_t accumulate;   // out_len[v] += (chunked host-output renders) instead of =
    uint32_t exact_fb;     // constant-coefficient feedback by the serial recurrence instead of the scan
    uint8_t* done;         // [n_voices] or NULL: voices that already returned short in this call
    // time-axis split: `n_voices` counts virtual voices; virtual voice vv is segment (vv % vsplit) of real
    // voice (vv / vsplit): parameters, noise streams and the output row belong to the real voice, the row
    // starts vseg samples further per segment; state and out_len are per virtual voice.
    // A launch may cover a RANGE of a voice's segments (time sharding over GPUs: every rank renders its own
    // range): segment sl of the launch is segment vseg_lo + sl of the voice, whose state block is number
    // v * vsplit_total + vseg_lo + sl.
    uint32_t vsplit;       // segments per voice in this launch; 0 or 1: every voice is a real voice
    uint32_t vsplit_total; // segments per voice in all (>= vseg_lo + vsplit)
    uint32_t vseg_lo;      // first segment of this launch
    uint64_t vseg;         // samples per segment
    uint32_t state_only;   // render for the final state alone (a summary pass): out == NULL is not "mixdown"
    uint32_t fm_sums;      // fused FM voice, summary pass: the carrier's phase sum alone (lanes.cuh run_fm_sums)
    unsigned long long* vsnap;  // ... which records the accumulator after vsnap_at samples in vsnap[state block]
    uint64_t vsnap_at;
    double lane_kscale;    // 2^44 / (TAU * sample_rate) (common.cuh SineK::kscale), made by the host: an operand straight out of
                           // the constant bank instead of two registers of the warps that live on 64 (lanes_fm_ws.cu)
};
