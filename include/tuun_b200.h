/*
 * tuun_b200.h — C ABI of the B200-native renderer for Tuun's waveform-generation hot path.
 *
 * This is the drop-in boundary for the reference path
 *     generator::initialize_state            (src/lib/generator.rs:39)
 *     Generator::new / generate / length     (src/lib/generator.rs:68, :86, :620)
 *     waveform::set_state(.., Initial)       (src/lib/waveform.rs:322)
 *     tracker mix loop  out[j] += tmp[j]     (src/lib/tracker.rs:597-642)
 * The reference has no FFI for this path (it is a Rust method on a value type), so the
 * entry points below are what a Rust `extern "C"` block would bind (see INTEGRATION.md and
 * rust/tuun-b200-sys/src/lib.rs).  Plain pointers and sizes only; no C++/torch types.
 *
 * A `Waveform<MarkId, State>` tree (src/lib/waveform.rs:23-100) crosses the ABI as a flat
 * array of `tb_node` in topological order (children before parents, root = last node) —
 * one node per enum variant instance, one-for-one, nothing re-associated or folded.
 *
 * All entry points return 0 (TB_OK) or a negative tb_status; they never unwind and never
 * fall back to a CPU implementation: without a usable CUDA device every compute call
 * fails with TB_ERR_CUDA.
 */
#ifndef TUUN_B200_H
#define TUUN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 3: tb_program_info grew (lane_fm_ws_capacity, fm_ws_launches); sample rates must be below 2^31 */
#define TB_ABI_VERSION 3u

/* enum Waveform variants, src/lib/waveform.rs:23-100 (same order). */
typedef enum tb_kind {
    TB_CONST = 0,    /* Const(f32)                                   waveform.rs:25 */
    TB_TIME = 1,     /* Time(State)                                  waveform.rs:27 */
    TB_NOISE = 2,    /* Noise                                        waveform.rs:29 */
    TB_FIXED = 3,    /* Fixed(Vec<f32>, State)                       waveform.rs:31 */
    TB_FIN = 4,      /* Fin{length=a, waveform=b}                    waveform.rs:35 */
    TB_APPEND = 5,   /* Append(a, b, State)                          waveform.rs:41 */
    TB_SINE = 6,     /* Sine{frequency=a, phase=b, state}            waveform.rs:51 */
    TB_FILTER = 7,   /* Filter{waveform=a, feed_forward, feedback}   waveform.rs:59 */
    TB_BINARY = 8,   /* BinaryPointOp(op, a, b)                      waveform.rs:67 */
    TB_RESET = 9,    /* Reset{trigger=a, waveform=b, state}          waveform.rs:74 */
    TB_ALT = 10,     /* Alt{trigger=a, positive=b, negative=c}       waveform.rs:81 */
    TB_MARKED = 11,  /* Marked{id=mark_id, waveform=a}               waveform.rs:90 */
    TB_CAPTURED = 12 /* Captured{file_stem=#mark_id, waveform=a}     waveform.rs:96 */
} tb_kind;

/* enum Operator, src/lib/waveform.rs:5-19 (same order). */
typedef enum tb_operator {
    TB_ADD = 0,
    TB_SUBTRACT = 1,
    TB_MULTIPLY = 2,
    TB_DIVIDE = 3, /* b == 0 yields 0                                generator.rs:266 */
    TB_MERGE = 4,  /* Add with zero-extension to the longer input    generator.rs:263 */
    TB_POWER = 5   /* f32::powf                                      generator.rs:269 */
} tb_operator;

/*
 * One node of the op list.  `a`, `b`, `c` are indices of child nodes (or -1).  For TB_FILTER
 * the coefficient waveforms are listed in the program's `lists` array:
 *   feed_forward[i] = lists[list_off + i]                 for i in 0..ff_count   (b_0, b_1, ...)
 *   feedback[j]     = lists[list_off + ff_count + j]      for j in 0..fb_count   (a_1, a_2, ...)
 * For TB_FIXED the samples are fixed_pool[fixed_off .. fixed_off + fixed_len].
 * For TB_CONST, `param_slot >= 0` makes the constant a per-voice parameter: voice v renders
 * with value params[v * n_params + param_slot] (a batch of identically shaped trees that
 * differ only in their constants); `value` is then the default used when no table is given.
 */
typedef struct tb_node {
    uint32_t kind;       /* tb_kind */
    uint32_t op;         /* tb_operator, TB_BINARY only */
    int32_t a, b, c;     /* children, -1 when unused */
    float value;         /* TB_CONST */
    int32_t param_slot;  /* TB_CONST: -1, or column of the parameter table */
    uint32_t list_off;   /* TB_FILTER */
    uint32_t ff_count;   /* TB_FILTER, >= 1 (generator.rs:233) */
    uint32_t fb_count;   /* TB_FILTER */
    uint32_t mark_id;    /* TB_MARKED / TB_CAPTURED */
    uint32_t reserved;   /* must be 0 */
    uint64_t fixed_off;  /* TB_FIXED */
    uint64_t fixed_len;  /* TB_FIXED */
} tb_node;

typedef enum tb_status {
    TB_OK = 0,
    TB_ERR_INVALID = -1,     /* malformed op list / bad argument (the reference would panic) */
    TB_ERR_UNSUPPORTED = -2, /* well-formed but outside what the device path implements */
    TB_ERR_CUDA = -3,        /* no device, kernel image missing, or a CUDA call failed */
    TB_ERR_NOMEM = -4,
    TB_ERR_STATE = -5        /* call sequence error (e.g. voice count changed mid-stream) */
} tb_status;

/* tb_render flags */
#define TB_OUT_DEVICE 1u   /* `out`/`mix` are device pointers (default: host, copied back) */
#define TB_PARAMS_DEVICE 2u /* `params` is a device pointer */
#define TB_NO_VOICE_OUT 4u /* tb_render_mix only: do not materialise per-voice rows */

typedef struct tb_program tb_program;

/*
 * initialize_state + ownership of the tree (generator.rs:39, waveform.rs:179).
 * Validates and lowers the op list to the device byte-code; all persistent render state
 * (the reference's per-node `State`, generator.rs:12-35) lives behind the handle.
 * `device` is the CUDA ordinal (-1 = current device).
 */
int tb_program_create(const tb_node* nodes, uint32_t n_nodes,
                      const int32_t* lists, uint32_t n_lists,
                      const float* fixed_pool, uint64_t fixed_len,
                      uint32_t sample_rate, int device, tb_program** out_program);

void tb_program_destroy(tb_program* p);

/*
 * Generator::generate over a batch of voices (generator.rs:86).
 * Renders the next `n_samples` samples of every voice into out[v * out_stride + i].
 * out_len[v] (may be NULL) receives the number of samples generated; a value smaller than
 * n_samples means voice v has finished, and out[v][out_len[v]..] is undefined, exactly as
 * the reference states (generator.rs:79-81).  State is carried, so successive calls continue
 * where the previous one stopped ("pick up where this one left off", generator.rs:76-78).
 * `params` may be NULL (use the constants in the tree); n_voices must stay the same for the
 * life of the stream (until tb_reset).
 *
 * Few voices, many samples: a steady program (sines, clocks, noise, point operators, Alt,
 * constant-coefficient filters) is cut along TIME as well — every voice becomes S segments rendered
 * side by side, each from the state the reference's serial walk would have reached there (phase sums:
 * exclusive u64 prefix sums of per-segment increments, exact; filter histories: a scan of affine maps
 * X' = M^L X + z over the segments).  Sample values of sines do not depend on S; filtered samples differ
 * from the serial f32 recurrence by its round-off noise.  Automatic when it pays; TUUN_B200_SPLIT=0
 * disables it, TUUN_B200_SPLIT=S forces S segments.  tb_program_info reports what happened.
 */
int tb_render(tb_program* p, const float* params, uint32_t n_params, uint32_t n_voices,
              uint64_t n_samples, float* out, uint64_t out_stride, uint64_t* out_len,
              uint32_t flags);

/*
 * tb_render plus the tracker's mix loop (tracker.rs:597-642): mix[i] = sum over voices of
 * out[v][i] for i < out_len[v], in f32.  With rows (`out` given) the adds run in voice index order
 * exactly like the tracker's serial `out[j] += tmp[j]` (tracker.rs:617-619).  With TB_NO_VOICE_OUT
 * (`out` may be NULL) a large steady batch never materialises its rows: each warp of the lane kernel
 * adds its 32 voices on the chip in voice order and the per-warp partial rows are added in warp
 * order (programs other than a single fused FM voice: the first 256-271 samples of a stream are
 * rendered by the general kernel and mixed in plain voice order).  That is the same sum
 * RE-ASSOCIATED in blocks of 32 voices ((v0 + .. + v31) + (v32 + .. + v63) + ..): deterministic,
 * equal to the serial order within f32 re-association (|difference| <= a few ulp of the mix x
 * log2(voices)), not bit-identical to it.  Callers that need the tracker's exact order pass rows.
 * `mix` holds n_samples floats, is overwritten, and lives where `out` lives (TB_OUT_DEVICE).
 */
int tb_render_mix(tb_program* p, const float* params, uint32_t n_params, uint32_t n_voices,
                  uint64_t n_samples, float* out, uint64_t out_stride, uint64_t* out_len,
                  float* mix, uint32_t flags);

/*
 * Generator::length (generator.rs:620): advance every voice by `max` samples without
 * producing output; len[v] = number of samples the voice would have generated.
 */
int tb_length(tb_program* p, const float* params, uint32_t n_params, uint32_t n_voices,
              uint64_t max, uint64_t* len, uint32_t flags);

/*
 * Noise (generator.rs:113-118) draws `fastrand::f32() * 2 - 1` from the reference's unseeded
 * thread-local generator, so the reference itself never produces the same noise twice.  Here
 * every Noise node of every voice owns a seeded stream of the same generator (fastrand 2.3.0 =
 * wyrand), a pure function of (seed, node index, voice index, sample count): renders are
 * reproducible, tb_reset replays them, and a batch split over several programs / GPUs draws what
 * the unsplit batch would when each part passes the index of its first voice.  Defaults: a fixed
 * seed, first_voice 0.
 */
int tb_seed_noise(tb_program* p, uint64_t seed, uint64_t first_voice);

/* waveform::set_state(root, State::Initial) for every voice (waveform.rs:322). */
int tb_reset(tb_program* p);

/*
 * waveform::substitute(&mut w, &mark_id, &Const(value)) (waveform.rs:396-462), mid-stream: the contents
 * of every Marked node whose id is `mark_id` become Const(value); every other node keeps its state and the
 * stream continues ("it picks up where it would have been", generator.rs:1398-1463 — which is why Fin
 * advances both of its children).  This is the form the reference itself uses: slider and amplitude values
 * are Marked constants replaced by constants (player.rs:110, :265-288).  `*n_replaced` (may be NULL) receives
 * the number of Marked nodes changed; 0 is not an error (substitute replaces "zero or more parts").
 * TB_ERR_UNSUPPORTED when a matching Marked node holds anything but a Const (the tree would change shape;
 * create a new program for that), or when the new value changes how the tree lowers.
 */
int tb_substitute(tb_program* p, uint32_t mark_id, float value, uint32_t* n_replaced);

/*
 * Time-segment sharding over several GPUs (one process per GPU): few voices, long renders.  A steady
 * program's carried state (generator.rs:12-35) has an associative form over time — positions, sums of phase
 * increments, affine maps of filter histories — so the next n_segments * seg_samples samples of every voice
 * can be rendered as n_segments segments by different devices: rank r renders segments [seg_lo, seg_hi) of
 * every voice; after each pass the ranks exchange the segments' final states (a few dozen bytes per stateful
 * node and segment: one all-gather over NVLink) and every rank runs the same scan over them.
 *
 *     tb_segments_begin(p, params, .., n_voices, n_segments, seg_samples, flags, &n_passes)
 *     for pass in 1 ..= n_passes:
 *         tb_segments_pass(p, pass, seg_lo, seg_hi, out, stride, TB_OUT_DEVICE)    // rows only on the last pass
 *         all-gather the state blocks of [seg_lo, seg_hi) (tb_segments_states)     // caller: NCCL / MPI / nothing
 *         if pass < n_passes: tb_segments_fix(p, pass)
 *     tb_segments_end(p)                       // the stream continues behind the last segment on every rank
 *
 * Every rank holds the same program, voices and stream position (call tb_render for the first 256 samples
 * of a stream on every rank: filters read ahead on their first call).  seg_samples is a multiple of 512.
 * States: a device array of n_voices * n_segments blocks of *bytes_per_segment bytes, block (v, s) at index
 * v * n_segments + s.  On the last pass `out` receives the rows of the rank's own segments:
 * out[v * out_stride + (s - seg_lo) * seg_samples + i].  With one rank and [0, n_segments) this is what
 * tb_render does by itself for few voices.  TB_ERR_UNSUPPORTED: the program is not steady.
 */
int tb_segments_begin(tb_program* p, const float* params, uint32_t n_params, uint32_t n_voices,
                      uint32_t n_segments, uint64_t seg_samples, uint32_t flags, uint32_t* n_passes);
int tb_segments_pass(tb_program* p, uint32_t pass, uint32_t seg_lo, uint32_t seg_hi, float* out,
                     uint64_t out_stride, uint32_t flags);
int tb_segments_states(tb_program* p, void** states, uint64_t* bytes_per_segment);
int tb_segments_fix(tb_program* p, uint32_t pass);
int tb_segments_end(tb_program* p);

/*
 * Stream the render is enqueued on (a cudaStream_t).  A program creates its own NON-BLOCKING stream: device-side work
 * of the caller that touches `out` / `params` / `mix` (a fill, a copy, a consumer kernel) is not ordered against a
 * render unless the caller orders it — wait for this stream (or an event on it), or hand the program the caller's own
 * stream with tb_set_stream.  Calls that return host data (out_len, host rows, tb_length) synchronize before returning;
 * with TB_OUT_DEVICE and no out_len a call returns as soon as its launches are queued.
 */
void* tb_stream(tb_program* p);
/* Run subsequent renders of this program on a caller-owned cudaStream_t. */
int tb_set_stream(tb_program* p, void* cuda_stream);

/* Introspection used by tests, bench.py and DESIGN.md numbers. */
typedef struct tb_program_info {
    uint32_t n_nodes;
    uint32_t n_code_words;    /* device byte-code length */
    uint32_t n_slots;         /* shared-memory tile slots per CTA */
    uint32_t state_words;     /* 32-bit words of carried state per voice */
    uint32_t tile;            /* samples per tile */
    uint32_t threads;         /* threads per CTA */
    uint32_t smem_bytes;      /* dynamic shared memory per CTA */
    uint32_t n_params;        /* highest param_slot + 1 */
    uint64_t kernel_launches; /* cumulative launches of this library's kernels */
    uint64_t lane_launches;   /* of those, launches of the lane-per-voice kernel (large steady batches) */
    uint32_t lane_smem_bytes; /* its dynamic shared memory per CTA; 0 when the program does not qualify */
    uint32_t lane_min_voices; /* batches of at least this many voices take it */
    uint32_t lane_capacity;   /* 64-voice CTAs of the lane interpreter kernels the device holds at once (more: work queue) */
    uint32_t lane_fm_capacity;/* same for the fused-FM-voice kernel; 0 when the program is not one fused FM voice */
    /* time-axis split (few voices, long calls: a voice rendered as S segments side by side, see tb_render) */
    uint32_t split_passes;    /* render passes a split call makes (summary passes + the one that writes samples); 0: never split */
    uint32_t split_segments;  /* S of the first (largest) round of the most recent split call; 0: no call was split yet */
    uint64_t split_seg_samples; /* samples per segment of that round */
    uint64_t split_rounds;    /* split rounds so far */
    /* a root sequence — Append(Fin{len, a}, Append(Fin{len', b}, ..)) with analytic, voice-independent lengths: what
       `<[a, b, ..]>` evaluates to — is held as one program per part, each rendered where it starts */
    uint32_t sequence_parts;  /* 0: the tree is one program */
    uint32_t split_fm_rounds; /* of split_rounds: those in the form for fused FM voices (phase-sum pass, filter warm-up, samples) */
    uint64_t sequence_renders; /* generate launches that went part by part */
    uint32_t lane_fm_ws_capacity; /* 32-voice CTAs of the fused-FM-voice kernel's two-warps-a-voice form the device holds at once; 0: not applicable */
    uint32_t reserved0;
    uint64_t fm_ws_launches;  /* of lane_launches: those of that form (tb_render_lanes_fm_ws_kernel) */
} tb_program_info;
int tb_program_get_info(const tb_program* p, tb_program_info* info);

/*
 * Device durations (milliseconds, CUDA events on the program's stream) of the most recent launches
 * of the lane-per-voice kernel, oldest first; at most `cap` (and at most 64) values, `*n` of them
 * written.  Synchronizes the stream.  Measurement aid for bench.py's roofline line: the kernel's
 * own launch time, apart from the two small bracketing launches of a call.
 */
int tb_lane_kernel_times(tb_program* p, float* ms, uint32_t cap, uint32_t* n);

/*
 * Validate and lower an op list WITHOUT touching a device: the status a tb_program_create of the
 * same tree would return from its host half (TB_ERR_INVALID / TB_ERR_UNSUPPORTED with
 * tb_last_error(), else TB_OK), and the launch geometry it would use.  The reference's analogue is
 * the panics of initialize_state / generate on a malformed tree (generator.rs:112,132,189,...),
 * surfaced before any sample is produced.  `tile` reports 512 when the tree qualifies for the
 * steady-state interpreter, else 256.  (This entry point has no sample rate: where a length in
 * samples decides the lowering — an envelope timeline under a note, see DESIGN.md section 2c — it is
 * taken at 44,100 Hz, the rate of the reference's tracker.)
 */
int tb_lower_check(const tb_node* nodes, uint32_t n_nodes, const int32_t* lists, uint32_t n_lists,
                   uint64_t fixed_len, tb_program_info* info);

/* Human-readable text of the last error on the calling thread (never NULL). */
const char* tb_last_error(void);
uint32_t tb_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TUUN_B200_H */
