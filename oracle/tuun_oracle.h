/*
 * tuun_oracle.h — CPU oracle for Tuun's waveform-generation hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A scalar restatement of the reference generator (src/lib/generator.rs:86-862) and of the
 * state helpers it needs (src/lib/waveform.rs:179-392), consuming the same `tb_node` op list
 * as the product C ABI (include/tuun_b200.h) so a parity test can feed ONE program to both.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; nothing under tuun_b200/ links or calls it.
 *
 * Parity status: pinned against the reference's own known-answer vectors
 * (generator.rs:1353-1925, chunk sizes 1/2/4/8) by tests/test_oracle_golden.py.
 * `Noise` is "parity unpinned": the reference draws from fastrand 2.3.0's unseeded
 * thread-local generator (generator.rs:115, Cargo.lock:372), which no test pins; the oracle
 * restates that crate's published generator (wyrand) and gives every Noise node of every voice
 * its own seeded stream of it (tuun_oracle.cpp, struct Rng) — same distribution, reproducible.
 * The Rust reference itself cannot be compiled here (no cargo/rustc), so there is no
 * oracle/_ref build; libm is glibc's instead of Rust std's (same correctly-rounded-to-<1ulp
 * f64 sin; test_sine pins it to 1e-5 only).
 */
#ifndef TUUN_ORACLE_H
#define TUUN_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "../include/tuun_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tbo_program tbo_program;

/* initialize_state over the op list (generator.rs:39).  Returns 0 or a tb_status. */
int tbo_program_create(const tb_node* nodes, uint32_t n_nodes, const int32_t* lists,
                       uint32_t n_lists, const float* fixed_pool, uint64_t fixed_len,
                       uint32_t sample_rate, tbo_program** out_program);
void tbo_program_destroy(tbo_program* p);

/* Bind one voice's row of the parameter table (Const nodes with param_slot >= 0). */
int tbo_set_params(tbo_program* p, const float* params, uint32_t n_params);

/* Generator::generate (generator.rs:86): one call == one reference call on `out[..n]`. */
uint64_t tbo_generate(tbo_program* p, float* out, uint64_t n);
/* Generator::length (generator.rs:620). */
uint64_t tbo_length(tbo_program* p, uint64_t max);
/* waveform::set_state(root, Initial) then nothing else (waveform.rs:322; keeps quirk A10). */
void tbo_set_state_initial(tbo_program* p);
/* initialize_state: every node incl. filter coefficients back to Initial (waveform.rs:179). */
void tbo_initialize_state(tbo_program* p);
/* waveform::substitute for Marked nodes whose id matches (waveform.rs:397): replaces the
 * marked subtree by Const(value) — the only form the reference tests use (generator.rs:1424). */
int tbo_substitute_const(tbo_program* p, uint32_t mark_id, float value);
/* Generator.allocations (generator.rs:53). */
uint64_t tbo_allocations(const tbo_program* p);
void tbo_seed_noise(tbo_program* p, uint64_t seed);
/* Zero the scratch-buffer tails the reference reads past a producer's returned length (see Gen::clean_tails). */
void tbo_set_clean_tails(tbo_program* p, int on);
/* Voice index of this program inside a batch (selects the Noise streams). */
void tbo_set_voice(tbo_program* p, uint64_t voice);

/*
 * The reference's offline shape (benches/tracker_benches.rs:19-34, tracker.rs:597-642): for
 * every voice, initialize_state, bind params[v], then call generate on `block`-sample buffers
 * until n_samples are produced or the voice finishes.  Voices are spread over `n_threads`
 * host threads.  out may be NULL (render into a per-thread scratch block and discard — the
 * bench shape); mix may be NULL, otherwise mix[i] += out[v][i] in voice order per thread and
 * threads combined in thread order.  Returns total samples generated.
 */
uint64_t tbo_render_batch(const tbo_program* p, const float* params, uint32_t n_params,
                          uint32_t n_voices, uint64_t n_samples, uint32_t block, float* out,
                          uint64_t out_stride, uint64_t* out_len, float* mix, uint32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif
