// lower.h — result of lowering a tb_node op list to the warp byte-code (see lower.cpp).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/tuun_b200.h"
#include "program.h"

namespace tb {

struct Lowered {
    std::vector<tb_insn> code;
    uint32_t pc_gen = 0, pc_len = 0, pc_steady = 0;
    std::vector<tb_cexpr> cexpr;
    std::vector<tb_aux> aux;
    uint32_t aux_words = 0;
    std::vector<tb_goe> goe;
    std::vector<int32_t> goe_steps;
    std::vector<tb_filter_tab> filt;
    std::vector<tb_fixed_tab> fixed;
    uint32_t n_slots = 0, state_words = 0, n_params = 0, n_nodes = 0;
    uint32_t pure_len = 1;
    uint32_t steady_ok = 0;  // the generate program runs through the steady-state interpreter too
    // lane-per-voice plan of the steady stream (lanes.cuh); lane_ok = 0 when it does not apply
    uint32_t lane_ok = 0;
    std::vector<tb_insn> lane_code;
    std::vector<tb_lane_aux> lane_aux;
    uint32_t lane_w_words = 0, lane_q_units = 0, lane_slots = 0;
    uint32_t lane_clk = 0;      // the steady stream uses the clocked words (Reset): lane kernels only
    int32_t lane_fin_goe = -1;  // root Fin with an analytic length over a steady tree: its goe entry, else -1
    // time-axis split plan (program.h tb_split_entry); split_passes = 0 when the program does not qualify
    std::vector<tb_split_entry> split;
    uint32_t split_passes = 0;
    int status = 0;
    std::string error;
};

// Returns TB_OK or a negative tb_status (out.error holds the reason).  noise_ids (may be NULL): node -> number of
// its Noise stream, for an op list that is a part of a larger tree.
int lower(const tb_node* nodes, uint32_t n_nodes, const int32_t* lists, uint32_t n_lists,
          uint64_t pool_len, bool fast_sines, Lowered& out, const uint32_t* noise_ids = nullptr,
          uint32_t sample_rate = 0);  // sample_rate 0: not known (lengths in samples are not decided while lowering)

// A root SEQUENCE — a tree of Appends (nested either way, under Marked / Captured wrappers) whose leaves are
// Fin{len_0, a_0}, Fin{len_1, a_1}, ..., rest: what `<[a, b, c]>`, `a \ b` and Player::beats_waveform evaluate to
// (builtins.rs:208-299, optimizer.rs:212-229, player.rs:232-260) — with Fin lengths that are analytic and the same
// for every voice (Time +- literal constants, generator.rs:787-862) over waveforms that cannot end earlier.  The second
// arm of an Append starts from its own Initial state (generator.rs:169-188), so every part is an independent stream
// that begins at a known sample: parts[i] = {root node of the part (the Fin with its wrappers; the rest for the last
// one), its length in samples (~0 for the last)}.  False when the root is no such sequence (or has more than 4096 parts).
struct SeqPart {
    int root;
    uint64_t len;
};
bool sequence_parts(const tb_node* nodes, uint32_t n_nodes, const int32_t* lists, uint32_t n_lists, uint64_t pool_len,
                    uint32_t sample_rate, std::vector<SeqPart>& parts);

}  // namespace tb
