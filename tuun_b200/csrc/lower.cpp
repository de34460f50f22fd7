// lower.cpp — host lowering of the tb_node op list (a reference Waveform tree,
// src/lib/waveform.rs:23-100) to the warp byte-code of program.h.
//
// The emitters follow the reference's three tree walks:
//   emit_gen  <- Generator::generate   (generator.rs:86-380)
//   emit_len  <- Generator::length     (generator.rs:620-782)
//   emit_seg  <- generate re-entered run by run under a Reset (generator.rs:281-318)
// Everything the reference decides from the SHAPE of the tree is decided here once:
// is_const (generator.rs:574-612), the all-constant-coefficient test of Filter (:428-440), the
// form of a Fin length (greater_or_equals_at, :787-862).  Everything it decides from VALUES or
// STATE (lengths, Append hand-over, restarts) stays in the byte-code as warp-uniform control.
#include "lower.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>

namespace tb {

namespace {

struct Lowerer {
    const tb_node* nodes;
    uint32_t n_nodes;
    const int32_t* lists;
    uint32_t n_lists;
    uint64_t pool_len;
    Lowered& out;
    bool fast_sines;
    const uint32_t* noise_ids = nullptr;  // node -> number of its Noise stream (a part of a sequence keeps the numbers
                                          // its nodes have in the whole tree, lower.h sequence_parts); NULL: the index
    int noise_id(int i) const { return noise_ids ? (int)noise_ids[i] : i; }
    uint32_t sample_rate = 0;  // 0: not known (no timeline in the steady stream)

    std::vector<int> const_memo;   // node -> cval index, -2 = not computed, -1 = not const
    std::vector<int> state_off;    // node -> offset of its state block (or -1)
    std::vector<int> state_len;    // node -> words of that block
    std::vector<int> fixed_idx;    // node -> fixed table index
    std::vector<int> filt_idx;     // node -> filter table index
    std::vector<uint8_t> sensitive;  // node output reaches a phase / trigger / length / coefficient
    int slots_in_use = 0;

    Lowerer(const tb_node* n, uint32_t nn, const int32_t* l, uint32_t nl, uint64_t pl, Lowered& o, bool fs)
        : nodes(n), n_nodes(nn), lists(l), n_lists(nl), pool_len(pl), out(o), fast_sines(fs) {}

    [[noreturn]] void fail(int status, const std::string& msg) {
        out.status = status;
        out.error = msg;
        throw status;
    }

    int alloc_slot() {
        int s = slots_in_use++;
        out.n_slots = std::max<uint32_t>(out.n_slots, (uint32_t)slots_in_use);
        return s;
    }
    void free_slot() { slots_in_use--; }

    int emit(uint32_t op, int a = 0, int b = 0, int c = 0) {
        out.code.push_back(tb_insn{op, a, b, c});
        return (int)out.code.size() - 1;
    }
    int here() const { return (int)out.code.size(); }
    // A jump lands on the next instruction to be emitted: nothing emitted from now on may be fused
    // backwards into what precedes the label.
    int label_at = -1;
    int last_producer = -1;  // index of the last instruction that leaves a fresh accumulator
    int label() {
        label_at = here();
        return label_at;
    }
    void produced(int idx) { last_producer = idx; }
    // acc = acc (op) const as a post-op of the producing instruction when that is legal.
    void emit_binc(int op, int cidx) {
        const int h = here();
        if (last_producer >= 0 && label_at != h) {
            const int npost = (int)(out.code[last_producer].op >> 16);
            if (last_producer + 1 + npost == h && npost < 15) {
                out.code[last_producer].op += 1u << 16;
                out.code.push_back(tb_insn{(uint32_t)G_BINC, op, cidx, 0});
                return;
            }
        }
        produced(emit(G_BINC, op, cidx, 0));
    }

    // ---- validation -------------------------------------------------------------------
    void validate() {
        if (!nodes || n_nodes == 0) fail(TB_ERR_INVALID, "empty op list");
        if (!lists && n_lists > 0) fail(TB_ERR_INVALID, "lists is NULL but n_lists > 0");
        // A Waveform owns each child exactly once (Box, waveform.rs:23-100): the op list must be a TREE rooted in its last
        // node.  A node with two parents would share one state block between them, an unreachable one is not part of
        // the waveform at all.
        std::vector<uint32_t> parents(n_nodes, 0);
        auto child_ok = [&](int32_t c, uint32_t self) {
            if (c < 0 || (uint32_t)c >= self) return false;
            parents[c]++;
            return true;
        };
        for (uint32_t i = 0; i < n_nodes; i++) {
            const tb_node& n = nodes[i];
            if (n.reserved != 0) fail(TB_ERR_INVALID, "tb_node.reserved must be 0");
            bool ok = true;
            switch (n.kind) {
                case TB_CONST:
                    if (n.param_slot >= 0) out.n_params = std::max<uint32_t>(out.n_params, n.param_slot + 1);
                    break;
                case TB_TIME:
                case TB_NOISE: break;
                case TB_FIXED: ok = n.fixed_len <= pool_len && n.fixed_off <= pool_len - n.fixed_len; break;  // no u64 wrap
                case TB_FIN:
                case TB_APPEND:
                case TB_SINE:
                case TB_RESET: ok = child_ok(n.a, i) && child_ok(n.b, i); break;
                case TB_BINARY: ok = n.op <= TB_POWER && child_ok(n.a, i) && child_ok(n.b, i); break;
                case TB_ALT: ok = child_ok(n.a, i) && child_ok(n.b, i) && child_ok(n.c, i); break;
                case TB_MARKED:
                case TB_CAPTURED: ok = child_ok(n.a, i); break;
                case TB_FILTER:
                    ok = child_ok(n.a, i) && n.ff_count >= 1 &&
                         (uint64_t)n.list_off + n.ff_count + n.fb_count <= n_lists;
                    if (ok)
                        for (uint32_t j = 0; j < n.ff_count + n.fb_count; j++)
                            ok = ok && child_ok(lists[n.list_off + j], i);
                    break;
                default: ok = false;
            }
            if (!ok) fail(TB_ERR_INVALID, "malformed node " + std::to_string(i));
        }
        for (uint32_t i = 0; i + 1 < n_nodes; i++)
            if (parents[i] != 1)
                fail(TB_ERR_INVALID, "node " + std::to_string(i) + (parents[i] == 0 ? " is not reachable from the root"
                                                                                     : " has more than one parent") +
                                         ": the op list must be a tree (one node per Waveform variant instance)");
    }

    // ---- is_const (generator.rs:574-612), evaluated per voice through the cexpr table ----
    int new_cexpr(const tb_cexpr& e) {
        out.cexpr.push_back(e);
        return (int)out.cexpr.size() - 1;
    }
    int const_of(int i) {
        if (const_memo[i] != -2) return const_memo[i];
        const tb_node& n = nodes[i];
        int r = -1;
        switch (n.kind) {
            case TB_CONST:
                if (n.param_slot >= 0) r = new_cexpr(tb_cexpr{CE_PARAM, 0, n.param_slot, 0, n.value});
                else r = new_cexpr(tb_cexpr{CE_LIT, 0, 0, 0, n.value});
                break;
            case TB_BINARY: {
                int a = const_of(n.a), b = const_of(n.b);
                if (a >= 0 && b >= 0) r = new_cexpr(tb_cexpr{CE_BIN, n.op, a, b, 0.f});
                break;
            }
            case TB_APPEND: {  // (Some(f), Some(g)) if f == g — decidable here only for literals
                int a = const_of(n.a), b = const_of(n.b);
                if (a >= 0 && b >= 0 && out.cexpr[a].kind == CE_LIT && out.cexpr[b].kind == CE_LIT &&
                    out.cexpr[a].value == out.cexpr[b].value)
                    r = a;
                break;
            }
            case TB_MARKED:
                r = const_of(n.a);
                if (r >= 0) {  // tb_substitute may give this constant another value later
                    if (cexpr_marked.size() <= (size_t)r) cexpr_marked.resize((size_t)r + 1, 0);
                    cexpr_marked[r] = 1;
                }
                break;
            default: break;
        }
        const_memo[i] = r;
        return r;
    }
    int literal_cexpr(float v) { return new_cexpr(tb_cexpr{CE_LIT, 0, 0, 0, v}); }
    std::vector<char> cexpr_marked;  // cval index -> the constant of a Marked node
    bool is_marked_cexpr(int i) const { return (size_t)i < cexpr_marked.size() && cexpr_marked[i] != 0; }

    int new_aux(uint32_t kind, int a, int b, uint32_t words) {
        for (const tb_aux& e : out.aux)  // a node emitted twice (Filter pre-read, Fin) shares its constants
            if (e.kind == kind && e.a == a && e.b == b) return (int)e.off;
        tb_aux x{kind, a, b, out.aux_words};
        out.aux.push_back(x);
        out.aux_words += (words + 1u) & ~1u;  // blocks stay 16-byte aligned (double2 loads)
        return (int)x.off;
    }

    int state_of(int i, int words) {
        if (state_off[i] < 0) {
            state_off[i] = (int)out.state_words;
            state_len[i] = words;
            out.state_words += words;
        }
        return state_off[i];
    }
    // State blocks `waveform::set_state(w, Initial)` clears (waveform.rs:322-392): every node of the subtree
    // except the coefficient waveforms of a Filter (appendix A10: those iterators are dropped unconsumed).
    void collect_state(int i, std::vector<std::pair<int, int>>& r) const {
        const tb_node& n = nodes[i];
        if (state_off[i] >= 0 && n.kind != TB_NOISE) r.emplace_back(state_off[i], state_len[i]);  // a Noise has no tree state
        switch (n.kind) {
            case TB_FIN: case TB_APPEND: case TB_SINE: case TB_RESET: case TB_BINARY:
                collect_state(n.a, r);
                collect_state(n.b, r);
                break;
            case TB_ALT:
                collect_state(n.a, r);
                collect_state(n.b, r);
                collect_state(n.c, r);
                break;
            case TB_MARKED: case TB_CAPTURED: case TB_FILTER: collect_state(n.a, r); break;
            default: break;
        }
    }

    // ---- which sines may use the f32 polynomial ------------------------------------------
    void mark_sensitive(int i, bool s) {
        const tb_node& n = nodes[i];
        if (s) sensitive[i] = 1;
        switch (n.kind) {
            case TB_FIN: mark_sensitive(n.a, true); mark_sensitive(n.b, s); break;
            case TB_APPEND:
            case TB_BINARY: mark_sensitive(n.a, s); mark_sensitive(n.b, s); break;
            case TB_SINE: mark_sensitive(n.a, true); mark_sensitive(n.b, true); break;
            case TB_FILTER:
                mark_sensitive(n.a, s);
                for (uint32_t j = 0; j < n.ff_count + n.fb_count; j++) mark_sensitive(lists[n.list_off + j], true);
                break;
            case TB_RESET: mark_sensitive(n.a, true); mark_sensitive(n.b, s); break;
            case TB_ALT: mark_sensitive(n.a, true); mark_sensitive(n.b, s); mark_sensitive(n.c, s); break;
            case TB_MARKED:
            case TB_CAPTURED: mark_sensitive(n.a, s); break;
            default: break;
        }
    }
    uint32_t sine_flags(int i) const {
        return (fast_sines && !sensitive[i]) ? (TB_SINE_FAST << 8) : (TB_SINE_EXACT << 8);
    }

    // ---- greater_or_equals_at chain (generator.rs:787-862) --------------------------------
    int build_goe(int i) {
        tb_goe g{};
        g.term = GOE_MAYBE;
        g.step_off = (uint32_t)out.goe_steps.size();
        int cur = i;
        for (;;) {
            const tb_node& n = nodes[cur];
            int c = const_of(cur);
            if (c >= 0) {  // :796-802
                g.term = GOE_CONST;
                g.term_arg = c;
                break;
            }
            if (n.kind == TB_TIME) {
                g.term = GOE_TIME;
                g.term_arg = state_of(cur, 2);
                break;
            }
            if (n.kind == TB_APPEND) {  // looks at `a` only; None degrades to Maybe
                g.through_append = 1;
                cur = n.a;
                continue;
            }
            if (n.kind == TB_BINARY && (n.op == TB_ADD || n.op == TB_SUBTRACT)) {
                const bool ac = nodes[n.a].kind == TB_CONST, bc = nodes[n.b].kind == TB_CONST;
                if (n.op == TB_ADD && ac) {  // value - va, recurse into b
                    out.goe_steps.push_back(-1);
                    out.goe_steps.push_back(const_of(n.a));
                    g.n_steps++;
                    cur = n.b;
                    continue;
                }
                if (n.op == TB_ADD && bc) {
                    out.goe_steps.push_back(-1);
                    out.goe_steps.push_back(const_of(n.b));
                    g.n_steps++;
                    cur = n.a;
                    continue;
                }
                if (n.op == TB_SUBTRACT && bc) {  // value + vb
                    out.goe_steps.push_back(+1);
                    out.goe_steps.push_back(const_of(n.b));
                    g.n_steps++;
                    cur = n.a;
                    continue;
                }
            }
            g.term = GOE_MAYBE;  // :853, :857-860 (Marked, Multiply, ...)
            break;
        }
        out.goe.push_back(g);
        return (int)out.goe.size() - 1;
    }

    // ---- Filter tables -----------------------------------------------------------------------
    int filter_table(int i) {
        if (filt_idx[i] >= 0) return filt_idx[i];
        const tb_node& n = nodes[i];
        const uint32_t K = n.ff_count, J = n.fb_count;
        if (K > TB_MAX_K_GEN) fail(TB_ERR_UNSUPPORTED, "Filter with more than " + std::to_string(TB_MAX_K_GEN) + " feed-forward taps");
        if (K > TB_MAX_K)  // the long form reads its taps as constants (moving_average(n), std.tuun:114-115)
            for (uint32_t j = 0; j < K; j++)
                if (const_of(lists[n.list_off + j]) < 0)
                    fail(TB_ERR_UNSUPPORTED, "Filter with more than " + std::to_string(TB_MAX_K) + " feed-forward taps that are waveforms");
        if (J > TB_MAX_J_GEN) fail(TB_ERR_UNSUPPORTED, "Filter with more than " + std::to_string(TB_MAX_J_GEN) + " feedback taps");
        tb_filter_tab t{};
        t.K = K;
        t.J = J;
        t.all_const = 1;
        t.fb_const = 1;
        t.x_slot = t.u_slot = -1;
        for (uint32_t j = 0; j < K + J; j++) {
            int c = lists[n.list_off + j];
            if (nodes[c].kind != TB_CONST) t.all_const = 0;
            int k = const_of(c);
            t.coef[j] = k >= 0 ? TB_OPERAND_CONST(k) : 0;  // slots are filled in by emit_gen
            if (j >= K && k < 0) t.fb_const = 0;
        }
        if (J > TB_MAX_J) t.fb_const = 0;  // past the scan's matrix size: the serial recurrence
        out.filt.push_back(t);
        filt_idx[i] = (int)out.filt.size() - 1;
        if (t.fb_const && J > 0) {
            out.filt.back().pow_aux = out.aux_words;
            new_aux(AUX_FILT_POW, 0, filt_idx[i], 6 * J * J);
        }
        return filt_idx[i];
    }
    int filter_state(int i) {
        const tb_node& n = nodes[i];
        const int st = state_of(i, 2 + (n.ff_count - 1) + n.fb_count + 2);
        if (filt_idx[i] >= 0) out.filt[filt_idx[i]].state_off = (uint32_t)st;
        return st;
    }

    int fixed_table(int i) {
        if (fixed_idx[i] < 0) {
            out.fixed.push_back(tb_fixed_tab{nodes[i].fixed_off, nodes[i].fixed_len});
            fixed_idx[i] = (int)out.fixed.size() - 1;
        }
        return fixed_idx[i];
    }

    // ---- generate ------------------------------------------------------------------------------
    void emit_gen(int i) {
        const tb_node& n = nodes[i];
        switch (n.kind) {
            case TB_CONST: produced(emit(G_CONST, const_of(i))); break;
            case TB_TIME: produced(emit(G_TIME, state_of(i, 2))); break;
            case TB_FIXED: produced(emit(G_FIXED, state_of(i, 2), fixed_table(i))); break;
            case TB_NOISE: produced(emit(G_NOISE, state_of(i, 2), noise_id(i))); break;
            case TB_MARKED:
            case TB_CAPTURED: emit_gen(n.a); break;
            case TB_BINARY: {
                const int merge = n.op == TB_MERGE;
                const int cb = const_of(n.b);
                emit_gen(n.a);
                if (cb >= 0) {
                    if (merge) emit(G_BINC, (int)n.op, cb, merge);
                    else emit_binc((int)n.op, cb);
                } else {
                    const int s = alloc_slot();
                    const int b0 = emit(G_BIN_BEGIN, s, merge, 0);
                    emit_gen(n.b);
                    out.code[b0].c = emit(G_BIN_END, s, (int)n.op, merge);
                    produced(out.code[b0].c);
                    free_slot();
                }
                break;
            }
            case TB_SINE: {
                const int cf = const_of(n.a), cp = const_of(n.b);
                const int st = state_of(i, 2);
                const uint32_t fl = sine_flags(i);
                const int aux_inc = cf >= 0 ? new_aux(AUX_SINE_INC, cf, 0, 1) : 0;
                const int aux_ph = cp >= 0 ? new_aux(AUX_SINE_PHASE, cp, 0, 1) : 0;
                if (cf >= 0 && cp >= 0) {
                    produced(emit(G_SINE_CC | fl, st, aux_inc, aux_ph));
                } else if (cp >= 0) {
                    emit_gen(n.a);
                    produced(emit(G_SINE_AC | fl, st, 0, aux_ph));
                } else if (cf >= 0) {
                    emit_gen(n.b);
                    produced(emit(G_SINE_CA | fl, st, aux_inc, 0));
                } else {
                    emit_gen(n.a);
                    const int s = alloc_slot();
                    const int b0 = emit(G_SINE_BEGIN, s, 0, 0);
                    emit_gen(n.b);
                    out.code[b0].c = emit(G_SINE_END | fl, st, s, 0);
                    produced(out.code[b0].c);
                    free_slot();
                }
                break;
            }
            case TB_ALT: {
                const int cp = const_of(n.b), cn = const_of(n.c);
                emit_gen(n.a);
                if (cp >= 0 && cn >= 0) {
                    produced(emit(G_ALT_CC, cp, cn));
                } else {
                    const int st = alloc_slot();
                    const int b0 = emit(G_ALT_BEGIN, st, 0, 0);
                    int opp = TB_OPERAND_CONST(cp), opn = TB_OPERAND_CONST(cn);
                    int extra = 0;
                    if (cp < 0) {
                        emit_gen(n.b);
                        if (cn < 0) {
                            opp = alloc_slot();
                            extra = 1;
                            emit(G_ALT_POS, opp);
                        } else {
                            // the negative branch is constant: keep the positive one in a slot too
                            opp = alloc_slot();
                            extra = 1;
                            emit(G_ALT_POS, opp);
                        }
                    }
                    if (cn < 0) {
                        emit_gen(n.c);
                        opn = 0;
                    }
                    out.code[b0].c = emit(G_ALT_END, st, opp, opn);
                    produced(out.code[b0].c);
                    if (extra) free_slot();
                    free_slot();
                }
                break;
            }
            case TB_FILTER: {
                const int fi = filter_table(i);
                const int st = filter_state(i);
                const int K = (int)n.ff_count, J = (int)n.fb_count;
                const int p0 = emit(G_FILT_PRE, st, K, 0);
                emit_gen(n.a);
                emit(G_FILT_PRE_END, st, K, J);
                out.code[p0].c = label();
                emit_gen(n.a);
                bool any_code = false;
                for (int j = 0; j < K + J; j++) any_code |= const_of(lists[n.list_off + j]) < 0;
                const bool need_u = (J > 0 && !out.filt[fi].fb_const) || K > TB_MAX_K;
                if (!any_code) {
                    if (need_u) out.filt[fi].u_slot = alloc_slot();  // long FIR: the input tile, staged for taps 9..
                    produced(emit(G_FILT_RUN, st, fi, 0));
                    if (need_u) free_slot();
                } else {
                    int used = 0;
                    const int xs = alloc_slot();
                    used++;
                    out.filt[fi].x_slot = xs;
                    const int b0 = emit(G_FILT_BEGIN, st, fi, 0);
                    for (int j = 0; j < K + J; j++) {
                        const int c = lists[n.list_off + j];
                        if (const_of(c) >= 0) continue;
                        const int s = alloc_slot();
                        used++;
                        out.filt[fi].coef[j] = s;
                        emit_gen(c);
                        emit(G_FILT_COEF, s);
                    }
                    if (need_u) {
                        out.filt[fi].u_slot = alloc_slot();
                        used++;
                    }
                    out.code[b0].c = emit(G_FILT_RUN, st, fi, 1);
                    produced(out.code[b0].c);
                    while (used--) free_slot();
                }
                break;
            }
            case TB_FIN: {
                const int gi = build_goe(n.a);
                const tb_goe g = out.goe[gi];
                const bool may_maybe = g.term == GOE_MAYBE || g.through_append;
                const bool may_static = g.term != GOE_MAYBE;
                const int h0 = emit(G_FIN_HEAD, gi, 0, 0);
                int scan = -1;
                if (may_maybe) {
                    emit_gen(n.a);
                    scan = emit(G_FIN_SCAN, 0, 0, 0);
                }
                out.code[h0].c = label();
                if (may_static) {
                    emit_len(n.a);
                    emit(G_FIN_STATIC);
                }
                if (scan >= 0) out.code[scan].c = label();
                const int in0 = emit(G_FIN_INNER, 0, 0, 0);
                emit_gen(n.b);
                const int adv = emit(G_FIN_ADV, has_unreached_filter(n.b, false) ? 0 : 1, 0, 0);
                out.code[in0].c = adv;
                emit_len(n.b);
                emit(G_FIN_END);
                out.code[adv].c = label();
                break;
            }
            case TB_APPEND: {
                const int st = state_of(i, 1);
                const int s = alloc_slot();
                const int b0 = emit(G_APP_BEGIN, st, 0, 0);
                emit_gen(n.a);
                const int m0 = emit(G_APP_MID, st, s, 0);
                out.code[b0].c = m0;
                emit_gen(n.b);
                emit(G_APP_END, 0, s, 0);
                out.code[m0].c = label();
                free_slot();
                break;
            }
            case TB_RESET: {
                const int st = state_of(i, 2);
                emit_gen(n.a);
                // All runs of a tile at once (the segmented form) when the inner tree allows it; else run by run,
                // the way the reference does it (generator.rs:288-316): the inner tree's generate code over the
                // window of each run, its state cleared where a run starts.
                const Lowered out0 = out;
                const std::vector<int> memo0 = const_memo, off0 = state_off, len0 = state_len, fixed0 = fixed_idx,
                                       filt0 = filt_idx;
                const int slots0 = slots_in_use, label0 = label_at, prod0 = last_producer, gated0 = seg_gated;
                bool segmented = true;
                try {
                    const int s = alloc_slot();
                    const int b0 = emit(G_RESET_BEGIN, st, s, 0);
                    emit_seg(n.b);
                    out.code[b0].c = emit(G_RESET_END, st);
                    produced(out.code[b0].c);
                    free_slot();
                } catch (int status) {
                    if (status != TB_ERR_UNSUPPORTED) throw;
                    segmented = false;
                    out = out0;
                    const_memo = memo0, state_off = off0, state_len = len0, fixed_idx = fixed0, filt_idx = filt0;
                    slots_in_use = slots0, label_at = label0, last_producer = prod0, seg_gated = gated0;
                }
                if (!segmented) {
                    const int so = alloc_slot(), sr = alloc_slot();
                    const int b0 = emit(G_RUNS_BEGIN, st, so, 0);
                    const int body = label();
                    emit_gen(n.b);
                    std::vector<std::pair<int, int>> ranges, merged;
                    collect_state(n.b, ranges);
                    std::sort(ranges.begin(), ranges.end());
                    for (const auto& r : ranges) {
                        if (!merged.empty() && merged.back().first + merged.back().second == r.first) merged.back().second += r.second;
                        else merged.push_back(r);
                    }
                    if (merged.empty()) merged.emplace_back(0, 0);
                    out.code[b0].c = emit(G_RUNS_END, so, sr, body);
                    for (size_t k = 0; k < merged.size(); k++)  // data words, never executed
                        emit(G_RUNS_END, merged[k].first, merged[k].second, k + 1 == merged.size());
                    produced(-1);
                    (void)label();
                    free_slot();
                    free_slot();
                }
                break;
            }
            default: fail(TB_ERR_INVALID, "unknown node kind");
        }
    }

    // ---- length ----------------------------------------------------------------------------------
    // Does the subtree hold a Filter that a non-empty `generate` of the subtree may not have reached — one in the
    // second part of an Append, inside a nested Fin or Reset, in a length or coefficient waveform?  `length(w, 0)`
    // turns such a Filter from Initial into zero history without its look-ahead (generator.rs:690-704), so the
    // empty advance at the end of a Fin (:166) must still run for it.  Any other node is left as it is by
    // `length(w, 0)`.
    bool has_unreached_filter(int i, bool under) const {
        const tb_node& n = nodes[i];
        switch (n.kind) {
            case TB_FILTER:
                if (under) return true;
                for (uint32_t j = 0; j < n.ff_count + n.fb_count; j++)
                    if (has_unreached_filter(lists[n.list_off + j], true)) return true;
                return has_unreached_filter(n.a, under);
            case TB_APPEND: return has_unreached_filter(n.a, under) || has_unreached_filter(n.b, true);
            case TB_FIN: return has_unreached_filter(n.a, true) || has_unreached_filter(n.b, true);
            case TB_RESET: return has_unreached_filter(n.a, under) || has_unreached_filter(n.b, true);
            case TB_BINARY: case TB_SINE: return has_unreached_filter(n.a, under) || has_unreached_filter(n.b, under);
            case TB_ALT:
                return has_unreached_filter(n.a, under) || has_unreached_filter(n.b, under) || has_unreached_filter(n.c, under);
            case TB_MARKED: case TB_CAPTURED: return has_unreached_filter(n.a, under);
            default: return false;
        }
    }
    // Subtrees whose `length` (generator.rs:620-782) touches no state and returns `max`: constants, noise and
    // what is made of them (a Sine's length does not move its accumulator).  One L_INF stands for the whole
    // subtree, and a point operator / Sine / Alt with such an operand needs no PUSH .. MIN around the other.
    bool len_trivial(int i) const {
        const tb_node& n = nodes[i];
        switch (n.kind) {
            case TB_CONST: case TB_NOISE: return true;
            case TB_MARKED: case TB_CAPTURED: case TB_RESET: return len_trivial(n.a);
            case TB_BINARY: case TB_SINE: return len_trivial(n.a) && len_trivial(n.b);
            case TB_ALT: return len_trivial(n.a) && len_trivial(n.b) && len_trivial(n.c);
            default: return false;
        }
    }
    void emit_len(int i) {
        const tb_node& n = nodes[i];
        if (len_trivial(i)) {
            emit(L_INF);
            return;
        }
        switch (n.kind) {
            case TB_CONST:
            case TB_NOISE: emit(L_INF); break;
            case TB_TIME: emit(L_TIME, state_of(i, 2)); break;
            case TB_FIXED: emit(L_FIXED, state_of(i, 2), fixed_table(i)); break;
            case TB_MARKED:
            case TB_CAPTURED: emit_len(n.a); break;
            case TB_SINE:
                (void)state_of(i, 2);
                if (len_trivial(n.b)) { emit_len(n.a); break; }
                if (len_trivial(n.a)) { emit_len(n.b); break; }
                emit_len(n.a);
                emit(L_PUSH);
                emit_len(n.b);
                emit(L_MIN);
                break;
            case TB_BINARY:
                if (len_trivial(n.a) || len_trivial(n.b)) {  // min(x, max) = x; Merge: max(x, max) = max
                    emit_len(len_trivial(n.a) ? n.b : n.a);
                    if (n.op == TB_MERGE) emit(L_INF);
                    break;
                }
                emit_len(n.a);
                emit(L_PUSH);
                emit_len(n.b);
                emit(n.op == TB_MERGE ? L_MAX : L_MIN);
                break;
            case TB_RESET:
                (void)state_of(i, 2);
                emit_len(n.a);
                break;
            case TB_ALT:
                emit_len(n.a);
                if (len_trivial(n.b) && len_trivial(n.c)) break;  // nothing to advance in the branches
                emit(L_PUSH);
                emit_len(n.b);
                emit_len(n.c);
                emit(L_POP);
                break;
            case TB_FILTER: {
                (void)filter_table(i);
                const int st = filter_state(i);
                emit(L_FILT_HEAD, st, (int)n.ff_count, (int)n.fb_count);
                emit_len(n.a);
                const int m0 = emit(L_FILT_MID, 0, 0, 0);
                for (uint32_t j = 0; j < n.ff_count + n.fb_count; j++) emit_len(lists[n.list_off + j]);
                out.code[m0].c = emit(L_FILT_END);
                break;
            }
            case TB_APPEND: {
                const int st = state_of(i, 1);
                const int b0 = emit(L_APP_BEGIN, st, 0, 0);
                emit_len(n.a);
                out.code[b0].c = emit(L_APP_MID, st);
                emit_len(n.b);
                emit(L_APP_END);
                break;
            }
            case TB_FIN: {
                const int gi = build_goe(n.a);
                const tb_goe g = out.goe[gi];
                const bool may_maybe = g.term == GOE_MAYBE || g.through_append;
                const bool may_static = g.term != GOE_MAYBE;
                const int h0 = emit(G_FIN_HEAD, gi, 0, 0);
                int scan2 = -1;
                if (may_maybe) {
                    out.pure_len = 0;
                    const int s = alloc_slot();
                    emit(G_SAVE, s);
                    emit_gen(n.a);
                    emit(L_FIN_SCAN1);
                    emit(G_RESTORE, s);
                    free_slot();
                    emit_len(n.b);
                    scan2 = emit(L_FIN_SCAN2, 0, 0, 0);
                }
                out.code[h0].c = label();
                if (may_static) {
                    emit_len(n.b);
                    emit(L_PUSH);
                    emit_len(n.a);
                    emit(L_FIN_STATIC);
                }
                if (scan2 >= 0) out.code[scan2].c = label();
                break;
            }
            default: fail(TB_ERR_INVALID, "unknown node kind");
        }
    }

    // Control-stack words the interpreter (render.cu: PUSH / POP) may hold while it is inside node i — the
    // largest over the generate, length and segmented forms of each kind, so a bound, not a count.
    int ctl_need(int i) const {
        const tb_node& n = nodes[i];
        switch (n.kind) {
            case TB_MARKED: case TB_CAPTURED: return ctl_need(n.a);
            case TB_BINARY: case TB_SINE: return std::max(ctl_need(n.a), 1 + ctl_need(n.b));
            case TB_ALT: return std::max(ctl_need(n.a), 1 + std::max(ctl_need(n.b), ctl_need(n.c)));
            case TB_FIN: return 3 + std::max(ctl_need(n.a), ctl_need(n.b));
            case TB_APPEND: return std::max(1 + ctl_need(n.a), 2 + ctl_need(n.b));
            case TB_RESET: return std::max(ctl_need(n.a), 4 + ctl_need(n.b));
            case TB_FILTER: {
                int m = ctl_need(n.a);
                for (uint32_t j = 0; j < n.ff_count + n.fb_count; j++) m = std::max(m, ctl_need(lists[n.list_off + j]));
                return 3 + m;
            }
            default: return 0;
        }
    }

    // A waveform that never returns short: constants, clocks, noise and what is made of them.
    bool never_ends(int i) const {
        const tb_node& n = nodes[i];
        switch (n.kind) {
            case TB_CONST: case TB_TIME: case TB_NOISE: return true;
            case TB_MARKED: case TB_CAPTURED: return never_ends(n.a);
            case TB_SINE: return never_ends(n.a) && never_ends(n.b);
            case TB_BINARY: return n.op == TB_MERGE ? (never_ends(n.a) || never_ends(n.b)) : (never_ends(n.a) && never_ends(n.b));
            case TB_RESET: case TB_ALT: return never_ends(n.a);  // as long as the trigger (generator.rs:273-343)
            default: return false;
        }
    }

    // ---- segmented (inside a Reset) ----------------------------------------------------------------
    // seg_gated > 0 while lowering an operand the reference asks for fewer samples than the run has left
    // (the inner of a Fin, the parts of an Append, the right-hand side of a point operator / the phase of a
    // Sine / the branches of an Alt / the inner of a nested Reset whose first operand can end).  A Noise there
    // draws a data-dependent number of samples per run (generator.rs:113-118 draws exactly what it is asked
    // for), and S_NOISE counts whole tiles: not taken.
    int seg_gated = 0;
    void emit_seg_gated(int i, bool gated) {
        seg_gated += gated;
        emit_seg(i);
        seg_gated -= gated;
    }
    void emit_seg(int i) {
        const tb_node& n = nodes[i];
        switch (n.kind) {
            case TB_CONST: emit(S_CONST, const_of(i)); break;
            case TB_TIME: emit(S_TIME, state_of(i, 2)); break;
            case TB_FIXED: emit(S_FIXED, state_of(i, 2), fixed_table(i)); break;
            case TB_MARKED:
            case TB_CAPTURED: emit_seg(n.a); break;
            case TB_BINARY: {
                const int merge = n.op == TB_MERGE;
                const int cb = const_of(n.b);
                emit_seg(n.a);
                if (cb >= 0) {
                    emit(S_BINC, (int)n.op, cb, merge);
                } else {
                    const int s = alloc_slot();
                    emit(S_BIN_BEGIN, s);
                    emit_seg_gated(n.b, !merge && !never_ends(n.a));  // generator.rs:538-549: b renders a_len samples
                    emit(S_BIN_END, s, (int)n.op, merge);
                    free_slot();
                }
                break;
            }
            case TB_SINE: {
                const int cf = const_of(n.a), cp = const_of(n.b);
                const int st = state_of(i, 2);
                const uint32_t fl = sine_flags(i);
                const int aux_inc = cf >= 0 ? new_aux(AUX_SINE_INC, cf, 0, 1) : 0;
                const int aux_ph = cp >= 0 ? new_aux(AUX_SINE_PHASE, cp, 0, 1) : 0;
                if (cf >= 0 && cp >= 0) {
                    emit(S_SINE_CC | fl, st, aux_inc, aux_ph);
                } else if (cp >= 0) {
                    emit_seg(n.a);
                    emit(S_SINE_AC | fl, st, 0, aux_ph);
                } else if (cf >= 0) {
                    emit_seg(n.b);
                    emit(S_SINE_CA | fl, st, aux_inc, 0);
                } else {
                    emit_seg(n.a);
                    const int s = alloc_slot();
                    emit(S_SINE_BEGIN, s);
                    emit_seg_gated(n.b, !never_ends(n.a));  // generator.rs:206: the phase renders f_len samples
                    emit(S_SINE_END | fl, st, s, 0);
                    free_slot();
                }
                break;
            }
            case TB_ALT: {
                const int cp = const_of(n.b), cn = const_of(n.c);
                emit_seg(n.a);
                if (cp >= 0 && cn >= 0) {
                    emit(S_ALT_CC, cp, cn);
                } else {
                    const int st = alloc_slot();
                    emit(S_ALT_BEGIN, st);
                    int opp = TB_OPERAND_CONST(cp), opn = TB_OPERAND_CONST(cn);
                    int extra = 0;
                    const bool gated = !never_ends(n.a);  // generator.rs:320-343: both branches render t_len samples
                    if (cp < 0) {
                        emit_seg_gated(n.b, gated);
                        opp = alloc_slot();
                        extra = 1;
                        emit(S_ALT_POS, opp);
                    }
                    if (cn < 0) {
                        emit_seg_gated(n.c, gated);
                        opn = 0;
                    }
                    emit(S_ALT_END, st, opp, opn);
                    if (extra) free_slot();
                    free_slot();
                }
                break;
            }
            case TB_RESET: {
                const int st = state_of(i, 2);
                emit_seg(n.a);
                const int s = alloc_slot();
                emit(S_RESET_BEGIN, st, s);
                emit_seg_gated(n.b, !never_ends(n.a));
                emit(S_RESET_END, st);
                free_slot();
                break;
            }
            case TB_FIN: {
                const int gi = build_goe(n.a);
                const tb_goe g = out.goe[gi];
                if (g.term == GOE_MAYBE || g.through_append)
                    fail(TB_ERR_UNSUPPORTED, "Fin with a rendered (non-analytic) length inside a Reset");
                emit_seg_gated(n.b, true);  // generator.rs:164: the inner renders `len` samples
                emit(S_FIN, gi);
                break;
            }
            case TB_NOISE:
                if (seg_gated > 0)
                    fail(TB_ERR_UNSUPPORTED, "Noise under a Fin, an Append or a finite operand inside a Reset");
                emit(S_NOISE, state_of(i, 2), noise_id(i));
                break;
            case TB_FILTER: fail(TB_ERR_UNSUPPORTED, "Filter inside a Reset");
            case TB_APPEND: {
                // Append under a Reset (a retriggered envelope: `Fin(..) ++ Fin(..) ++ ..`).  Every run restarts
                // the Append; inside a run the second part starts, from its Initial state, where the first
                // one ends (generator.rs:169-188).  Evaluated for all runs of a tile at once when that point is
                // known without rendering: the first part is a Fin whose length is analytic in the run's own
                // Time (greater_or_equals_at, :787-862) over a waveform that cannot end earlier.
                int a = n.a;
                while (nodes[a].kind == TB_MARKED || nodes[a].kind == TB_CAPTURED) a = nodes[a].a;
                if (nodes[a].kind != TB_FIN)
                    fail(TB_ERR_UNSUPPORTED, "Append inside a Reset whose first part is not a Fin");
                const int gi = build_goe(nodes[a].a);
                const tb_goe g = out.goe[gi];
                if (g.term != GOE_TIME || g.through_append || !never_ends(nodes[a].b))
                    fail(TB_ERR_UNSUPPORTED,
                         "Append inside a Reset: the first part needs an analytic length over an infinite waveform");
                const int sa = alloc_slot(), so = alloc_slot();
                emit(S_APP_BEGIN, gi);
                emit_seg(n.a);
                emit(S_APP_MID, sa, so, gi);
                const uint32_t st_begin = out.state_words;
                emit_seg_gated(n.b, true);  // generator.rs:186: the second part renders what the first left
                const uint32_t st_count = out.state_words - st_begin;
                if (st_begin >= 0x10000u || st_count >= 0x8000u) fail(TB_ERR_UNSUPPORTED, "Append inside a Reset: too much state");
                emit(S_APP_END, sa, so, (int)(st_begin | (st_count << 16)));
                free_slot();
                free_slot();
                break;
            }
            default: fail(TB_ERR_INVALID, "unknown node kind");
        }
    }

    // ---- steady-state stream (steady.cuh) ---------------------------------------------------------
    // A second, straight-line rendering of the generate walk for trees made only of infinite,
    // window-free nodes (constants, clocks, sines, point operators, Alt, constant-coefficient
    // filters).  Returns false when the tree holds anything else; the caller then discards it.
    // Point operators with a constant right-hand side (generator.rs:538-549) become post-op words
    // of the instruction that produced the accumulator; `* c` followed by `+ c'` or `- c'` shares
    // one word (both operations still round separately, like the reference's two nodes).
    int s_slots = 0;
    int s_last = -1;  // last instruction that left a fresh accumulator
    int one_idx = -1, negzero_idx = -1;
    int s_alloc() {
        int s = s_slots++;
        out.n_slots = std::max<uint32_t>(out.n_slots, (uint32_t)s_slots);
        return s;
    }
    void s_produced(int idx) { s_last = idx; }
    void s_postop(uint32_t op, int cidx) {
        if (one_idx < 0) {
            one_idx = literal_cexpr(1.0f);
            negzero_idx = literal_cexpr(-0.0f);
        }
        tb_insn& prod = out.code[s_last];
        const int npost = (int)((prod.op >> 16) & 0xffu);
        if (op == TB_MERGE) op = TB_ADD;  // infinite operands: Merge is Add (generator.rs:263)
        if (op == TB_ADD || op == TB_SUBTRACT) {
            const int addend = op == TB_ADD ? cidx : new_cexpr(tb_cexpr{CE_NEG, 0, cidx, 0, 0.f});
            if (npost > 0) {
                tb_insn& prev = out.code[s_last + npost];
                if ((prev.op & 0xffu) == ST_AFFINE && prev.c == negzero_idx) {  // completes a multiply
                    prev.c = addend;
                    return;
                }
            }
            out.code[s_last].op += 1u << 16;
            emit(ST_AFFINE, 0, one_idx, addend);
        } else if (op == TB_MULTIPLY) {
            out.code[s_last].op += 1u << 16;
            emit(ST_AFFINE, 0, cidx, negzero_idx);
        } else {
            out.code[s_last].op += 1u << 16;
            emit(ST_OPC, (int)op, cidx, 0);
        }
    }
    // A Reset in the steady stream (lane kernels only): the trigger is rendered, ST_RESET_CLK turns it into a
    // per-sample local clock (samples since the last restart), and the inner tree is emitted with clk_slot set —
    // which admits only nodes that are closed-form in that clock: constants, Time, sines with constant rate
    // and phase, point operators, Alt, noise (not restarted by a Reset).  That covers sawtooth, pulse and
    // triangle of lib/v0/std.tuun (Appendix B of SURVEY.md); anything else keeps the general interpreter.
    int clk_slot = -1;
    bool lane_steady_root = false;  // the steady stream renders the whole (infinite) tree
    bool steady_nested_clk = false; // a Reset inside a Reset: no time-axis split
    // greater_or_equals_at (generator.rs:787-862) of a chain of literals against a clock: the sample the Fin ends at
    // (the arithmetic of the kernels' goe_eval); false when a step is a per-voice parameter.
    bool literal_target(const tb_goe& g, uint64_t* target) const {
        float value = 0.0f;
        for (uint32_t k = 0; k < g.n_steps; k++) {
            const tb_cexpr& e = out.cexpr[out.goe_steps[g.step_off + 2 * k + 1]];
            if (e.kind != CE_LIT) return false;
            value = out.goe_steps[g.step_off + 2 * k] > 0 ? value + e.value : value - e.value;
        }
        const float t = std::ceil(value * (float)sample_rate);
        if (!(t >= 1.0f) || t >= 2.0e9f) return false;
        *target = (uint64_t)t;
        return true;
    }
    // A timeline (program.h ST_SEG_*): Append(Fin{T - c0, e0}, Append(Fin{T - c1, e1}, .. last)) where `last` is one more
    // such Fin or a tree that never ends, under a root Fin that ends no later than the timeline does (so neither the
    // timeline's own end nor the zero-extension of a Merge around it is ever rendered: generator.rs:169-188, :520-570).
    // Every e_k must be closed-form in the clock of its piece, like the inner tree of a Reset.
    uint64_t steady_limit = 0;  // samples the root Fin lets through when that is a literal, else 0
    int tl_root_time = -1;      // state offset of the root Fin's Time
    bool in_piece = false;
    bool have_timeline = false;
    struct TlPiece {
        uint32_t start, next;                        // first sample of the piece and of the one behind it
        int fin_time, app;                           // state offsets or -1
        std::vector<std::pair<int, int>> ranges;     // its state ranges
    };
    std::vector<TlPiece> tl_pieces;
    bool emit_timeline(int i) {
        if (clk_slot >= 0 || in_piece || have_timeline || steady_limit == 0 || sample_rate == 0) return false;
        struct Piece { int tree, fin, fin_time, app; uint64_t start; };
        std::vector<Piece> pieces;
        uint64_t start = 0;
        for (int cur = i;;) {
            const tb_node& n = nodes[cur];
            const bool app = n.kind == TB_APPEND;
            const int f = app ? n.a : cur;
            if (nodes[f].kind == TB_FIN) {
                const tb_goe g = out.goe[build_goe(nodes[f].a)];
                uint64_t len = 0;
                if (g.term != GOE_TIME || g.through_append || !literal_target(g, &len)) return false;
                pieces.push_back(Piece{nodes[f].b, f, g.term_arg, app ? state_of(cur, 1) : -1, start});
                start += len;
            } else if (!app && never_ends(cur)) {
                pieces.push_back(Piece{cur, -1, -1, -1, start});
                start = ~0ull;
            } else {
                return false;
            }
            if (!app) break;
            cur = n.b;
            if (pieces.size() > 64) return false;
        }
        if (start < steady_limit) return false;  // the timeline would end inside the note
        while (pieces.back().start >= steady_limit) pieces.pop_back();  // never begun
        const int cs = s_alloc(), vs = s_alloc();
        if (vs > 0xff) return false;
        have_timeline = true;
        out.lane_clk = 1;
        for (size_t k = 0; k < pieces.size(); k++) {
            const Piece& pc = pieces[k];
            const bool last = k + 1 == pieces.size();
            emit(ST_SEG_CLK, 0, cs, (int)pc.start);
            // its second word (never run: ST_SEG_CLK takes it): a = where the piece ends, b = words from here to its
            // ST_SEG_SEL (build_lane_plan) — a tile that lies wholly outside the piece jumps there
            emit(ST_SEG_CLK | 0x200u, last ? 0x7fffffff : (int)pieces[k + 1].start, 0, 0);
            s_last = -1;
            in_piece = true;
            clk_slot = cs;
            const bool ok = emit_steady(pc.tree);
            clk_slot = -1;
            in_piece = false;
            if (!ok) return false;
            if (k == 0 && last) return false;  // (an Append has two parts)
            s_produced(emit(ST_SEG_SEL | (last ? 0x100u : 0u), 0, vs, (int)pc.start));
            if (!last) emit(ST_SAVE, vs);
            TlPiece t;
            t.start = (uint32_t)pc.start;
            t.next = last ? 0xffffffffu : (uint32_t)pieces[k + 1].start;
            t.fin_time = pc.fin_time;
            t.app = pc.app;
            collect_state(pc.fin >= 0 ? pc.fin : pc.tree, t.ranges);
            tl_pieces.push_back(t);
        }
        s_slots -= 2;
        return true;
    }
    bool emit_steady(int i) {
        const tb_node& n = nodes[i];
        switch (n.kind) {
            case TB_CONST: s_produced(emit(ST_CONST, const_of(i))); return true;
            case TB_TIME:
                if (clk_slot >= 0) s_produced(emit(ST_TIME_CLK, state_of(i, 2), clk_slot));
                else s_produced(emit(ST_TIME, state_of(i, 2)));
                return true;
            case TB_RESET: {
                // Nested (hard sync: a pulse restarted by another oscillator): the trigger is rendered against the
                // enclosing clock, and the inner clock restarts with it as well — set_state(inner, Initial) of the
                // enclosing Reset (generator.rs:311-313) makes this one "previously negative" again (:276-279).
                if (in_piece) return false;
                const int outer = clk_slot;
                if (!emit_steady(n.a)) return false;
                const int s = s_alloc();
                if (s > 0xff) return false;
                emit(ST_RESET_CLK, state_of(i, 2), s, outer);
                s_last = -1;
                clk_slot = s;
                const bool ok = emit_steady(n.b);
                clk_slot = outer;
                s_slots--;
                out.lane_clk = 1;
                if (outer >= 0) steady_nested_clk = true;
                return ok;
            }
            case TB_APPEND: return emit_timeline(i);
            case TB_NOISE:
                // (in a piece of a timeline its draw count would have to start with the piece: not closed-form in a clock)
                if (in_piece) return false;
                s_produced(emit(ST_NOISE, state_of(i, 2), noise_id(i)));
                return true;
            case TB_MARKED:
            case TB_CAPTURED: return emit_steady(n.a);
            case TB_BINARY: {
                const int cb = const_of(n.b);
                if (!emit_steady(n.a)) return false;
                if (cb >= 0) {
                    if (s_last < 0 || ((out.code[s_last].op >> 16) & 0xffu) >= 15) return false;
                    s_postop(n.op, cb);
                } else {
                    const int s = s_alloc();
                    emit(ST_SAVE, s);
                    if (!emit_steady(n.b)) return false;
                    s_produced(emit(ST_BIN, s, n.op == TB_MERGE ? (int)TB_ADD : (int)n.op));
                    s_slots--;
                }
                return true;
            }
            case TB_SINE: {
                const int cf = const_of(n.a), cp = const_of(n.b);
                const int st = state_of(i, 2);
                const uint32_t fl = sine_flags(i);
                const int aux_inc = cf >= 0 ? new_aux(AUX_SINE_INC, cf, 0, 1) : 0;
                const int aux_ph = cp >= 0 ? new_aux(AUX_SINE_PHASE, cp, 0, 1) : 0;
                if (clk_slot >= 0) {
                    if (cf < 0 || cp < 0) return false;  // a phase sum would have to restart with the clock
                    s_produced(emit(ST_SINE_CLK | fl | ((uint32_t)clk_slot << 24), st, aux_inc, aux_ph));
                    return true;
                }
                if (cf >= 0 && cp >= 0) {
                    s_produced(emit(ST_SINE_CC, st, new_aux(AUX_SINE_ROT, cf, 0, 2 + 2 * TB_CS), aux_ph));
                } else if (cp >= 0) {
                    if (!emit_steady(n.a)) return false;
                    s_produced(emit(ST_SINE_AC | fl, st, 0, aux_ph));
                } else if (cf >= 0) {
                    if (!emit_steady(n.b)) return false;
                    s_produced(emit(ST_SINE_CA | fl, st, aux_inc, 0));
                } else {
                    if (!emit_steady(n.a)) return false;
                    const int s = s_alloc();
                    emit(ST_SAVE, s);
                    if (!emit_steady(n.b)) return false;
                    s_produced(emit(ST_SINE_AA | fl, st, s, 0));
                    s_slots--;
                }
                return true;
            }
            case TB_ALT: {
                const int cp = const_of(n.b), cn = const_of(n.c);
                if (!emit_steady(n.a)) return false;
                if (cp >= 0 && cn >= 0) {
                    s_produced(emit(ST_ALT_CC, cp, cn));
                    return true;
                }
                const int st = s_alloc();
                emit(ST_SAVE, st);
                int opp = TB_OPERAND_CONST(cp), opn = TB_OPERAND_CONST(cn);
                int extra = 0;
                if (cp < 0) {
                    if (!emit_steady(n.b)) return false;
                    opp = s_alloc();
                    extra = 1;
                    emit(ST_SAVE, opp);
                }
                if (cn < 0) {
                    if (!emit_steady(n.c)) return false;
                    opn = 0;
                }
                s_produced(emit(ST_ALT, st, opp, opn));
                s_slots -= 1 + extra;
                return true;
            }
            case TB_FILTER: {
                if (clk_slot >= 0) return false;  // a filter's history restarts with its Reset
                if (n.ff_count > TB_MAX_K || n.fb_count > TB_MAX_J) return false;  // long filters: general interpreter only
                const int fi = filter_table(i);
                const int st = filter_state(i);
                const uint32_t K = n.ff_count, J = n.fb_count;
                for (uint32_t j = 0; j < K + J; j++)
                    if (const_of(lists[n.list_off + j]) < 0) return false;  // coefficient waveforms
                if (!emit_steady(n.a)) return false;
                const int coef = new_aux(AUX_FILT_COEF, 0, fi, (K + J + 1) / 2);
                s_produced(emit(ST_FILT | (K << 8) | (J << 12), st, coef, J > 0 ? (int)out.filt[fi].pow_aux : 0));
                return true;
            }
            default: return false;
        }
    }


    // ---- time-axis split plan (program.h tb_split_entry) ---------------------------------------------
    // Walks the tree the steady stream was emitted from.  Returns the first render pass in which node i's
    // output is right in every segment (1 = the first pass), recording one entry per stateful node.
    int split_clk_level = -1;  // >= 0 inside a Reset: the pass in which its trigger is right
    int split_clk_reset = -1;  // ... and the state block of that Reset
    int split_level(int i) {
        const tb_node& n = nodes[i];
        switch (n.kind) {
            case TB_CONST: return 1;
            case TB_TIME:
                if (split_clk_level >= 0) {  // the run's own clock
                    out.split.push_back(tb_split_entry{SP_CLK, (uint32_t)state_off[i], (uint32_t)split_clk_level, -1, split_clk_reset});
                    return split_clk_level + 1;
                }
                out.split.push_back(tb_split_entry{SP_POS, (uint32_t)state_off[i], 0u, 0, 0});
                return 1;
            case TB_NOISE:  // not restarted by a Reset
                out.split.push_back(tb_split_entry{SP_POS, (uint32_t)state_off[i], 0u, 0, 0});
                return 1;
            case TB_RESET: {
                const int lt = split_level(n.a);
                out.split.push_back(tb_split_entry{SP_RESET_SIGN, (uint32_t)state_off[i], (uint32_t)lt, 0, 0});
                split_clk_level = lt;
                split_clk_reset = state_off[i];
                const int li = split_level(n.b);
                split_clk_level = -1;
                return std::max(lt + 1, li);
            }
            case TB_MARKED:
            case TB_CAPTURED: return split_level(n.a);
            case TB_BINARY: {
                const int la = split_level(n.a);
                return const_of(n.b) >= 0 ? la : std::max(la, split_level(n.b));
            }
            case TB_ALT: {
                int lv = split_level(n.a);
                if (const_of(n.b) < 0) lv = std::max(lv, split_level(n.b));
                if (const_of(n.c) < 0) lv = std::max(lv, split_level(n.c));
                return lv;
            }
            case TB_SINE: {
                const int cf = const_of(n.a), cp = const_of(n.b);
                int lv = 1;
                if (split_clk_level >= 0) {  // constant rate and phase (emit_steady): accumulator = rate x local clock
                    out.split.push_back(tb_split_entry{SP_CLK, (uint32_t)state_off[i], (uint32_t)split_clk_level, cf, split_clk_reset});
                    return split_clk_level + 1;
                }
                if (cf >= 0) {
                    out.split.push_back(tb_split_entry{SP_SINE_CONST, (uint32_t)state_off[i], 0u, cf, 0});
                } else {
                    const int lf = split_level(n.a);  // the increments are right in pass lf: the sum after it
                    out.split.push_back(tb_split_entry{SP_SINE_VAR, (uint32_t)state_off[i], (uint32_t)lf, 0, 0});
                    lv = lf + 1;
                }
                if (cp < 0) lv = std::max(lv, split_level(n.b));
                return lv;
            }
            case TB_FILTER: {
                const int li = split_level(n.a);
                out.split.push_back(tb_split_entry{SP_FILTER, (uint32_t)state_off[i], (uint32_t)li, filt_idx[i], 0});
                return li + 1;
            }
            default: return 1 << 20;  // not reached: emit_steady took the tree
        }
    }

    // ---- lane-per-voice plan (lanes.cuh) -------------------------------------------------------------
    // The ST_* stream again, with every operand rewritten for a thread that keeps its own voice in
    // its own column of shared memory (program.h): constants at W[0, n_cval), the state block at
    // W[n_cval, n_cval + state_words), derived constants behind it.  Slots keep their indices.
    int lane_slots_max = 0;
    const tb_aux* aux_at(int off) const {
        for (const tb_aux& e : out.aux)
            if ((int)e.off == off) return &e;
        return nullptr;
    }
    int lane_new_aux(uint32_t kind, int a, uint32_t w_words, uint32_t q_units, uint32_t* q_off, int b = 0, int c = 0) {
        for (const tb_lane_aux& e : out.lane_aux)
            if (e.kind == kind && e.a == a && e.b == b && e.c == c) {
                if (q_off) *q_off = e.q_off;
                return (int)e.w_off;
            }
        tb_lane_aux x{kind, a, b, c, out.lane_w_words, out.lane_q_units};
        if (kind == LA_ROT) {
            // The rotation table depends on the frequency alone: sines of one frequency (the same constant, the
            // same literal, the same parameter column) share it; each keeps its own increment and carried pair.
            // (Shared memory per voice decides how many voices an SM holds: 144 bytes a table.)
            auto same = [&](int i, int j) {
                if (i == j) return true;
                const tb_cexpr &p = out.cexpr[i], &q = out.cexpr[j];
                if (p.kind == CE_LIT && q.kind == CE_LIT)  // (not a literal tb_substitute may replace)
                    return !is_marked_cexpr(i) && !is_marked_cexpr(j) && std::memcmp(&p.value, &q.value, sizeof(float)) == 0;
                return p.kind == CE_PARAM && q.kind == CE_PARAM && p.a == q.a;
            };
            for (const tb_lane_aux& e : out.lane_aux)
                if (e.kind == LA_ROT && same(e.a, a)) {
                    x.q_off = e.q_off;
                    q_units = 0;
                    break;
                }
        }
        out.lane_aux.push_back(x);
        out.lane_w_words += w_words;
        out.lane_q_units += q_units;
        if (q_off) *q_off = x.q_off;
        return (int)x.w_off;
    }
    bool build_lane_plan() {
        const int n_cval = (int)out.cexpr.size();
        const int st0 = n_cval;
        out.lane_w_words = (uint32_t)n_cval + out.state_words;
        out.lane_q_units = 0;
        out.lane_code.clear();
        out.lane_aux.clear();
        int max_slot = -1;
        auto slot = [&](int s) { max_slot = std::max(max_slot, s); return s; };
        for (size_t pc = out.pc_steady; pc < out.code.size(); pc++) {
            tb_insn in = out.code[pc];
            const uint32_t op = in.op & 0xffu;
            switch (op) {
                case ST_END: case ST_CONST: case ST_ALT_CC: case ST_AFFINE: case ST_OPC: break;
                case ST_TIME: case ST_NOISE: in.a += st0; break;
                case ST_TIME_CLK: in.a += st0; slot(in.b); break;
                case ST_RESET_CLK:
                    in.a += st0;
                    slot(in.b);
                    if (in.c >= 0) slot(in.c);
                    break;
                case ST_SEG_CLK: case ST_SEG_SEL:
                    if (in.op & 0x200u) break;  // second word of ST_SEG_CLK
                    if (tl_root_time < 0) return false;
                    in.a = lane_new_aux(LA_TL_POS, st0 + tl_root_time, 2, 0, nullptr);
                    slot(in.b);
                    break;
                case ST_SINE_CLK: {
                    const tb_aux* inc = aux_at(in.b);
                    const tb_aux* ph = aux_at(in.c);
                    if (!inc || !ph) return false;
                    in.a += st0;
                    in.b = lane_new_aux(LA_INC, inc->a, 2, 0, nullptr);
                    in.c = lane_new_aux(LA_PHASE, ph->a, 2, 0, nullptr);
                    slot((int)(in.op >> 24));
                    break;
                }
                case ST_SAVE: case ST_BIN: slot(in.a); break;
                case ST_SINE_CC: {
                    const tb_aux* rot = aux_at(in.b);
                    const tb_aux* ph = aux_at(in.c);
                    if (!rot || !ph) return false;
                    uint32_t q = 0;
                    in.a += st0;
                    in.b = lane_new_aux(LA_ROT, rot->a, 6, TB_LS / 2 + 1, &q, in.a, ph->a);
                    in.c = 0;
                    if (q > 0xffu) return false;
                    in.op = (in.op & ~0xff00u) | (q << 8);
                    break;
                }
                case ST_SINE_AC: {
                    const tb_aux* ph = aux_at(in.c);
                    if (!ph) return false;
                    in.a += st0;
                    in.c = lane_new_aux(LA_PHASE, ph->a, 2, 0, nullptr);
                    break;
                }
                case ST_SINE_CA: {
                    const tb_aux* inc = aux_at(in.b);
                    if (!inc) return false;
                    in.a += st0;
                    in.b = lane_new_aux(LA_INC, inc->a, 2, 0, nullptr);
                    break;
                }
                case ST_SINE_AA: in.a += st0; slot(in.b); break;
                case ST_ALT:
                    slot(in.a);
                    if (in.b >= 0) slot(in.b);
                    break;
                case ST_FILT: {
                    const tb_aux* cf = aux_at(in.b);
                    if (!cf) return false;
                    const uint32_t K = (in.op >> 8) & 0xfu, J = (in.op >> 12) & 0x7u;
                    in.a += st0;
                    in.b = lane_new_aux(LA_COEF, cf->b, K + J, 0, nullptr);
                    in.c = 0;
                    break;
                }
                default: return false;
            }
            out.lane_code.push_back(in);
            if (op == ST_END) break;
        }
        out.lane_slots = (uint32_t)(max_slot + 1);
        if (!tl_pieces.empty()) {  // what finish_lane leaves for the general interpreter (lanes.cuh)
            const uint32_t pos_w = (uint32_t)lane_new_aux(LA_TL_POS, st0 + tl_root_time, 2, 0, nullptr);
            for (const TlPiece& t : tl_pieces) {
                out.lane_aux.push_back(tb_lane_aux{LA_TL_PIECE, (int32_t)t.start, t.fin_time >= 0 ? st0 + t.fin_time : -1,
                                                   t.app >= 0 ? st0 + t.app : -1, pos_w, t.next});
                for (const auto& r : t.ranges)
                    if (r.second > 0)
                        out.lane_aux.push_back(tb_lane_aux{LA_TL_ZERO, (int32_t)t.start, st0 + r.first, r.second, pos_w, 0u});
            }
        }
        const char* fe = std::getenv("TUUN_B200_LANE_FUSE");  // diagnostics: "0" keeps the plain ST_* words
        if (!(fe && fe[0] == '0')) {
            fuse_lane_fm();
            fold_lane_postops();
        }
        // ST_SEG_CLK: distance to the piece's ST_SEG_SEL, in words of the program as it now stands
        auto width = [&](size_t i) {
            const uint32_t op = out.lane_code[i].op & 0xffu;
            return (size_t)1 + (op == ST_END ? 0u : ((out.lane_code[i].op >> 16) & 0xffu)) + (op == LN_FM ? 1u : 0u);
        };
        for (size_t i = 0; i < out.lane_code.size(); i += width(i)) {
            if ((out.lane_code[i].op & 0xffu) != ST_SEG_CLK || (out.lane_code[i].op & 0x200u)) continue;
            size_t j = i + 2;
            while (j < out.lane_code.size() && (out.lane_code[j].op & 0xffu) != ST_SEG_SEL) j += width(j);
            if (j >= out.lane_code.size()) return false;
            // Not the last piece: past its ST_SEG_SEL (and the ST_SAVE folded behind it) as well — a tile in front of
            // the piece keeps the running result (the pieces before it, already saved), a tile behind it keeps nothing.
            // The last piece's ST_SEG_SEL runs every tile: it advances the position.
            if (!(out.lane_code[j].op & 0x100u)) j += width(j);
            out.lane_code[i + 1].b = (int32_t)(j - (i + 2));
        }
        return true;
    }
    // In the lane program an instruction that only transforms the running result in place — Alt with two
    // constant branches (square waves), a biquad with constant coefficients — rides as one more post-op
    // word of the instruction that produced the result: one dispatch and one trip through shared memory
    // less per tile each (`square(f) | lpf(..)` becomes a single instruction with two post-op words).
    void fold_lane_postops() {
        std::vector<tb_insn> in = out.lane_code, res;
        long last = -1;  // index in res of the last instruction that may take post-ops
        for (size_t i = 0; i < in.size();) {
            const uint32_t op = in[i].op & 0xffu;
            const size_t words = (op == LN_FM ? 2 : 1);
            const uint32_t np = op == ST_END ? 0 : ((in[i].op >> 16) & 0xffu);
            // (ST_SAVE — the running result into a slot — and ST_BIN — slot (op) running result — ride along too: what
            //  they need is in registers where the producer ends, and every word that is not a dispatch saves one)
            const bool foldable = op == ST_ALT_CC || (in[i].op & 0xffffu) == (ST_FILT | (3u << 8) | (2u << 12)) ||
                                  op == ST_SAVE || op == ST_BIN;
            if (foldable && last >= 0 && ((res[last].op >> 16) & 0xffu) + 1 + np <= 0xffu) {
                res[last].op += (1u + np) << 16;
                tb_insn w = in[i];
                w.op &= 0xffffu;  // a post-op word carries no count of its own
                res.push_back(w);
                for (uint32_t k = 0; k < np; k++) res.push_back(in[i + 1 + k]);
            } else {
                last = (op == ST_SAVE || op == ST_RESET_CLK || op == ST_SEG_CLK || op == ST_END) ? -1 : (long)res.size();
                for (size_t k = 0; k < words + np && i + k < in.size(); k++) res.push_back(in[i + k]);
            }
            i += words + np;
        }
        out.lane_code = res;
    }

    // Peephole over the lane program: a constant-rate sine, scaled and offset, driving the frequency of
    // a sine with constant phase (vibrato, FM: `$(c + m * $f)`) or the phase of a sine with constant
    // frequency (PM: `sine(f, m * $g)`), optionally straight into a biquad (every filter of
    // lib/v0/std.tuun), becomes one LN_FM instruction (program.h).
    void fuse_lane_fm() {
        std::vector<tb_insn> in = out.lane_code, res;
        for (size_t i = 0; i < in.size();) {
            const tb_insn& cc = in[i];
            const bool cand = (cc.op & 0xffu) == ST_SINE_CC && (cc.op >> 16) == 1 && i + 2 < in.size() &&
                              (in[i + 1].op & 0xffu) == ST_AFFINE &&
                              ((in[i + 2].op & 0xffu) == ST_SINE_AC || (in[i + 2].op & 0xffu) == ST_SINE_CA);
            if (!cand) {
                // copy the instruction with its post-op words
                const size_t n = 1 + (((cc.op & 0xffu) == ST_END) ? 0 : ((cc.op >> 16) & 0xffu));
                for (size_t k = 0; k < n && i + k < in.size(); k++) res.push_back(in[i + k]);
                i += n;
                continue;
            }
            const tb_insn& aff = in[i + 1];
            const tb_insn& ac = in[i + 2];
            uint32_t np = ac.op >> 16;
            size_t next = i + 3;  // first post-op word of the carrier, if any
            tb_insn w0{}, w1{};
            const bool pm = (ac.op & 0xffu) == ST_SINE_CA;  // the scaled sine drives the phase, not the frequency
            w0.op = LN_FM | (cc.op & 0xff00u) | (((ac.op >> 8) & 0xffu) << 24) | (pm ? TB_LN_FM_PHASE << 24 : 0u);
            w0.a = cc.b + 2;
            w0.b = ac.a;
            w0.c = pm ? ac.b : ac.c;
            w1.op = 0;
            w1.a = aff.b;
            w1.b = aff.c;
            w1.c = -1;
            if (np == 0 && next < in.size() && (in[next].op & 0xffffu) == (ST_FILT | (3u << 8) | (2u << 12))) {
                w1.c = in[next].a;
                w1.op = (uint32_t)in[next].b;
                np = in[next].op >> 16;
                next++;
            }
            w0.op |= np << 16;
            res.push_back(w0);
            res.push_back(w1);
            for (uint32_t k = 0; k < np; k++) res.push_back(in[next + k]);
            i = next + np;
        }
        out.lane_code = res;
    }

    void run() {
        validate();
        const_memo.assign(n_nodes, -2);
        state_off.assign(n_nodes, -1);
        state_len.assign(n_nodes, 0);
        fixed_idx.assign(n_nodes, -1);
        filt_idx.assign(n_nodes, -1);
        sensitive.assign(n_nodes, 0);
        const int root = (int)n_nodes - 1;
        if (ctl_need(root) > TB_CTL_DEPTH)
            fail(TB_ERR_UNSUPPORTED, "tree nested too deeply: needs " + std::to_string(ctl_need(root)) +
                                         " control-stack words, the kernel has " + std::to_string(TB_CTL_DEPTH));
        mark_sensitive(root, false);
        out.pure_len = 1;
        out.pc_gen = (uint32_t)here();
        emit_gen(root);
        emit(OP_END);
        out.pc_len = (uint32_t)here();
        emit_len(root);
        emit(OP_END);
        {   // the steady-state stream, when the whole tree qualifies
            const size_t code0 = out.code.size(), cexpr0 = out.cexpr.size(), aux0 = out.aux.size();
            const uint32_t aux_words0 = out.aux_words, slots0 = out.n_slots;
            out.pc_steady = (uint32_t)here();
            // A root Fin whose length is found analytically (generator.rs:787-862: Time, constants, +- constants)
            // over a steady tree — `$440 * Qw`, any note with a fixed duration — is rendered by the lane kernel
            // as its inner tree; the lane kernel only has to count how much of it belongs to the voice
            // (tb_launch::lane_fin_goe).  The warp-per-voice steady interpreter does not take such trees.
            int sroot = root;
            while (nodes[sroot].kind == TB_MARKED || nodes[sroot].kind == TB_CAPTURED) sroot = nodes[sroot].a;
            bool fin_ok = false;
            if (nodes[sroot].kind == TB_FIN) {
                const size_t goe0 = out.goe.size(), steps0 = out.goe_steps.size();
                const int gi = build_goe(nodes[sroot].a);
                const tb_goe g = out.goe[gi];
                uint64_t target = 0;
                steady_limit = (g.term == GOE_TIME && !g.through_append && sample_rate != 0 && literal_target(g, &target)) ? target : 0;
                tl_root_time = g.term == GOE_TIME ? g.term_arg : -1;
                if ((g.term == GOE_TIME || g.term == GOE_CONST) && !g.through_append && emit_steady(nodes[sroot].b)) {
                    emit(ST_END);
                    out.lane_fin_goe = gi;
                    fin_ok = true;
                } else {
                    out.goe.resize(goe0);
                    out.goe_steps.resize(steps0);
                    tl_pieces.clear();
                    have_timeline = false;
                }
                steady_limit = 0;
            }
            if (fin_ok) {
                out.steady_ok = 0;
            } else if (nodes[sroot].kind != TB_FIN && emit_steady(root)) {
                emit(ST_END);
                out.steady_ok = out.lane_clk ? 0 : 1;  // clocked words: lane kernels only
                lane_steady_root = true;
                {   // clocked words run on the lane kernels only: lane_ok decides below whether the plan stands
                    const int passes = split_level(root);
                    bool ok = passes <= 8 && !steady_nested_clk;
                    for (const tb_split_entry& e : out.split) ok = ok && (int)e.state_off >= 0;
                    if (ok) out.split_passes = (uint32_t)passes;
                    else out.split.clear();
                }
            } else {
                out.lane_clk = 0;
                tl_pieces.clear();
                have_timeline = false;
                out.code.resize(code0);
                out.cexpr.resize(cexpr0);
                for (int& m : const_memo)
                    if (m >= (int)cexpr0) m = -2;
                one_idx = negzero_idx = -1;
                out.aux.resize(aux0);
                out.aux_words = aux_words0;
                out.n_slots = slots0;
                out.pc_steady = 0;
                out.steady_ok = 0;
            }
        }
        if (out.state_words == 0) out.state_words = 1;
        if (out.cexpr.empty()) literal_cexpr(0.f);
        if (out.aux_words == 0) out.aux_words = 1;
        if (out.n_slots == 0) out.n_slots = 1;
        out.lane_ok = ((out.steady_ok || out.lane_fin_goe >= 0 || lane_steady_root) && build_lane_plan()) ? 1u : 0u;
        if (!out.lane_ok) {
            out.lane_fin_goe = -1;
            if (out.lane_clk) {  // a split of clocked words needs the lane kernels
                out.split.clear();
                out.split_passes = 0;
            }
            out.lane_clk = 0;
        }
        out.n_nodes = n_nodes;
    }
};

}  // namespace

// The parts of a root sequence (lower.h).
bool sequence_parts(const tb_node* nodes, uint32_t n_nodes, const int32_t* lists, uint32_t n_lists, uint64_t pool_len,
                    uint32_t sample_rate, std::vector<SeqPart>& parts) {
    parts.clear();
    Lowered tmp;
    Lowerer L(nodes, n_nodes, lists, n_lists, pool_len, tmp, true);
    try {
        L.validate();
        L.const_memo.assign(n_nodes, -2);
        L.state_off.assign(n_nodes, -1);
        L.state_len.assign(n_nodes, 0);
        L.fixed_idx.assign(n_nodes, -1);
        L.filt_idx.assign(n_nodes, -1);
        L.sensitive.assign(n_nodes, 0);
        // the leaves of the root's tree of Appends, in order (Marked / Captured are pass-throughs, generator.rs:344-358)
        std::vector<int> leaves;
        std::function<void(int)> collect = [&](int i) {
            int c = i;
            while (nodes[c].kind == TB_MARKED || nodes[c].kind == TB_CAPTURED) c = nodes[c].a;
            if (nodes[c].kind == TB_APPEND) {
                collect(nodes[c].a);
                collect(nodes[c].b);
            } else {
                leaves.push_back(i);  // with its wrappers
            }
        };
        collect((int)n_nodes - 1);
        if (leaves.size() < 2 || leaves.size() > 4096) return false;
        int cur = leaves.back();
        for (size_t q = 0; q + 1 < leaves.size(); q++) {
            int a = leaves[q];
            while (nodes[a].kind == TB_MARKED || nodes[a].kind == TB_CAPTURED) a = nodes[a].a;
            if (nodes[a].kind != TB_FIN || !L.never_ends(nodes[a].b)) return false;
            const int gi = L.build_goe(nodes[a].a);
            const tb_goe g = tmp.goe[gi];
            if (g.term != GOE_TIME || g.through_append) return false;
            float value = 0.0f;  // greater_or_equals_at (generator.rs:787-862), the arithmetic of the kernels' goe_eval
            for (uint32_t k = 0; k < g.n_steps; k++) {
                const tb_cexpr& e = tmp.cexpr[tmp.goe_steps[g.step_off + 2 * k + 1]];
                if (e.kind != CE_LIT) return false;  // a per-voice length: the parts would start at per-voice offsets
                value = tmp.goe_steps[g.step_off + 2 * k] > 0 ? value + e.value : value - e.value;
            }
            const float t = std::ceil(value * (float)sample_rate);
            if (!(t >= 1.0f) || t >= 9.0e15f) return false;
            parts.push_back(SeqPart{leaves[q], (uint64_t)t});
        }
        if (parts.empty()) return false;
        parts.push_back(SeqPart{cur, ~0ull});
        return true;
    } catch (int) {
        parts.clear();
        return false;
    }
}

int lower(const tb_node* nodes, uint32_t n_nodes, const int32_t* lists, uint32_t n_lists,
          uint64_t pool_len, bool fast_sines, Lowered& out, const uint32_t* noise_ids, uint32_t sample_rate) {
    out = Lowered();
    Lowerer L(nodes, n_nodes, lists, n_lists, pool_len, out, fast_sines);
    L.noise_ids = noise_ids;
    L.sample_rate = sample_rate;
    try {
        L.run();
    } catch (int status) {
        return status;
    } catch (const std::bad_alloc&) {
        out.status = TB_ERR_NOMEM;
        out.error = "out of host memory while lowering";
        return TB_ERR_NOMEM;
    }
    return TB_OK;
}

}  // namespace tb
