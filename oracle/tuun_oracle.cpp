// tuun_oracle.cpp — CPU oracle (TEST INFRASTRUCTURE ONLY; see tuun_oracle.h).
//
// Scalar restatement of /root/reference/src/lib/generator.rs:86-862 and
// src/lib/waveform.rs:179-392 over the tb_node op list.  Each function cites the lines it
// follows.  Build: g++ -O2 -ffp-contract=off -fno-fast-math (oracle/Makefile) so every f32
// operation rounds exactly once, as rustc emits it.
#include "tuun_oracle.h"

#include <atomic>
#include <cmath>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <thread>
#include <vector>

namespace {

// generator.rs:12-35
enum class St { Initial, Position, Finished, Samples, Phase, Sign };

struct Node {
    uint32_t kind = 0, op = 0;
    Node *a = nullptr, *b = nullptr, *c = nullptr;
    float value = 0.f;
    int32_t param_slot = -1;
    uint32_t mark_id = 0;
    std::vector<Node*> ff, fb;
    const float* fixed = nullptr;
    size_t fixed_len = 0;
    // State
    St st = St::Initial;
    size_t position = 0;
    bool finished = false;
    std::deque<float> input, output;
    double accumulator = 0.0;
    float signum = 0.f;
    // Noise: index of this node in the op list and the number of samples drawn so far
    uint32_t index = 0;
    uint64_t noise_drawn = 0;
};

enum class MO { Some, None, Maybe };  // generator.rs:58-63
struct MaybeOption {
    MO tag;
    size_t v;
};

// fastrand 2.3.0 (Cargo.lock:372-373; not vendored under /root/reference) — its published
// generator is wyrand: state += C0; t = state * (state ^ C1) as u128; out = lo(t) ^ hi(t);
// f32() = from_bits(0x3F800000 | (u32 >> 9)) - 1.0 in [0, 1).  The reference calls the UNSEEDED
// thread-local instance (generator.rs:115), so no sequence of it is reproducible: parity unpinned.
// Here every Noise node of every voice owns a stream of that generator (state advances by a
// constant, so sample k is a pure function of the stream's seed and k):
//     state_k = seed + NODE_K * (node_index + 1) + VOICE_K * voice + C0 * (k + 1)
struct Rng {
    static constexpr uint64_t C0 = 0x2d358dccaa6c78a5ull, C1 = 0x8bb84b93962eacc9ull;
    static constexpr uint64_t NODE_K = 0x9e3779b97f4a7c15ull, VOICE_K = 0xd6e8feb86659fd93ull;
    uint64_t seed = 0x7475756E2545F491ull;
    uint64_t voice = 0;
    static float f32_at(uint64_t state) {
        const unsigned __int128 t = (unsigned __int128)state * (unsigned __int128)(state ^ C1);
        const uint64_t r = (uint64_t)t ^ (uint64_t)(t >> 64);
        const uint32_t bits = 0x3F800000u | ((uint32_t)r >> 9);
        float f;
        memcpy(&f, &bits, 4);
        return f - 1.0f;
    }
    float sample(uint32_t node_index, uint64_t k) const {
        return f32_at(seed + NODE_K * (uint64_t)(node_index + 1) + VOICE_K * voice + C0 * (k + 1));
    }
};

// Rust `f as usize` (saturating, NaN -> 0).
inline size_t f32_as_usize(float f) {
    if (!(f > 0.f)) return 0;
    if (f >= 18446744073709551616.0f) return std::numeric_limits<size_t>::max();
    return (size_t)f;
}
// Rust f32::signum
inline float rust_signum(float x) {
    if (std::isnan(x)) return x;
    return std::signbit(x) ? -1.0f : 1.0f;
}
// Rust f64::rem_euclid
inline double rem_euclid(double x, double rhs) {
    double r = std::fmod(x, rhs);
    return r < 0.0 ? r + std::fabs(rhs) : r;
}
const double TAU = 6.283185307179586476925286766559;

struct Gen {
    uint32_t sample_rate;
    size_t allocations = 0;
    Rng* rng;
    // The reference reads scratch buffers past the length their producer returned (generator.rs:555-567,
    // 206-220, 320-343); what sits there is whatever the producer's in-place rendering left behind ("tail
    // undefined", :76-95) on top of the vec![0.0] initialisation.  clean_tails zeroes those tails instead —
    // the documented semantics ("missing samples are zeros", overview.md:147-160) — so that tests can tell a
    // tree that depends on such leftovers from a real difference.
    bool clean_tails = false;

    static float apply(uint32_t op, float a, float b) {  // generator.rs:262-270
        switch (op) {
            case TB_ADD:
            case TB_MERGE: return a + b;
            case TB_SUBTRACT: return a - b;
            case TB_MULTIPLY: return a * b;
            case TB_DIVIDE: return b == 0.0f ? 0.0f : a / b;
            default: return powf(a, b);
        }
    }

    // generator.rs:86-380
    size_t generate(Node* w, float* out, size_t n) {
        if (n == 0) return 0;  // :93
        switch (w->kind) {
            case TB_CONST:  // :97
                for (size_t i = 0; i < n; i++) out[i] = w->value;
                return n;
            case TB_TIME:  // :101-111
                if (w->st == St::Initial) {
                    w->st = St::Position;
                    w->position = 0;
                }
                for (size_t i = 0; i < n; i++)
                    out[i] = (float)(w->position + i) / (float)sample_rate;
                w->position += n;
                return n;
            case TB_NOISE:  // :113-118
                for (size_t i = 0; i < n; i++) out[i] = rng->sample(w->index, w->noise_drawn + i) * 2.0f - 1.0f;
                w->noise_drawn += n;
                return n;
            case TB_FIXED: {  // :119-131
                if (w->st == St::Initial) {
                    w->st = St::Position;
                    w->position = 0;
                }
                if (w->position >= w->fixed_len) return 0;
                size_t len = std::min(w->fixed_len - w->position, n);
                memcpy(out, w->fixed + w->position, len * sizeof(float));
                w->position += len;
                return len;
            }
            case TB_FIN: {  // :133-168
                Node dummy;
                dummy.kind = TB_CONST;
                dummy.value = 0.0f;
                Node* inner = w->b;
                w->b = &dummy;
                size_t len = length(w, n);
                w->b = inner;
                size_t inner_len = generate(inner, out, len);
                (void)length(inner, n - len);
                return inner_len;
            }
            case TB_APPEND: {  // :169-188
                if (w->st == St::Initial) {
                    w->st = St::Finished;
                    w->finished = false;
                }
                size_t a_len;
                if (!w->finished) {
                    a_len = generate(w->a, out, n);
                    if (a_len == n) return a_len;
                    w->finished = true;
                } else {
                    a_len = 0;
                }
                size_t b_len = generate(w->b, out + a_len, n - a_len);
                return a_len + b_len;
            }
            case TB_SINE: {  // :191-221
                if (w->st == St::Initial) {
                    w->st = St::Phase;
                    w->accumulator = 0.0;
                }
                size_t f_len = generate(w->a, out, n);
                std::vector<float> ph_out(f_len, 0.0f);
                allocations += f_len;
                size_t ph_len = generate(w->b, ph_out.data(), f_len);
                if (clean_tails) for (size_t i = ph_len; i < f_len; i++) ph_out[i] = 0.0f;
                for (size_t i = 0; i < f_len; i++) {
                    float sample = (float)std::sin(w->accumulator + (double)ph_out[i]);
                    double f = (double)out[i];
                    double phase_inc = f / (double)sample_rate;
                    out[i] = sample;
                    w->accumulator = rem_euclid(w->accumulator + phase_inc, TAU);
                }
                return ph_len;
            }
            case TB_FILTER: {  // :223-258
                if (w->st == St::Initial) {
                    size_t ff_count = w->ff.size();
                    std::vector<float> input(ff_count - 1, 0.0f);
                    allocations += ff_count - 1;
                    size_t inner_len = generate(w->a, input.data(), input.size());
                    input.resize(inner_len);
                    w->input.assign(input.begin(), input.end());
                    w->output.assign(w->fb.size(), 0.0f);
                    allocations += w->fb.size();
                    w->st = St::Samples;
                }
                return generate_filter(w, out, n);
            }
            case TB_BINARY:  // :260-272
                return generate_binary_op(w->op, w->a, w->b, w->op == TB_MERGE, out, n);
            case TB_RESET: {  // :273-318
                if (w->st == St::Initial) {
                    w->st = St::Sign;
                    w->signum = -1.0f;
                }
                size_t t_len = generate(w->a, out, n);
                size_t generated = 0;
                while (generated < t_len) {
                    bool reset_inner_position = false;
                    size_t inner_desired = t_len - generated;
                    for (size_t i = 0; generated + i < n; i++) {  // out[generated..] — to the block end
                        float x = out[generated + i];
                        if (w->signum < 0.0f && x >= 0.0f) {
                            inner_desired = i;
                            reset_inner_position = true;
                            w->signum = rust_signum(x);
                            break;
                        } else if (w->signum >= 0.0f && x < 0.0f) {
                            w->signum = rust_signum(x);
                        }
                    }
                    size_t inner_len = generate(w->b, out + generated, inner_desired);
                    for (size_t i = generated + inner_len; i < generated + inner_desired; i++)
                        out[i] = 0.0f;
                    if (reset_inner_position) set_state_initial(w->b);
                    generated += inner_desired;
                }
                return t_len;
            }
            case TB_ALT: {  // :320-343
                size_t t_len = generate(w->a, out, n);
                std::vector<float> pos(t_len, 0.0f), neg(t_len, 0.0f);
                allocations += 2 * t_len;
                const size_t p_len = generate(w->b, pos.data(), t_len);
                const size_t n_len = generate(w->c, neg.data(), t_len);
                if (clean_tails) {
                    for (size_t i = p_len; i < t_len; i++) pos[i] = 0.0f;
                    for (size_t i = n_len; i < t_len; i++) neg[i] = 0.0f;
                }
                for (size_t i = 0; i < t_len; i++) out[i] = out[i] >= 0.0f ? pos[i] : neg[i];
                return t_len;
            }
            case TB_MARKED:  // :344-346
            case TB_CAPTURED:  // :347-378 with capture_state == None
                return generate(w->a, out, n);
        }
        return 0;
    }

    // generator.rs:382-515
    size_t generate_filter(Node* w, float* out, size_t n) {
        std::deque<float>& input = w->input;
        std::deque<float>& output = w->output;
        size_t inner_len = generate(w->a, out, n);
        size_t out_len = std::min(n, inner_len + input.size());
        size_t extra_samples_read = n - inner_len;
        for (size_t i = inner_len; i < inner_len + extra_samples_read; i++) out[i] = 0.0f;

        size_t ff_count = w->ff.size();
        size_t input_padding = 0;
        if (input.size() != ff_count - 1) {
            // assert_eq!(0, inner_len) in the reference (:414)
            input_padding = (ff_count - 1) - input.size();
        }
        input.resize(input.size() + input_padding, 0.0f);
        size_t fb_count = w->fb.size();

        bool all_const = true;  // :428-440 — literally Const, not Marked(Const)
        for (Node* c : w->ff) all_const &= c->kind == TB_CONST;
        for (Node* c : w->fb) all_const &= c->kind == TB_CONST;
        std::vector<float> ff_coeffs(ff_count, 0.0f), fb_coeffs(fb_count, 0.0f);
        std::vector<std::vector<float>> ff_outs, fb_outs;
        if (all_const) {
            for (size_t j = 0; j < ff_count; j++) ff_coeffs[j] = w->ff[j]->value;
            for (size_t j = 0; j < fb_count; j++) fb_coeffs[j] = w->fb[j]->value;
        } else {
            for (Node* c : w->ff) {
                std::vector<float> o(out_len, 0.0f);
                allocations += out_len;
                const size_t c_len = generate(c, o.data(), out_len);
                if (clean_tails) for (size_t i = c_len; i < out_len; i++) o[i] = 0.0f;
                ff_outs.push_back(std::move(o));
            }
            for (Node* c : w->fb) {
                std::vector<float> o(out_len, 0.0f);
                allocations += out_len;
                const size_t c_len = generate(c, o.data(), out_len);
                if (clean_tails) for (size_t i = c_len; i < out_len; i++) o[i] = 0.0f;
                fb_outs.push_back(std::move(o));
            }
        }
        for (size_t i = 0; i < out_len; i++) {  // :482-508
            if (!all_const) {
                for (size_t j = 0; j < ff_count; j++) ff_coeffs[j] = ff_outs[j][i];
                for (size_t j = 0; j < fb_count; j++) fb_coeffs[j] = fb_outs[j][i];
            }
            float x = out[i];
            input.push_back(x);
            x = x * ff_coeffs[0];
            for (size_t j = 0; j + 1 < ff_count; j++) {
                float t = ff_coeffs[j + 1] * input[(ff_count - 1) - (j + 1)];
                x = x + t;
            }
            for (size_t j = 0; j < fb_count; j++) {
                float t = fb_coeffs[j] * output[(fb_count - 1) - j];
                x = x - t;
            }
            out[i] = x;
            input.pop_front();
            output.push_back(x);
            output.pop_front();
        }
        // :513 — usize subtraction; release builds wrap (Cargo.toml sets no overflow-checks),
        // and Vec/VecDeque::truncate with len > self.len() is a no-op.
        size_t drop = input_padding + extra_samples_read;
        if (drop <= input.size()) input.resize(input.size() - drop);
        return out_len;
    }

    // generator.rs:520-570
    size_t generate_binary_op(uint32_t op, Node* a, Node* b, bool extend, float* out, size_t n) {
        size_t a_len = generate(a, out, n);
        if (a_len == 0 && extend) return generate(b, out, n);
        size_t len = extend ? n : a_len;
        float f;
        if (is_const(b, &f)) {
            for (size_t i = a_len; i < len; i++) out[i] = 0.0f;
            for (size_t i = 0; i < len; i++) out[i] = apply(op, out[i], f);
            return len;
        }
        std::vector<float> b_out(len, 0.0f);
        allocations += len;
        size_t b_len = generate(b, b_out.data(), len);
        if (clean_tails) for (size_t i = b_len; i < b_out.size(); i++) b_out[i] = 0.0f;
        len = extend ? std::max(a_len, b_len) : std::min(a_len, b_len);
        for (size_t i = a_len; i < len; i++) out[i] = 0.0f;
        for (size_t i = 0; i < len; i++) out[i] = apply(op, out[i], b_out[i]);
        return len;
    }

    // generator.rs:574-612
    bool is_const(const Node* w, float* f) const {
        switch (w->kind) {
            case TB_CONST: *f = w->value; return true;
            case TB_BINARY: {
                float x, y;
                if (is_const(w->a, &x) && is_const(w->b, &y)) {
                    *f = apply(w->op, x, y);
                    return true;
                }
                return false;
            }
            case TB_APPEND: {
                float x, y;
                if (is_const(w->a, &x) && is_const(w->b, &y) && x == y) {
                    *f = x;
                    return true;
                }
                return false;
            }
            case TB_MARKED: return is_const(w->a, f);
            default: return false;
        }
    }

    // generator.rs:620-782
    size_t length(Node* w, size_t max) {
        switch (w->kind) {
            case TB_CONST: return max;
            case TB_TIME:
                if (w->st == St::Initial) {
                    w->st = St::Position;
                    w->position = 0;
                }
                w->position += max;
                return max;
            case TB_NOISE: return max;
            case TB_FIXED: {
                if (w->st == St::Initial) {
                    w->st = St::Position;
                    w->position = 0;
                }
                if (w->position >= w->fixed_len) return 0;
                size_t len = std::min(max, w->fixed_len - w->position);
                w->position += len;
                return len;
            }
            case TB_FIN: {  // :649-689
                MaybeOption m = greater_or_equals_at(w->a, 0.0f, max);
                if (m.tag == MO::Some) {
                    size_t inner_len = length(w->b, max);
                    (void)length(w->a, max);
                    return std::min(m.v, inner_len);
                } else if (m.tag == MO::None) {
                    size_t inner_len = length(w->b, max);
                    (void)length(w->a, max);
                    return inner_len;
                }
                std::vector<float> length_out(max, 0.0f);
                allocations += max;
                size_t length_len = generate(w->a, length_out.data(), max);
                size_t inner_len = length(w->b, max);
                for (size_t i = 0; i < max; i++) {
                    if (i == length_len || length_out[i] >= 0.0f || i == inner_len) return i;
                }
                return max;
            }
            case TB_FILTER: {  // :690-724
                if (w->st == St::Initial) {
                    w->input.assign(w->ff.size() - 1, 0.0f);
                    w->output.assign(w->fb.size(), 0.0f);
                    w->st = St::Samples;
                    return length(w->a, max);
                }
                size_t inner_len = length(w->a, max);
                for (Node* c : w->ff) (void)length(c, max);
                for (Node* c : w->fb) (void)length(c, max);
                return inner_len;
            }
            case TB_APPEND: {  // :725-742
                if (w->st == St::Initial) {
                    w->st = St::Finished;
                    w->finished = false;
                }
                size_t a_len;
                if (!w->finished) {
                    a_len = length(w->a, max);
                    if (a_len < max) w->finished = true;
                } else {
                    a_len = 0;
                }
                size_t b_len = length(w->b, max - a_len);
                return a_len + b_len;
            }
            case TB_SINE: {  // :743-749
                size_t f_len = length(w->a, max);
                size_t ph_len = length(w->b, max);
                return std::min(f_len, ph_len);
            }
            case TB_BINARY: {  // :750-761
                size_t a_len = length(w->a, max);
                size_t b_len = length(w->b, max);
                return w->op == TB_MERGE ? std::max(a_len, b_len) : std::min(a_len, b_len);
            }
            case TB_RESET: return length(w->a, max);  // :762-767
            case TB_ALT: {  // :768-778
                size_t len = length(w->a, max);
                (void)length(w->b, max);
                (void)length(w->c, max);
                return len;
            }
            case TB_MARKED:
            case TB_CAPTURED: return length(w->a, max);
        }
        return 0;
    }

    // generator.rs:787-862
    MaybeOption greater_or_equals_at(const Node* w, float value, size_t max) const {
        float f;
        if (is_const(w, &f)) return f >= value ? MaybeOption{MO::Some, 0} : MaybeOption{MO::None, 0};
        switch (w->kind) {
            case TB_TIME: {  // :806-817
                size_t position = w->st == St::Initial ? 0 : w->position;
                float current_value = (float)position / (float)sample_rate;
                if (current_value >= value) return {MO::Some, 0};
                size_t target = f32_as_usize(std::ceil(value * (float)sample_rate));
                return {MO::Some, std::min(max, target - position)};
            }
            case TB_APPEND: {  // :818-839
                MaybeOption m = greater_or_equals_at(w->a, value, max);
                if (m.tag == MO::None) return {MO::Maybe, 0};
                return m;
            }
            case TB_BINARY:
                if (w->op == TB_ADD || w->op == TB_SUBTRACT) {  // :840-855
                    const Node *a = w->a, *b = w->b;
                    bool ac = a->kind == TB_CONST, bc = b->kind == TB_CONST;
                    if (w->op == TB_ADD) {
                        if (ac && bc)
                            return a->value + b->value >= value ? MaybeOption{MO::Some, 0}
                                                                : MaybeOption{MO::None, 0};
                        if (ac) return greater_or_equals_at(b, value - a->value, max);
                        if (bc) return greater_or_equals_at(a, value - b->value, max);
                    } else {
                        if (ac && bc)
                            return a->value - b->value >= value ? MaybeOption{MO::Some, 0}
                                                                : MaybeOption{MO::None, 0};
                        if (bc) return greater_or_equals_at(a, value + b->value, max);
                    }
                }
                return {MO::Maybe, 0};
            default: return {MO::Maybe, 0};  // :857-860
        }
    }

    // waveform.rs:322-392 with new_state = Initial.  Filter coefficient subtrees are NOT
    // visited: the reference builds `iter_mut().map(..)` and drops it unconsumed (:357-360).
    static void set_state_initial(Node* w) {
        switch (w->kind) {
            case TB_CONST:
            case TB_NOISE: return;
            case TB_TIME:
            case TB_FIXED: w->st = St::Initial; return;
            case TB_FIN:
                set_state_initial(w->a);
                set_state_initial(w->b);
                return;
            case TB_APPEND:
            case TB_SINE:
            case TB_RESET:
                set_state_initial(w->a);
                set_state_initial(w->b);
                w->st = St::Initial;
                return;
            case TB_FILTER:
                set_state_initial(w->a);
                w->st = St::Initial;
                return;
            case TB_BINARY:
                set_state_initial(w->a);
                set_state_initial(w->b);
                return;
            case TB_ALT:
                set_state_initial(w->a);
                set_state_initial(w->b);
                set_state_initial(w->c);
                return;
            case TB_MARKED:
            case TB_CAPTURED: set_state_initial(w->a); return;
        }
    }
};

}  // namespace

struct tbo_program {
    std::vector<std::unique_ptr<Node>> nodes;
    std::vector<float> pool;
    Node* root = nullptr;
    uint32_t sample_rate = 0;
    Rng rng;
    size_t allocations = 0;
    bool clean_tails = false;
    // source for cloning
    std::vector<tb_node> src;
    std::vector<int32_t> lists;
};

static int build(tbo_program* p) {
    const size_t n = p->src.size();
    p->nodes.clear();
    for (size_t i = 0; i < n; i++) p->nodes.emplace_back(new Node());
    auto child = [&](int32_t idx, size_t self, Node** out) -> bool {
        if (idx < 0 || (size_t)idx >= self) return false;
        *out = p->nodes[idx].get();
        return true;
    };
    for (size_t i = 0; i < n; i++) {
        const tb_node& s = p->src[i];
        Node* w = p->nodes[i].get();
        w->index = (uint32_t)i;
        w->kind = s.kind;
        w->op = s.op;
        w->value = s.value;
        w->param_slot = s.param_slot;
        w->mark_id = s.mark_id;
        bool ok = true;
        switch (s.kind) {
            case TB_CONST:
            case TB_TIME:
            case TB_NOISE: break;
            case TB_FIXED:
                if (s.fixed_len > p->pool.size() || s.fixed_off > p->pool.size() - s.fixed_len) return TB_ERR_INVALID;  // no u64 wrap
                w->fixed = p->pool.data() + s.fixed_off;
                w->fixed_len = s.fixed_len;
                break;
            case TB_FIN:
            case TB_APPEND:
            case TB_SINE:
            case TB_RESET: ok = child(s.a, i, &w->a) && child(s.b, i, &w->b); break;
            case TB_BINARY:
                ok = s.op <= TB_POWER && child(s.a, i, &w->a) && child(s.b, i, &w->b);
                break;
            case TB_ALT: ok = child(s.a, i, &w->a) && child(s.b, i, &w->b) && child(s.c, i, &w->c); break;
            case TB_MARKED:
            case TB_CAPTURED: ok = child(s.a, i, &w->a); break;
            case TB_FILTER: {
                ok = child(s.a, i, &w->a) && s.ff_count >= 1 &&
                     (size_t)s.list_off + s.ff_count + s.fb_count <= p->lists.size();
                if (!ok) break;
                for (uint32_t j = 0; j < s.ff_count + s.fb_count; j++) {
                    Node* c = nullptr;
                    if (!child(p->lists[s.list_off + j], i, &c)) return TB_ERR_INVALID;
                    (j < s.ff_count ? w->ff : w->fb).push_back(c);
                }
                break;
            }
            default: ok = false;
        }
        if (!ok) return TB_ERR_INVALID;
    }
    p->root = p->nodes.back().get();
    return TB_OK;
}

extern "C" {

int tbo_program_create(const tb_node* nodes, uint32_t n_nodes, const int32_t* lists,
                       uint32_t n_lists, const float* fixed_pool, uint64_t fixed_len,
                       uint32_t sample_rate, tbo_program** out_program) {
    if (!nodes || n_nodes == 0 || !out_program || sample_rate == 0) return TB_ERR_INVALID;
    std::unique_ptr<tbo_program> p(new tbo_program());
    p->src.assign(nodes, nodes + n_nodes);
    if (lists && n_lists) p->lists.assign(lists, lists + n_lists);
    if (fixed_pool && fixed_len) p->pool.assign(fixed_pool, fixed_pool + fixed_len);
    p->sample_rate = sample_rate;
    int rc = build(p.get());
    if (rc != TB_OK) return rc;
    *out_program = p.release();
    return TB_OK;
}

void tbo_program_destroy(tbo_program* p) { delete p; }

int tbo_set_params(tbo_program* p, const float* params, uint32_t n_params) {
    for (auto& w : p->nodes) {
        if (w->kind == TB_CONST && w->param_slot >= 0) {
            if ((uint32_t)w->param_slot >= n_params) return TB_ERR_INVALID;
            w->value = params[w->param_slot];
        }
    }
    return TB_OK;
}

uint64_t tbo_generate(tbo_program* p, float* out, uint64_t n) {
    Gen g{p->sample_rate, 0, &p->rng, p->clean_tails};
    size_t len = g.generate(p->root, out, n);
    p->allocations += g.allocations;
    return len;
}

uint64_t tbo_length(tbo_program* p, uint64_t max) {
    Gen g{p->sample_rate, 0, &p->rng, p->clean_tails};
    size_t len = g.length(p->root, max);
    p->allocations += g.allocations;
    return len;
}

void tbo_set_state_initial(tbo_program* p) { Gen::set_state_initial(p->root); }

void tbo_initialize_state(tbo_program* p) {
    for (auto& w : p->nodes) {
        w->st = St::Initial;
        w->input.clear();
        w->output.clear();
        w->noise_drawn = 0;  // a fresh tree replays its noise streams (the reference's global RNG would not)
    }
}

int tbo_substitute_const(tbo_program* p, uint32_t mark_id, float value) {
    int hits = 0;
    const size_t n = p->nodes.size();  // new Const nodes are appended past n; root is tracked separately
    for (size_t i = 0; i < n; i++) {
        Node* w = p->nodes[i].get();
        if (w->kind == TB_MARKED && w->mark_id == mark_id) {
            p->nodes.emplace_back(new Node());
            Node* c = p->nodes.back().get();
            c->kind = TB_CONST;
            c->value = value;
            w->a = c;
            hits++;
        }
    }
    return hits;
}

uint64_t tbo_allocations(const tbo_program* p) { return p->allocations; }
void tbo_seed_noise(tbo_program* p, uint64_t seed) { p->rng.seed = seed; }
void tbo_set_clean_tails(tbo_program* p, int on) { p->clean_tails = on != 0; }
void tbo_set_voice(tbo_program* p, uint64_t voice) { p->rng.voice = voice; }

uint64_t tbo_render_batch(const tbo_program* p, const float* params, uint32_t n_params,
                          uint32_t n_voices, uint64_t n_samples, uint32_t block, float* out,
                          uint64_t out_stride, uint64_t* out_len, float* mix, uint32_t n_threads) {
    if (n_threads == 0) n_threads = 1;
    if (block == 0) block = 1024;
    std::atomic<uint32_t> next(0);
    std::atomic<uint64_t> total(0);
    std::vector<std::vector<float>> mixes(mix ? n_threads : 0);
    auto worker = [&](uint32_t tid) {
        tbo_program* q = nullptr;
        if (tbo_program_create(p->src.data(), (uint32_t)p->src.size(), p->lists.data(),
                               (uint32_t)p->lists.size(), p->pool.data(), p->pool.size(),
                               p->sample_rate, &q) != TB_OK)
            return;
        std::vector<float> scratch(block);
        if (mix) mixes[tid].assign(n_samples, 0.0f);
        uint64_t mine = 0;
        for (;;) {
            uint32_t v = next.fetch_add(1);
            if (v >= n_voices) break;
            tbo_initialize_state(q);
            q->rng.seed = p->rng.seed;
            q->rng.voice = p->rng.voice + v;  // voice v of the batch draws its own noise streams
            if (params) tbo_set_params(q, params + (size_t)v * n_params, n_params);
            uint64_t done = 0;
            while (done < n_samples) {
                uint64_t want = std::min<uint64_t>(block, n_samples - done);
                float* dst = out ? out + (size_t)v * out_stride + done : scratch.data();
                uint64_t got = tbo_generate(q, dst, want);
                if (mix)
                    for (uint64_t i = 0; i < got; i++) mixes[tid][done + i] += dst[i];
                done += got;
                if (got < want) break;
            }
            if (out_len) out_len[v] = done;
            mine += done;
        }
        total += mine;
        tbo_program_destroy(q);
    };
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < n_threads; t++) th.emplace_back(worker, t);
    worker(0);
    for (auto& t : th) t.join();
    if (mix) {
        for (uint64_t i = 0; i < n_samples; i++) {
            float s = 0.0f;
            for (uint32_t t = 0; t < n_threads; t++) s += mixes[t][i];
            mix[i] = s;
        }
    }
    return total.load();
}

}  // extern "C"
