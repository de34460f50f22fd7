// tuun_b200.hpp — C++17 host side above the C ABI (include/tuun_b200.h), header only.
//
// The reference is compiled code (Rust) whose toolchain is absent from this image, so the compiled
// host mirror of its generator interface is C++:
//
//   enum Waveform / enum Operator      src/lib/waveform.rs:5-100     -> tuun_b200::Waveform, Operator
//   generator::initialize_state        src/lib/generator.rs:39       -> Generator::initialize_state
//   Generator::new / generate / length src/lib/generator.rs:68,86,620-> Generator
//   waveform::set_state(.., Initial)   src/lib/waveform.rs:322       -> Program::reset
//
// Same names, argument meaning and error behaviour: `generate` fills the slice and returns the
// number of samples generated (short = finished, tail undefined, generator.rs:76-95); what the
// reference reports by panicking (malformed trees, generator.rs:112,132,189,...) surfaces here as
// tuun_b200::Error carrying the tb_status.  Trees are immutable shared nodes; like the reference's
// boxed trees, a subtree that appears twice is two instances with their own state.
#ifndef TUUN_B200_HPP
#define TUUN_B200_HPP

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "tuun_b200.h"

namespace tuun_b200 {

enum class Operator : uint32_t {  // waveform.rs:5-19
    Add = TB_ADD, Subtract = TB_SUBTRACT, Multiply = TB_MULTIPLY, Divide = TB_DIVIDE, Merge = TB_MERGE,
    Power = TB_POWER
};

struct Node;
using Waveform = std::shared_ptr<const Node>;

struct Node {  // one variant instance of `enum Waveform` (waveform.rs:23-100)
    tb_kind kind = TB_CONST;
    Operator op = Operator::Add;
    float value = 0.0f;
    int32_t param_slot = -1;
    uint32_t mark_id = 0;
    Waveform a, b, c;
    std::vector<Waveform> feed_forward, feedback;
    std::vector<float> samples;
};

namespace detail {
inline Waveform make(tb_kind k, Waveform a = nullptr, Waveform b = nullptr, Waveform c = nullptr) {
    auto n = std::make_shared<Node>();
    n->kind = k;
    n->a = std::move(a);
    n->b = std::move(b);
    n->c = std::move(c);
    return n;
}
}  // namespace detail

// Constructors named after the variants.
inline Waveform Const(float v, int32_t param_slot = -1) {
    auto n = std::make_shared<Node>();
    n->kind = TB_CONST;
    n->value = v;
    n->param_slot = param_slot;
    return n;
}
inline Waveform Time() { return detail::make(TB_TIME); }
inline Waveform Noise() { return detail::make(TB_NOISE); }
inline Waveform Fixed(std::vector<float> samples) {
    auto n = std::make_shared<Node>();
    n->kind = TB_FIXED;
    n->samples = std::move(samples);
    return n;
}
inline Waveform Fin(Waveform length, Waveform waveform) { return detail::make(TB_FIN, std::move(length), std::move(waveform)); }
inline Waveform Append(Waveform a, Waveform b) { return detail::make(TB_APPEND, std::move(a), std::move(b)); }
inline Waveform Sine(Waveform frequency, Waveform phase) { return detail::make(TB_SINE, std::move(frequency), std::move(phase)); }
inline Waveform Filter(Waveform waveform, std::vector<Waveform> feed_forward, std::vector<Waveform> feedback = {}) {
    auto n = std::make_shared<Node>();
    n->kind = TB_FILTER;
    n->a = std::move(waveform);
    n->feed_forward = std::move(feed_forward);
    n->feedback = std::move(feedback);
    return n;
}
inline Waveform BinaryPointOp(Operator op, Waveform a, Waveform b) {
    auto n = std::make_shared<Node>();
    n->kind = TB_BINARY;
    n->op = op;
    n->a = std::move(a);
    n->b = std::move(b);
    return n;
}
inline Waveform Reset(Waveform trigger, Waveform waveform) { return detail::make(TB_RESET, std::move(trigger), std::move(waveform)); }
inline Waveform Alt(Waveform trigger, Waveform positive, Waveform negative) {
    return detail::make(TB_ALT, std::move(trigger), std::move(positive), std::move(negative));
}
inline Waveform Marked(uint32_t id, Waveform waveform) {
    auto n = std::make_shared<Node>();
    n->kind = TB_MARKED;
    n->mark_id = id;
    n->a = std::move(waveform);
    return n;
}
inline Waveform Captured(uint32_t file_stem_id, Waveform waveform) {
    auto n = std::make_shared<Node>();
    n->kind = TB_CAPTURED;
    n->mark_id = file_stem_id;
    n->a = std::move(waveform);
    return n;
}

// The flat op list that crosses the ABI: children before parents, root last.
struct OpList {
    std::vector<tb_node> nodes;
    std::vector<int32_t> lists;
    std::vector<float> fixed_pool;
};

namespace detail {
inline int32_t flatten_into(const Waveform& w, OpList& o) {
    if (!w) throw std::invalid_argument("null waveform");
    tb_node n{};
    n.kind = (uint32_t)w->kind;
    n.a = n.b = n.c = -1;
    n.param_slot = -1;
    switch (w->kind) {
        case TB_CONST:
            n.value = w->value;
            n.param_slot = w->param_slot;
            break;
        case TB_TIME:
        case TB_NOISE: break;
        case TB_FIXED:
            n.fixed_off = o.fixed_pool.size();
            n.fixed_len = w->samples.size();
            o.fixed_pool.insert(o.fixed_pool.end(), w->samples.begin(), w->samples.end());
            break;
        case TB_FILTER: {
            n.a = flatten_into(w->a, o);
            std::vector<int32_t> idx;
            for (const auto& c : w->feed_forward) idx.push_back(flatten_into(c, o));
            for (const auto& c : w->feedback) idx.push_back(flatten_into(c, o));
            n.list_off = (uint32_t)o.lists.size();
            n.ff_count = (uint32_t)w->feed_forward.size();
            n.fb_count = (uint32_t)w->feedback.size();
            o.lists.insert(o.lists.end(), idx.begin(), idx.end());
            break;
        }
        case TB_BINARY:
            n.op = (uint32_t)w->op;
            n.a = flatten_into(w->a, o);
            n.b = flatten_into(w->b, o);
            break;
        case TB_ALT:
            n.a = flatten_into(w->a, o);
            n.b = flatten_into(w->b, o);
            n.c = flatten_into(w->c, o);
            break;
        case TB_MARKED:
        case TB_CAPTURED:
            n.mark_id = w->mark_id;
            n.a = flatten_into(w->a, o);
            break;
        default:  // Fin, Append, Sine, Reset
            n.a = flatten_into(w->a, o);
            n.b = flatten_into(w->b, o);
            break;
    }
    o.nodes.push_back(n);
    return (int32_t)o.nodes.size() - 1;
}
}  // namespace detail

inline OpList flatten(const Waveform& root) {
    OpList o;
    detail::flatten_into(root, o);
    return o;
}

class Error : public std::runtime_error {
public:
    Error(int status, const std::string& what) : std::runtime_error(what), status_(status) {}
    int status() const { return status_; }

private:
    int status_;
};

inline void check(int status) {
    if (status != TB_OK) throw Error(status, tb_last_error());
}

// Host-only validation + lowering (no device needed).
inline tb_program_info lower_check(const Waveform& w) {
    OpList o = flatten(w);
    tb_program_info info{};
    check(tb_lower_check(o.nodes.data(), (uint32_t)o.nodes.size(), o.lists.data(), (uint32_t)o.lists.size(),
                         o.fixed_pool.size(), &info));
    return info;
}

// `Waveform<M, State>`: the tree plus its carried state, living on the device.
class Program {
public:
    Program(const Waveform& w, uint32_t sample_rate, int device = -1) {
        OpList o = flatten(w);
        check(tb_program_create(o.nodes.data(), (uint32_t)o.nodes.size(), o.lists.data(), (uint32_t)o.lists.size(),
                                o.fixed_pool.data(), o.fixed_pool.size(), sample_rate, device, &h_));
    }
    ~Program() { tb_program_destroy(h_); }
    Program(Program&& other) noexcept : h_(other.h_) { other.h_ = nullptr; }
    Program& operator=(Program&& other) noexcept {
        if (this != &other) {
            tb_program_destroy(h_);
            h_ = other.h_;
            other.h_ = nullptr;
        }
        return *this;
    }
    Program(const Program&) = delete;
    Program& operator=(const Program&) = delete;

    // Batch render into host rows: out[v * stride + i]; returns per-voice generated lengths.
    std::vector<uint64_t> render(float* out, uint32_t n_voices, uint64_t n_samples, uint64_t stride,
                                 const float* params = nullptr, uint32_t n_params = 0, uint32_t flags = 0) {
        std::vector<uint64_t> len(n_voices);
        check(tb_render(h_, params, n_params, n_voices, n_samples, out, stride, len.data(), flags));
        return len;
    }
    std::vector<uint64_t> render_mix(float* mix, uint32_t n_voices, uint64_t n_samples, const float* params = nullptr,
                                     uint32_t n_params = 0, float* rows = nullptr, uint64_t stride = 0) {
        std::vector<uint64_t> len(n_voices);
        check(tb_render_mix(h_, params, n_params, n_voices, n_samples, rows, stride, len.data(), mix,
                            rows ? 0u : TB_NO_VOICE_OUT));
        return len;
    }
    std::vector<uint64_t> length(uint32_t n_voices, uint64_t max, const float* params = nullptr, uint32_t n_params = 0) {
        std::vector<uint64_t> len(n_voices);
        check(tb_length(h_, params, n_params, n_voices, max, len.data(), 0));
        return len;
    }
    void reset() { check(tb_reset(h_)); }  // waveform::set_state(root, Initial)
    // waveform::substitute(&mut w, &mark_id, &Const(value)) (waveform.rs:396): returns the nodes replaced
    uint32_t substitute(uint32_t mark_id, float value) {
        uint32_t n = 0;
        check(tb_substitute(h_, mark_id, value, &n));
        return n;
    }
    tb_program_info info() const {
        tb_program_info i{};
        check(tb_program_get_info(h_, &i));
        return i;
    }
    tb_program* handle() { return h_; }

private:
    tb_program* h_ = nullptr;
};

// `Generator::new(sample_rate)` — one waveform at a time, like the reference.
class Generator {
public:
    explicit Generator(uint32_t sample_rate, int device = -1) : sample_rate_(sample_rate), device_(device) {}
    Program initialize_state(const Waveform& w) const { return Program(w, sample_rate_, device_); }
    // Fills out[0..n) and returns the number of samples generated (generator.rs:86).
    size_t generate(Program& w, float* out, size_t n) const {
        if (n == 0) return 0;  // generator.rs:93-95
        return (size_t)w.render(out, 1, n, n)[0];
    }
    size_t generate(Program& w, std::vector<float>& out) const { return generate(w, out.data(), out.size()); }
    // Advance by `max` samples without output; returns the length generated (generator.rs:620).
    size_t length(Program& w, size_t max) const { return (size_t)w.length(1, max)[0]; }

private:
    uint32_t sample_rate_;
    int device_;
};

}  // namespace tuun_b200
#endif  // TUUN_B200_HPP
