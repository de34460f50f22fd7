// golden_test.cpp — the reference's own known-answer tests for the generator
// (/root/reference/src/lib/generator.rs:1284-1925), restated over the C++ host mirror
// (include/tuun_b200.hpp) and run through the C ABI on the device.
//
//   golden_test              run every case on cuda:0 (chunk sizes 1, 2, 4, 8; exact f32 equality)
//   golden_test --host-only  no device: flatten + tb_lower_check of every case
//
// Same harness shape as the reference's `run_tests` (generator.rs:1284-1351): sample_rate 1, the
// output pre-filled with +inf, `check_length` first, then chunked generation from Initial state.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "../../include/tuun_b200.hpp"

using namespace tuun_b200;

static const float F32_TAU = 6.28318530717958647692f;  // f32::consts::TAU
static const float F32_PI = 3.14159265358979323846f;

static Waveform sin_waveform(float frequency, float phase) {  // generator.rs:1467-1477
    return Sine(BinaryPointOp(Operator::Multiply, Const(F32_TAU), Const(frequency)), Const(phase));
}
static Waveform time_minus(float c) { return BinaryPointOp(Operator::Subtract, Time(), Const(c)); }
static std::vector<Waveform> consts(int n, float v) {
    std::vector<Waveform> r;
    for (int i = 0; i < n; i++) r.push_back(Const(v));
    return r;
}
static std::vector<float> rep(std::vector<float> v, int times) {
    std::vector<float> r;
    for (int i = 0; i < times; i++) r.insert(r.end(), v.begin(), v.end());
    return r;
}

struct Case {
    std::string name;
    Waveform w;
    std::vector<float> expected;
};

static std::vector<Case> cases() {
    std::vector<Case> c;
    auto add = [&](const char* name, Waveform w, std::vector<float> e) { c.push_back({name, std::move(w), std::move(e)}); };
    add("time", Time(), {0, 1, 2, 3, 4, 5, 6, 7});                                     // test_time :1354
    add("fixed", Fixed({1, 2, 3, 4, 5}), {1, 2, 3, 4, 5});                             // test_fixed :1360
    add("fin_marked",                                                                    // test_fin :1375-1396
        BinaryPointOp(Operator::Multiply, Const(2.0f),
                      Append(Fin(BinaryPointOp(Operator::Subtract, Time(), Marked(1, Const(4.0f))), Const(1.0f)),
                             Fixed({1.0f, 0.75f, 0.5f, 0.25f}))),
        {2, 2, 2, 2, 2, 1.5f, 1, 0.5f});
    // test_reset :1543-1599
    add("reset_time", Reset(sin_waveform(0.25f, 0.0f), Time()), {0, 1, 2, 3, 0, 1, 2, 3});
    add("reset_fin_trigger", Reset(Fin(time_minus(6.0f), sin_waveform(0.25f, 0.0f)), Time()), {0, 1, 2, 3, 0, 1});
    add("reset_fin_inner", Reset(sin_waveform(0.25f, 0.0f), Fin(time_minus(3.0f), Time())), {0, 1, 2, 0, 0, 1, 2, 0});
    add("reset_phase_pi", Reset(sin_waveform(0.25f, F32_PI), Time()), {0, 1, 0, 1, 2, 3, 0, 1});
    add("reset_16", Reset(sin_waveform(0.25f, 0.0f), Time()), rep({0, 1, 2, 3}, 4));
    add("append", Append(Fixed({1, 1, 1}), Fixed({2, 2, 2})), {1, 1, 1, 2, 2, 2});   // test_append :1602
    // test_sum :1624-1675
    add("add_consts", BinaryPointOp(Operator::Add, Const(1.0f), Const(2.0f)), rep({3.0f}, 8));
    add("add_fixed_const", BinaryPointOp(Operator::Add, Fixed({1, 2, 3}), Const(10.0f)), {11, 12, 13});
    add("add_short_long", BinaryPointOp(Operator::Add, Fixed({1, 2}), Fixed({10, 20, 30})), {11, 22});
    add("add_long_short", BinaryPointOp(Operator::Add, Fixed({1, 2, 3}), Fixed({10, 20})), {11, 22});
    add("add_fin", Fin(time_minus(4.0f), BinaryPointOp(Operator::Add, Const(1.0f), Const(2.0f))), rep({3.0f}, 4));
    add("add_empty", BinaryPointOp(Operator::Add, Fixed({}), Const(5.0f)), {});
    // test_dot_product :1678-1736
    add("mul_fin", Fin(time_minus(8.0f), BinaryPointOp(Operator::Multiply, Const(3.0f), Const(2.0f))), rep({6.0f}, 8));
    add("mul_fixed_const", BinaryPointOp(Operator::Multiply, Fixed({3, 4, 5}), Const(2.0f)), {6, 8, 10});
    add("mul_short_long", BinaryPointOp(Operator::Multiply, Fixed({3, 4}), Fixed({2, 5, 1})), {6, 20});
    add("mul_empty", BinaryPointOp(Operator::Multiply, Fixed({}), Const(5.0f)), {});
    // test_merge :1739-1777
    add("merge_consts", BinaryPointOp(Operator::Merge, Const(1.0f), Const(2.0f)), rep({3.0f}, 8));
    add("merge_short_long", BinaryPointOp(Operator::Merge, Fixed({1, 2}), Fixed({10, 20, 30})), {11, 22, 30});
    add("merge_fixed_const", BinaryPointOp(Operator::Merge, Fixed({1, 2}), Const(10.0f)), {11, 12, 10, 10, 10, 10, 10, 10});
    add("merge_same", BinaryPointOp(Operator::Merge, Fixed({1, 2}), Fixed({10, 20})), {11, 22});
    add("merge_empty", BinaryPointOp(Operator::Merge, Fixed({}), Fixed({10, 20})), {10, 20});
    // test_filter :1780-1903
    add("fir3_time", Filter(Time(), consts(3, 2.0f)), {6, 12, 18, 24, 30, 36, 42, 48});
    add("fir3_fin5", Filter(Fin(time_minus(5.0f), Time()), consts(3, 2.0f)), {6, 12, 18, 14, 8});
    add("fir5_fin8", Filter(Fin(time_minus(8.0f), Time()), consts(5, 2.0f)), {20, 30, 40, 50, 44, 36, 26, 14});
    add("fir2_over_reset", Filter(Reset(sin_waveform(1.0f / 3.0f, 3.0f * F32_PI / 2.0f), Time()), consts(2, 2.0f)),
        {0, 2, 6, 4, 2, 6, 4, 2});
    add("moving_average", Filter(Const(1.0f), consts(5, 0.2f)), rep({1.0f}, 8));
    add("iir_1_1", Filter(Time(), {Const(0.5f)}, {Const(-0.5f)}),
        {0.0f, 0.5f, 1.25f, 2.125f, 3.0625f, 4.03125f, 5.015625f, 6.0078125f});
    add("iir_cascade", Filter(Filter(Time(), {Const(0.5f)}, {Const(-0.5f)}), {Const(0.4f)}, {Const(-0.6f)}),
        {0.0f, 0.2f, 0.62f, 1.222f, 1.9582f, 2.7874203f, 3.6787024f, 4.610347f});
    add("fir_time_coeff", Filter(Const(1.0f), {Const(1.0f), Time()}), {1, 2, 3, 4, 5, 6, 7, 8});
    add("fir_finite_coeffs", Filter(Fixed({1, 1, 1}), {Const(1.0f), Fixed({2.0f}), Fixed({3.0f, 3.0f})}), {6, 3, 0});
    return c;
}

static int failures = 0;
#define EXPECT(cond, ...)                    \
    do {                                     \
        if (!(cond)) {                       \
            failures++;                      \
            std::fprintf(stderr, "FAIL: ");  \
            std::fprintf(stderr, __VA_ARGS__); \
            std::fprintf(stderr, "\n");      \
        }                                    \
    } while (0)

// generator.rs:1284-1351
static void run_tests(const Case& c) {
    Generator generator(1);
    {
        Program w = generator.initialize_state(c.w);  // check_length, :1290
        EXPECT(generator.length(w, c.expected.size()) == c.expected.size(), "%s: length", c.name.c_str());
    }
    for (size_t size : {1u, 2u, 4u, 8u}) {
        Program w = generator.initialize_state(c.w);
        std::vector<float> out(c.expected.size(), std::numeric_limits<float>::infinity());
        for (size_t n = 0; n < out.size() / size + 1; n++) {  // :1294-1298
            const size_t begin = n * size, end = std::min(out.size(), (n + 1) * size);
            const size_t got = generator.generate(w, out.data() + begin, end - begin);
            EXPECT(got == end - begin, "%s: chunk %zu of size %zu generated %zu", c.name.c_str(), n, size, got);
        }
        for (size_t i = 0; i < out.size(); i++)
            EXPECT(out[i] == c.expected[i], "%s (chunk %zu): sample %zu is %.9g, expected %.9g", c.name.c_str(), size, i,
                   out[i], c.expected[i]);
    }
}

// test_sine :1498-1540 — tolerance 1e-5 at 44.1 kHz against the analytic value.
static void test_sine() {
    Generator generator(44100);
    const int n = 100;
    const double tau = 6.283185307179586476925286766559;
    struct S { const char* name; Waveform w; double (*f)(double, int); };
    auto run = [&](const char* name, Waveform wf, auto analytic) {
        Program w = generator.initialize_state(wf);
        std::vector<float> out(n, std::numeric_limits<float>::infinity());
        EXPECT(generator.generate(w, out) == (size_t)n, "%s: length", name);
        for (int i = 0; i < n; i++) {
            const double want = analytic(i);
            EXPECT(std::fabs((double)out[i] - want) <= 1e-5, "%s: sample %d is %.9g, expected %.9g", name, i, out[i], want);
        }
    };
    run("sine_1hz", sin_waveform(1.0f, 0.0f), [&](int i) { return std::sin(tau * i / 44100.0); });
    run("sine_chirp",
        Sine(BinaryPointOp(Operator::Multiply, BinaryPointOp(Operator::Add, Time(), Const(10.0f)), Const(F32_TAU)), Const(0.0f)),
        [&](int i) { const double t = i / 44100.0; return std::sin(tau * (0.5 * t * t + 10.0 * t)); });
    run("sine_phase_pi", sin_waveform(0.25f, F32_PI), [&](int i) { return std::sin(tau * 0.25 * i / 44100.0 + M_PI); });
}

// test_fixed :1364-1371 — an exhausted Fixed generates nothing, forever.
static void test_fixed_exhausted() {
    Generator generator(1);
    Program w = generator.initialize_state(Fixed({1, 2, 3, 4, 5}));
    std::vector<float> out(8, 0.0f);
    EXPECT(generator.generate(w, out) == 5, "fixed: first call");
    EXPECT(generator.generate(w, out) == 0, "fixed: exhausted");
    EXPECT(generator.generate(w, out.data(), 0) == 0, "fixed: empty slice");
}

int main(int argc, char** argv) {
    const bool host_only = argc > 1 && std::strcmp(argv[1], "--host-only") == 0;
    const std::vector<Case> cs = cases();
    if (host_only) {
        for (const Case& c : cs) {
            try {
                const OpList o = flatten(c.w);
                EXPECT(!o.nodes.empty() && o.nodes.back().kind == (uint32_t)c.w->kind, "%s: root is not last", c.name.c_str());
                const tb_program_info info = lower_check(c.w);
                EXPECT(info.n_nodes == o.nodes.size(), "%s: node count", c.name.c_str());
            } catch (const Error& e) {
                EXPECT(false, "%s: status %d: %s", c.name.c_str(), e.status(), e.what());
            }
        }
        try {  // what the reference rejects by panicking comes back as a status (generator.rs:233)
            lower_check(Filter(Time(), {}));
            EXPECT(false, "empty feed_forward accepted");
        } catch (const Error& e) {
            EXPECT(e.status() == TB_ERR_INVALID, "empty feed_forward: status %d", e.status());
        }
        try {  // a Filter inside a Reset lowers to the run-by-run form (generator.rs:288-316)
            lower_check(Reset(sin_waveform(1.0f, 0.0f), Filter(Noise(), consts(2, 0.5f))));
        } catch (const Error& e) {
            EXPECT(false, "Filter inside a Reset: status %d: %s", e.status(), e.what());
        }
        try {
            lower_check(Filter(Time(), consts(40, 0.025f)));
            EXPECT(false, "40-tap Filter accepted");
        } catch (const Error& e) {
            EXPECT(e.status() == TB_ERR_UNSUPPORTED, "40-tap Filter: status %d", e.status());
        }
    } else {
        try {
            for (const Case& c : cs) run_tests(c);
            test_sine();
            test_fixed_exhausted();
        } catch (const Error& e) {
            EXPECT(false, "status %d: %s", e.status(), e.what());
        }
    }
    std::printf("%s: %zu cases, %d failures\n", host_only ? "host-only" : "device", cs.size(), failures);
    return failures == 0 ? 0 : 1;
}
