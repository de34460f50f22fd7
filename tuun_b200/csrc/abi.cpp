// abi.cpp — the C ABI of include/tuun_b200.h over the lowering (lower.cpp) and the kernels
// (render.cu).  Plain CUDA runtime; no torch, no CPU fallback: every compute entry point fails
// with TB_ERR_CUDA when no device is usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/tuun_b200.h"
#include "lower.h"
#include "program.h"
#include "split.h"

extern "C" size_t tb_kernel_smem_bytes(uint32_t n_code, uint32_t n_slots, uint32_t aux_words,
                                       uint32_t n_cval, uint32_t state_words, uint32_t steady_ok,
                                       uint32_t warps);
extern "C" cudaError_t tb_kernel_launch(const tb_launch* P, size_t smem, uint32_t warps, cudaStream_t stream);
extern "C" size_t tb_lanes_smem_bytes(uint32_t n_lane_code, uint32_t w_words, uint32_t q_units, uint32_t slots);
extern "C" cudaError_t tb_lanes_launch(const tb_launch* P, size_t smem, int kind, cudaStream_t stream);
extern "C" cudaError_t tb_lanes_occupancy(size_t smem, int kind, int* blocks_per_sm, int* n_sm);
// lanes_fm_ws.cu: the fused FM voice as a phase warp and a tone warp per 32 voices
extern "C" size_t tb_lanes_fm_ws_smem_bytes(uint32_t n_lane_code, uint32_t w_words, uint32_t q_units, uint32_t slots);
extern "C" void tb_lanes_fm_ws_run(const tb_launch* P, size_t smem, cudaStream_t stream);
extern "C" cudaError_t tb_lanes_fm_ws_occupancy(size_t smem, int* blocks_per_sm);
extern "C" void tb_lanes_fm_ws_split_run(const tb_launch* P, size_t smem, cudaStream_t stream);  // lanes_fm_ws_split.cu
extern "C" cudaError_t tb_lanes_fm_ws_split_occupancy(size_t smem, int* blocks_per_sm);
extern "C" cudaError_t tb_len_set(unsigned long long* out_len, uint32_t n, unsigned long long add, int accumulate,
                                  cudaStream_t stream);
extern "C" cudaError_t tb_mix_launch(const float* rows, uint64_t stride, const unsigned long long* lens,
                                     uint32_t n_voices, uint64_t n_samples, uint64_t t0, float* mix,
                                     int accumulate, cudaStream_t stream);

namespace {

thread_local std::string g_error = "";

int set_error(int status, const std::string& msg) {
    g_error = msg;
    return status;
}
int cuda_fail(cudaError_t e, const char* what) {
    return set_error(TB_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                       \
    do {                                               \
        cudaError_t e_ = (call);                       \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

template <class T>
int upload(const std::vector<T>& v, T** dst) {
    *dst = nullptr;
    const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    CU(cudaMalloc(reinterpret_cast<void**>(dst), bytes));
    if (!v.empty()) CU(cudaMemcpy(*dst, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return TB_OK;
}

}  // namespace

struct tb_program {
    tb::Lowered low;
    std::vector<uint32_t> noise_ids;  // a part of a sequence: the numbers of its nodes in the whole tree
    std::vector<tb_node> nodes;   // the op list as given (tb_substitute re-lowers it)
    std::vector<int32_t> lists;
    uint64_t fixed_len = 0;
    bool fast_sines = true;
    uint32_t sample_rate = 0;
    int device = 0;
    tb_insn* d_code = nullptr;
    tb_cexpr* d_cexpr = nullptr;
    tb_aux* d_aux = nullptr;
    tb_goe* d_goe = nullptr;
    int32_t* d_goe_steps = nullptr;
    tb_filter_tab* d_filt = nullptr;
    tb_fixed_tab* d_fixed = nullptr;
    float* d_pool = nullptr;
    tb_insn* d_lane_code = nullptr;
    tb_lane_aux* d_lane_aux = nullptr;
    size_t lane_smem = 0;           // 0: the lane-per-voice kernel does not apply to this program
    uint32_t lane_min_voices = 0;   // batches at least this large take it
    uint32_t lane_capacity = 0;     // CTAs of the lane interpreter kernels the device holds at once
    uint32_t lane_fm_capacity = 0;  // same for the fused-FM-voice kernel; 0: the program is not one fused FM voice
    uint32_t lane_fm_ws_capacity = 0;  // 32-voice CTAs of its two-warps-a-voice form (lanes_fm_ws.cu) the device holds; 0: not applicable
    size_t lane_fm_ws_smem = 0;
    uint32_t lane_fm_ws_split_capacity = 0;  // the same for virtual voices (lanes_fm_ws_split.cu)
    uint32_t lane_fm_ws_min_voices = 0;
    uint64_t fm_ws_launches = 0;
    uint32_t* d_lane_queue = nullptr;  // work queue of the persistent form (program.h tb_launch::lane_queue)
    size_t lane_queue_cap = 0;
    uint32_t* h_fault = nullptr;    // mapped pinned counter written by the lane kernel
    uint32_t* d_fault = nullptr;
    uint64_t lane_launches = 0;
    static constexpr uint32_t kLaneEvents = 64;  // ring of (start, stop) events around the last lane launches
    cudaEvent_t lane_ev[kLaneEvents][2] = {};
    uint32_t* d_state = nullptr;
    uint32_t n_voices = 0;
    bool fresh = true;  // no samples generated since create / tb_reset
    uint64_t stream_pos = 0;   // samples every voice has generated since create / tb_reset
    bool pos_known = true;     // false once tb_length has advanced the state (positions then differ per node)
    float* d_params = nullptr;
    size_t params_cap = 0;
    unsigned long long* d_len = nullptr;
    uint8_t* d_done = nullptr;
    float* d_mix = nullptr;
    size_t mix_cap = 0;
    float* d_stage[2] = {nullptr, nullptr};
    size_t stage_cap = 0;  // floats per staging buffer
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev_render[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    size_t smem = 0;
    uint32_t warps = TB_WARPS_PER_CTA;  // voices per CTA; fewer when one voice needs a lot of shared memory
    uint64_t launches = 0;
    uint64_t noise_seed = 0x7475756E2545F491ull, noise_first_voice = 0;  // tb_seed_noise
    uint32_t fast_mode = 2;  // FAST-class sines: 1 = f32 polynomial, 2 = MUFU (TUUN_B200_FAST_SINES)
    // A root sequence (lower.h sequence_parts): every part is a program of its own, rendered where it starts.
    std::vector<tb_program*> part_prog;
    std::vector<uint64_t> part_len;   // samples; ~0 for the last part
    uint64_t seq_renders = 0;         // generate launches that went part by part
    bool seq_only = false;            // the tree lowers only part by part (as one program it is nested too deeply): `low` is a stand-in
    static constexpr int kSeqStreams = 16;  // the parts of one call are independent streams: rendered side by side
    cudaStream_t seq_stream[kSeqStreams] = {};
    cudaEvent_t seq_fork = nullptr, seq_join[kSeqStreams] = {};
    // time-axis split (split.cu): per-segment state blocks and scratch, grown on demand
    tb_split_entry* d_split = nullptr;
    uint32_t* d_vs = nullptr;
    uint32_t* d_vi = nullptr;
    unsigned long long* d_vlen = nullptr;
    size_t vstate_cap = 0;          // virtual voices the three buffers hold
    uint32_t* d_vw = nullptr;       // fused FM voice: warm-up states, carrier snapshots, longest warm-up needed
    unsigned long long* d_snap = nullptr;
    uint32_t* d_warm_need = nullptr;
    size_t vwarm_cap = 0;
    uint64_t split_fm_rounds = 0;   // of split_rounds, those of the FM form (summary of phase sums + filter warm-up)
    uint64_t split_last_warm = 0;
    float* d_split_cval = nullptr;
    unsigned long long* d_split_inc = nullptr;
    size_t split_real_cap = 0;      // real voices the two scratch tables hold
    tb_split_args split_args = {};  // the split in progress (split_begin .. split_finish)
    uint32_t seg_voices = 0;        // tb_segments_begin .. tb_segments_end: voices of the sharded render, 0 = none
    bool seg_fm = false;            // ... in the form of render_split_fm (phase-sum pass, filter warm-up, samples)
    const float* seg_params = nullptr;
    uint32_t seg_n_params = 0;
    uint64_t split_rounds = 0;      // split renders so far (tb_program_info)
    uint32_t split_last_segments = 0;
    uint64_t split_last_seg_samples = 0;

    ~tb_program() {
        for (tb_program* q : part_prog) delete q;
        cudaSetDevice(device);
        for (int i = 0; i < kSeqStreams; i++) {
            if (seq_stream[i]) cudaStreamDestroy(seq_stream[i]);
            if (seq_join[i]) cudaEventDestroy(seq_join[i]);
        }
        if (seq_fork) cudaEventDestroy(seq_fork);
        cudaFree(d_code); cudaFree(d_cexpr); cudaFree(d_aux); cudaFree(d_goe); cudaFree(d_goe_steps);
        cudaFree(d_filt); cudaFree(d_fixed); cudaFree(d_pool); cudaFree(d_state); cudaFree(d_params);
        cudaFree(d_len); cudaFree(d_done); cudaFree(d_mix); cudaFree(d_stage[0]); cudaFree(d_stage[1]);
        cudaFree(d_lane_code); cudaFree(d_lane_aux); cudaFree(d_lane_queue);
        cudaFree(d_split); cudaFree(d_vs); cudaFree(d_vi); cudaFree(d_vlen); cudaFree(d_split_cval); cudaFree(d_split_inc);
        cudaFree(d_vw); cudaFree(d_snap); cudaFree(d_warm_need);
        if (h_fault) cudaFreeHost(h_fault);
        for (auto& e : lane_ev) {
            if (e[0]) cudaEventDestroy(e[0]);
            if (e[1]) cudaEventDestroy(e[1]);
        }
        for (int i = 0; i < 2; i++) {
            if (ev_render[i]) cudaEventDestroy(ev_render[i]);
            if (ev_copy[i]) cudaEventDestroy(ev_copy[i]);
        }
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }
};

namespace {

// Voices (warps) per CTA and the CTA's dynamic shared memory: as many warps as fit, up to
// TB_WARPS_PER_CTA.
bool size_cta(const tb::Lowered& low, uint32_t* warps, size_t* smem) {
    for (uint32_t w = TB_WARPS_PER_CTA; w >= 1; w >>= 1) {
        const size_t s = tb_kernel_smem_bytes((uint32_t)low.code.size(), low.n_slots, low.aux_words,
                                              (uint32_t)low.cexpr.size(), low.state_words, low.steady_ok || (low.lane_fin_goe >= 0 && !low.lane_clk), w);
        if (s <= 220 * 1024) {
            *warps = w;
            *smem = s;
            return true;
        }
    }
    return false;
}

int ensure_voices(tb_program* p, uint32_t n_voices) {
    if (p->n_voices == n_voices && p->d_state) {
        p->fresh = false;
        return TB_OK;
    }
    if (!p->fresh) return set_error(TB_ERR_STATE, "n_voices changed mid-stream; call tb_reset first");
    cudaFree(p->d_state);
    cudaFree(p->d_len);
    cudaFree(p->d_done);
    p->d_state = nullptr;
    p->d_len = nullptr;
    p->d_done = nullptr;
    p->n_voices = 0;
    const size_t words = (size_t)n_voices * p->low.state_words;
    CU(cudaMalloc(reinterpret_cast<void**>(&p->d_state), words * 4));
    CU(cudaMemsetAsync(p->d_state, 0, words * 4, p->stream));
    CU(cudaMalloc(reinterpret_cast<void**>(&p->d_len), (size_t)n_voices * 8));
    CU(cudaMalloc(reinterpret_cast<void**>(&p->d_done), (size_t)n_voices));
    p->n_voices = n_voices;
    p->fresh = false;
    return TB_OK;
}

int stage_params(tb_program* p, const float* params, uint32_t n_params, uint32_t n_voices, uint32_t flags,
                 const float** d_out) {
    *d_out = nullptr;
    if (!params) {
        if (p->low.n_params > 0 && n_params != 0) return set_error(TB_ERR_INVALID, "params is NULL but n_params != 0");
        return TB_OK;
    }
    if (n_params < p->low.n_params)
        return set_error(TB_ERR_INVALID, "parameter table has fewer columns than the program's param_slots");
    if (flags & TB_PARAMS_DEVICE) {
        *d_out = params;
        return TB_OK;
    }
    const size_t bytes = (size_t)n_voices * n_params * 4;
    if (bytes > p->params_cap) {
        cudaFree(p->d_params);
        p->d_params = nullptr;
        p->params_cap = 0;
        CU(cudaMalloc(reinterpret_cast<void**>(&p->d_params), std::max<size_t>(bytes, 16)));
        p->params_cap = bytes;
    }
    CU(cudaMemcpyAsync(p->d_params, params, bytes, cudaMemcpyHostToDevice, p->stream));
    *d_out = p->d_params;
    return TB_OK;
}

void fill_launch(const tb_program* p, tb_launch* L) {
    std::memset(L, 0, sizeof(*L));
    L->code = p->d_code;
    L->n_code = (uint32_t)p->low.code.size();
    L->pc_gen = p->low.pc_gen;
    L->pc_len = p->low.pc_len;
    L->pc_steady = p->low.pc_steady;
    L->cexpr = p->d_cexpr;
    L->n_cval = (uint32_t)p->low.cexpr.size();
    L->aux = p->d_aux;
    L->n_aux = (uint32_t)p->low.aux.size();
    L->aux_words = p->low.aux_words;
    L->goe = p->d_goe;
    L->goe_steps = p->d_goe_steps;
    L->filt = p->d_filt;
    L->fixed = p->d_fixed;
    L->pool = p->d_pool;
    L->n_slots = p->low.n_slots;
    L->state_words = p->low.state_words;
    L->sample_rate = p->sample_rate;
    L->pure_len = p->low.pure_len;
    L->n_filt = (uint32_t)p->low.filt.size();
    L->steady_ok = p->low.steady_ok;
    L->fast_mode = p->fast_mode;
    L->noise_seed = p->noise_seed;
    L->voice_base = p->noise_first_voice;
    L->state = p->d_state;
    L->lane_code = p->d_lane_code;
    L->lane_aux = p->d_lane_aux;
    L->n_lane_code = (uint32_t)p->low.lane_code.size();
    L->n_lane_aux = (uint32_t)p->low.lane_aux.size();
    L->lane_w_words = p->low.lane_w_words;
    L->lane_q_units = p->low.lane_q_units;
    L->lane_slots = p->low.lane_slots;
    L->fault = p->d_fault;
    L->lane_fin_goe = p->low.lane_fin_goe;
    L->lane_clk = p->low.lane_clk;
    L->lane_kscale = 17592186044416.0 / (6.283185307179586476925286766559 * (double)p->sample_rate);
}

int launch(tb_program* p, const tb_launch& L) {
    // Up to TB_WARPS_PER_CTA voices share a CTA (and its copy of the byte-code); a batch too small to give every
    // SM a CTA that way is spread out instead.
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0)
            n_sm = 148;
    }
    uint32_t warps = p->warps;
    size_t smem = p->smem;
    while (warps > 1 && (L.n_voices + warps - 1) / warps < 2u * (uint32_t)n_sm) warps >>= 1;
    if (warps != p->warps)
        smem = tb_kernel_smem_bytes((uint32_t)p->low.code.size(), p->low.n_slots, p->low.aux_words,
                                    (uint32_t)p->low.cexpr.size(), p->low.state_words,
                                    p->low.steady_ok || (p->low.lane_fin_goe >= 0 && !p->low.lane_clk), warps);
    cudaError_t e = tb_kernel_launch(&L, smem, warps, p->stream);
    if (e != cudaSuccess) return cuda_fail(e, "tb_render_kernel launch");
    p->launches++;
    return TB_OK;
}

// One launch of the lane kernel over B.n_samples (a multiple of TB_LS).  When the batch has more
// 64-voice groups than the device holds CTAs, the launch is cut into time segments handed out through a
// work queue (lanes_queue.cu): a plain grid would run a second, mostly empty wave for the whole
// duration of the render.
// One fused FM voice and a batch the device holds at once: the kernel of its own (lanes_fm.cu).
bool fm_kernel_applies(const tb_program* p, uint32_t n_voices) {
    const uint32_t groups = (n_voices + TB_LANE_THREADS - 1) / TB_LANE_THREADS;
    const char* qe = std::getenv("TUUN_B200_LANE_QUEUE");
    return p->lane_fm_capacity != 0 && groups <= p->lane_fm_capacity && !(qe && qe[0] == '1');
}

int launch_lanes(tb_program* p, tb_launch& B) {
    const uint32_t groups = (B.n_voices + TB_LANE_THREADS - 1) / TB_LANE_THREADS;
    const bool fm = fm_kernel_applies(p, B.n_voices);
    // ... as a phase warp and a tone warp per 32 voices when the device holds all those CTAs at once (lanes_fm_ws.cu)
    // (virtual voices — the segments of a time-axis split — too: lanes_fm_ws_split.cu; the phase-sum pass has its own kernel)
    const bool vs = B.vsplit_total > 1;
    const uint32_t ws_cap = vs ? p->lane_fm_ws_split_capacity : p->lane_fm_ws_capacity;
    const bool fm_ws = fm && ws_cap != 0 && (B.n_voices + 31u) / 32u <= ws_cap && !(vs && (B.fm_sums || B.mix_partial)) &&
                       B.n_voices >= p->lane_fm_ws_min_voices;
    const char* qe = std::getenv("TUUN_B200_LANE_QUEUE");  // diagnostics: "0" never, "1" always
    const bool want = !fm && B.vsplit_total <= 1 && (qe ? qe[0] == '1' : groups > p->lane_capacity);
    B.lane_queue = nullptr;
    if (want && B.n_samples >= 4 * 2 * TB_LS) {
        // 16 segments (or fewer, of at least 1024 samples): the last wave of units wastes < 1/16 of the time
        uint64_t segs = std::min<uint64_t>(16, std::max<uint64_t>(1, B.n_samples / 1024));
        uint64_t seg = (B.n_samples + segs - 1) / segs;
        seg = (seg + 2 * TB_LS - 1) / (2 * TB_LS) * (2 * TB_LS);
        segs = (B.n_samples + seg - 1) / seg;
        const size_t words = (size_t)groups + 1;
        if (words > p->lane_queue_cap) {
            cudaFree(p->d_lane_queue);
            p->d_lane_queue = nullptr;
            p->lane_queue_cap = 0;
            CU(cudaMalloc(reinterpret_cast<void**>(&p->d_lane_queue), words * 4));
            p->lane_queue_cap = words;
        }
        CU(cudaMemsetAsync(p->d_lane_queue, 0, words * 4, p->stream));
        B.lane_queue = p->d_lane_queue;
        B.lane_groups = groups;
        B.lane_segs = (uint32_t)segs;
        B.lane_seg_samples = seg;
        B.lane_grid = std::max<uint32_t>(1, p->lane_capacity);
    }
    cudaEvent_t* ev = p->lane_ev[p->lane_launches % tb_program::kLaneEvents];
    if (!ev[0]) {
        CU(cudaEventCreate(&ev[0]));
        CU(cudaEventCreate(&ev[1]));
    }
    CU(cudaEventRecord(ev[0], p->stream));
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return cuda_fail(e, "before the lane kernel launch");
    if (std::getenv("TUUN_B200_DEBUG"))
        std::fprintf(stderr, "[tuun_b200] lane launch: voices %u samples %llu smem %zu groups %u queue %d fm %d out %p stride %llu\n",
                     B.n_voices, (unsigned long long)B.n_samples, p->lane_smem, groups, B.lane_queue != nullptr, (int)fm,
                     (void*)B.out, (unsigned long long)B.out_stride);
    if (fm_ws) {
        if (vs) tb_lanes_fm_ws_split_run(&B, p->lane_fm_ws_smem, p->stream);
        else tb_lanes_fm_ws_run(&B, p->lane_fm_ws_smem, p->stream);
        e = cudaGetLastError();
        p->fm_ws_launches++;
    } else {
        e = tb_lanes_launch(&B, p->lane_smem, fm ? 2 : (B.lane_queue ? 1 : 0), p->stream);
    }
    if (e != cudaSuccess) return cuda_fail(e, "lane kernel launch");
    CU(cudaEventRecord(ev[1], p->stream));
    p->launches++;
    p->lane_launches++;
    return TB_OK;
}

// A generate launch.  Large batches of steady-state voices go through the lane-per-voice kernel
// (lanes.cuh).  `pos` = samples the voices have generated before this launch: the first general tile of
// a stream (filter pre-reads, generator.rs:234-252) stays on the warp-per-voice kernel, and so do the
// < TB_LS samples of a call that do not fill a lane tile; a stream that is past its first tile goes
// straight to the lane kernel (a caller streaming 1024-sample blocks pays one launch per block).
// State blocks are shared, so all launches continue one stream.
int launch_generate_seq(tb_program* p, const tb_launch& L, uint64_t pos) {
    // (a split pass of a program with clocked words — a Reset in the steady stream — has no other kernel to run on)
    const bool clk_split = p->low.lane_clk != 0 && L.vsplit_total > 1;
    const bool big = p->lane_smem != 0 && (L.n_voices >= p->lane_min_voices || clk_split) && (L.out != nullptr || L.state_only);
    if (!big) return launch(p, L);
    // The fused-FM-voice kernel starts a stream itself (the filter's read-ahead, run_fm_voice) and takes
    // the samples that do not fill a tile: one launch for the whole call.  (A root Fin keeps its general
    // head tile: the reference's per-call "already over?" test lives there.)
    if (fm_kernel_applies(p, L.n_voices) && p->low.lane_fin_goe < 0 && p->pos_known && L.n_samples >= TB_LS) {
        tb_launch B = L;
        B.done = nullptr;
        return launch_lanes(p, B);
    }
    // Primed: every filter holds its full history.  A root Fin decides "already over?" by the reference's
    // per-call test (generator.rs:808-809), which the general tile at the head of every call applies.
    const bool primed = p->pos_known && pos >= (uint64_t)TB_TILE && p->low.lane_fin_goe < 0;
    uint64_t head = primed ? 0 : TB_TILE;
    if (L.n_samples < head + TB_LS) {  // too short for a lane tile: same arithmetic, general tiles
        tb_launch G = L;
        G.exact_fb = 1;
        return launch(p, G);
    }
    // The samples that do not fill a lane tile go in front when that keeps the rows of the lane launch
    // 16-byte aligned (one general launch instead of two), else behind.
    const uint64_t rem = (L.n_samples - head) % TB_LS;
    if ((rem & 3) == 0) head += rem;
    const uint64_t bulk = (L.n_samples - head) / TB_LS * TB_LS;
    const uint64_t tail = L.n_samples - head - bulk;
    int rc = TB_OK;
    if (head) {
        tb_launch H = L;
        H.n_samples = head;
        H.exact_fb = 1;  // the bracketing tiles run the reference's recurrence too (see lanes.cuh)
        if ((rc = launch(p, H))) return rc;
    }
    tb_launch B = L;
    B.out = L.out + head;
    B.n_samples = bulk;
    B.accumulate = head ? 1 : L.accumulate;
    B.call_pos = L.call_pos + head;
    B.done = nullptr;  // the lane kernel renders finished voices too (their tails are undefined) and counts
    if ((rc = launch_lanes(p, B))) return rc;
    if (tail) {
        tb_launch T = L;
        T.out = L.out + head + bulk;
        T.n_samples = tail;
        T.accumulate = 1;
        T.exact_fb = 1;
        T.mid_call = 1;
        if ((rc = launch(p, T))) return rc;
    }
    return TB_OK;
}

// ---- time-axis split (split.cu) -------------------------------------------------------------------
// Few voices and many samples: a voice is one warp (or one thread), so a 60-second render of one voice
// would leave the device idle.  A steady program's state obeys associative laws over time (program.h
// tb_split_entry), so the call is cut into S = 2^k segments per voice, rendered side by side as V S
// "virtual voices" of the same kernels: `split_passes - 1` summary passes (final states only, no samples
// stored), each followed by a scan over the segments that makes one more level of state right, then the
// pass that writes the samples.  Phase sums are exact (u64), so sines are the samples of the unsplit render;
// filter histories come out of an f64 scan and differ from the serial f32 recurrence by its round-off noise.
struct SplitPlan {
    uint32_t n_seg = 0;   // S
    uint64_t seg = 0;     // samples per segment, a multiple of the steady tile
};
int n_sm_of_device() {
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0)
            n_sm = 148;
    }
    return n_sm;
}
// Segments for a call of n samples over V voices whose stream is past its first tile; false: render serially.
// TUUN_B200_SPLIT: "0" never, "S" about S segments when the call is long enough, unset: when the batch is small
// enough that the passes pay (V <= warps the device holds / (2 passes)).  Segments are whole steady tiles, so
// they start where the unsplit render's tiles start; what does not fill S segments is left to the caller
// (another round, or the serial form).
bool plan_split(const tb_program* p, const tb_launch& L, uint64_t n, SplitPlan* plan) {
    if (p->low.split_passes == 0 || !p->d_split || L.out == nullptr || L.done != nullptr || L.mode != 0 || L.vsplit > 1)
        return false;
    const uint64_t V = L.n_voices;
    const uint64_t cap = (uint64_t)n_sm_of_device() * 16;  // resident warps of the warp-per-voice kernel
    const char* env = std::getenv("TUUN_B200_SPLIT");
    uint64_t want = 0, min_seg = TB_TILE_S, grain = TB_TILE_S;
    const bool clk = p->low.lane_clk != 0;  // clocked words (a Reset in the steady stream): lane kernels only
    if (clk && p->lane_smem == 0) return false;
    bool lanes = clk;
    bool one_wave = false;
    if (clk) {  // one thread per segment, tiles of 16 samples (the serial lane path tiles the stream from sample 256 too)
        min_seg = 128;
        grain = 2 * TB_LS;
    }
    if (env) {
        want = std::strtoull(env, nullptr, 10);
        if (want < 2) return false;
    } else if (clk) {
        if (n < 4096 || V * 2 * p->low.split_passes > cap) return false;
        want = 32 * cap / V;
    } else {
        if (n < 4096 || V * 2 * p->low.split_passes > cap) return false;
        want = cap / V;     // one resident wave of warps: a second, partial wave costs a whole segment's time
        one_wave = true;
        // enough voice-samples for the lane-per-voice kernels (2.4 x the rate on FM voices): 4 lane_min_voices
        // virtual voices of at least 4096 samples
        if (p->lane_smem != 0 && p->lane_min_voices > 0 && V * (n / 4096) >= 4ull * p->lane_min_voices) {
            want = (4ull * p->lane_min_voices + V - 1) / V;
            min_seg = 4096;
            lanes = true;
            one_wave = false;
        }
    }
    // at least `want` segments (one_wave: at most) of whole tiles, none shorter than min_seg
    uint64_t seg = one_wave ? ((n + want - 1) / want + grain - 1) / grain * grain : n / want / grain * grain;
    seg = std::max<uint64_t>(seg, min_seg);
    uint64_t S = n / seg;
    // split.cu scans a voice's segments with one warp: beyond a few thousand the scan would take longer than the
    // passes (measured: 16,384 segments of a filtered pulse, 9 ms against 3.9 ms with 4,096)
    if (S > 4096) {
        S = 4096;
        seg = n / S / grain * grain;
    }
    (void)lanes;
    if (S < 2 || V * S > 0x7fffffffull) return false;
    plan->n_seg = (uint32_t)S;
    plan->seg = seg;
    return true;
}
int ensure_split_buffers(tb_program* p, uint64_t n_virtual, uint32_t n_real) {
    if (n_virtual > p->vstate_cap) {
        cudaFree(p->d_vs); cudaFree(p->d_vi); cudaFree(p->d_vlen);
        p->d_vs = p->d_vi = nullptr;
        p->d_vlen = nullptr;
        p->vstate_cap = 0;
        const size_t bytes = (size_t)n_virtual * p->low.state_words * 4;
        CU(cudaMalloc(reinterpret_cast<void**>(&p->d_vs), bytes));
        CU(cudaMalloc(reinterpret_cast<void**>(&p->d_vi), bytes));
        CU(cudaMalloc(reinterpret_cast<void**>(&p->d_vlen), (size_t)n_virtual * 8));
        p->vstate_cap = n_virtual;
    }
    if (n_real > p->split_real_cap) {
        cudaFree(p->d_split_cval); cudaFree(p->d_split_inc);
        p->d_split_cval = nullptr;
        p->d_split_inc = nullptr;
        p->split_real_cap = 0;
        CU(cudaMalloc(reinterpret_cast<void**>(&p->d_split_cval), (size_t)n_real * std::max<size_t>(p->low.cexpr.size(), 1) * 4));
        CU(cudaMalloc(reinterpret_cast<void**>(&p->d_split_inc), (size_t)n_real * std::max<size_t>(p->low.split.size(), 1) * 8));
        p->split_real_cap = n_real;
    }
    return TB_OK;
}
// The steps of a split render.  split_begin seeds the initial state of every segment; split_pass renders the
// segments [seg_lo, seg_hi) of every voice from them (pass = 1 .. split_passes; only the last one stores
// samples, at out_base[v * stride + s * seg + i]); split_fix makes the entries of level `pass` right from the
// final states of ALL segments; split_finish hands the last segment's final state back to the voice.
int split_begin(tb_program* p, const tb_launch& L, const SplitPlan& plan) {
    const uint32_t V = L.n_voices;
    const uint64_t nv = (uint64_t)V * plan.n_seg;
    int rc = ensure_split_buffers(p, nv, V);
    if (rc) return rc;
    tb_split_args& A = p->split_args;
    std::memset(&A, 0, sizeof(A));
    A.cexpr = p->d_cexpr;
    A.n_cval = (uint32_t)p->low.cexpr.size();
    A.entries = p->d_split;
    A.n_entries = (uint32_t)p->low.split.size();
    A.filt = p->d_filt;
    A.params = L.params;
    A.n_params = L.n_params;
    A.sample_rate = p->sample_rate;
    A.state_words = p->low.state_words;
    A.n_real = V;
    A.n_seg = plan.n_seg;
    A.seg = plan.seg;
    A.real_state = L.state;
    A.vi = p->d_vi;
    A.vs = p->d_vs;
    A.cval = p->d_split_cval;
    A.inc = p->d_split_inc;
    cudaError_t e = tb_split_seed(&A, p->stream);
    if (e != cudaSuccess) return cuda_fail(e, "tb_split_seed");
    p->launches += 2;
    return TB_OK;
}
struct PassOpt {
    bool fm_sums = false;      // fused FM voice: phase sums only, snapshot after snap_at samples (split.h)
    uint64_t snap_at = 0;
    uint32_t* states = nullptr;  // render these state blocks in place (no copy from d_vi) for `samples` samples
    uint64_t samples = 0;
};
int split_pass(tb_program* p, const tb_launch& L, uint32_t pass, uint32_t seg_lo, uint32_t seg_hi, float* out_base,
               uint64_t pos, const PassOpt& opt = PassOpt()) {
    const tb_split_args& A = p->split_args;
    const bool last = pass == p->low.split_passes;
    const size_t state_bytes = (size_t)A.n_real * A.n_seg * p->low.state_words * 4;
    if (!opt.states) CU(cudaMemcpyAsync(p->d_vs, p->d_vi, state_bytes, cudaMemcpyDeviceToDevice, p->stream));
    tb_launch B = L;
    B.state = opt.states ? opt.states : p->d_vs;
    B.n_voices = A.n_real * (seg_hi - seg_lo);
    B.n_samples = opt.states ? opt.samples : A.seg;
    B.fm_sums = opt.fm_sums ? 1u : 0u;
    B.vsnap = opt.fm_sums ? A.snap : nullptr;
    B.vsnap_at = opt.snap_at;
    B.vsplit = seg_hi - seg_lo;
    B.vsplit_total = A.n_seg;
    B.vseg_lo = seg_lo;
    B.vseg = A.seg;
    B.out = last ? out_base : nullptr;
    B.state_only = last ? 0 : 1;
    B.out_len = p->d_vlen;
    B.accumulate = 0;
    B.mid_call = 1;
    B.done = nullptr;
    return launch_generate_seq(p, B, std::max<uint64_t>(pos, TB_TILE));
}
int split_fix(tb_program* p, uint32_t pass) {
    cudaError_t e = tb_split_fix(&p->split_args, pass, p->stream);
    if (e != cudaSuccess) return cuda_fail(e, "tb_split_fix");
    p->launches++;
    return TB_OK;
}
int split_finish(tb_program* p, uint32_t* real_state, unsigned long long* out_len, uint64_t n, bool accumulate) {
    cudaError_t e = tb_split_finish(&p->split_args, real_state, out_len, n, accumulate ? 1 : 0, p->stream);
    if (e != cudaSuccess) return cuda_fail(e, "tb_split_finish");
    p->launches++;
    return TB_OK;
}
// One round: S seg samples of every voice of L (stream primed), rows at L.out.
int render_split_round(tb_program* p, const tb_launch& L, const SplitPlan& plan, uint64_t pos) {
    int rc = split_begin(p, L, plan);
    if (rc) return rc;
    for (uint32_t pass = 1; pass <= p->low.split_passes; pass++) {
        if ((rc = split_pass(p, L, pass, 0, plan.n_seg, L.out, pos))) return rc;
        if (pass < p->low.split_passes && (rc = split_fix(p, pass))) return rc;
    }
    if ((rc = split_finish(p, L.state, L.out_len, plan.seg * plan.n_seg, L.accumulate != 0))) return rc;
    p->split_rounds++;
    return TB_OK;
}

// ---- a batch of fused FM voices with a biquad, too small to fill the lane kernel (strong scaling: 65,536 voices over 8
// GPUs are 8,192 each, 1.7 warps an SM) ---------------------------------------------------------------------------------
// The general split would render such a batch three times.  Here the two summaries are cheap: (1) the carrier's phase
// sums come from a pass that computes nothing else (run_fm_sums: a third of a tile's instructions), and (2) the
// biquad's history at a segment's start comes from a warm-up — `warm` samples before the segment, from zero history,
// long enough for the filter to forget it (split.cu split_fm_need_kernel) — instead of a pass over the whole
// segment and an affine scan.  Sines stay bit-identical to the serial render (exact phase sums); the filter differs
// by < 1e-11 of its state plus, like any change of partition, its own round-off noise.
bool plan_split_fm(const tb_program* p, const tb_launch& L, uint64_t n, SplitPlan* plan) {
    if (p->low.split_passes != 3 || !p->d_split || p->lane_fm_capacity == 0 || L.out == nullptr || L.done != nullptr ||
        L.mode != 0 || L.vsplit_total > 1 || p->low.filt.size() != 1)
        return false;
    const char* env = std::getenv("TUUN_B200_SPLIT_FM");  // "0": never; "S": S segments whatever the batch
    uint64_t S = 0;
    const uint64_t V = L.n_voices;
    if (env) {
        S = std::strtoull(env, nullptr, 10);
        if (S < 2) return false;
    } else {
        if (std::getenv("TUUN_B200_SPLIT")) return false;          // the general form was asked for (or none)
        if (V > 12288 || n < 65536) return false;                   // large batches fill the device as they are
        S = 65536 / V;
    }
    uint64_t k = 0;
    while ((2ull << k) <= S) k++;
    S = 1ull << k;
    if (S > 4096) S = 4096;
    while (S >= 2 && n / S < 32768) S >>= 1;                        // room for the warm-up (a few thousand samples)
    if (S < 2) return false;
    // a warm-up needs long segments, so a handful of voices cannot be cut finely enough to fill the lane kernel this
    // way: they take the general form (three passes, segments of any length)
    if (!env && V * S < 16384) return false;
    plan->n_seg = (uint32_t)S;
    plan->seg = n / S / (2 * TB_LS) * (2 * TB_LS);
    return true;
}
int ensure_warm_buffers(tb_program* p, uint64_t nv) {
    if (nv > p->vwarm_cap) {
        cudaFree(p->d_vw); cudaFree(p->d_snap);
        p->d_vw = nullptr;
        p->d_snap = nullptr;
        p->vwarm_cap = 0;
        CU(cudaMalloc(reinterpret_cast<void**>(&p->d_vw), (size_t)nv * p->low.state_words * 4));
        CU(cudaMalloc(reinterpret_cast<void**>(&p->d_snap), (size_t)nv * 8));
        p->vwarm_cap = nv;
    }
    if (!p->d_warm_need) CU(cudaMalloc(reinterpret_cast<void**>(&p->d_warm_need), 4));
    return TB_OK;
}
// After split_begin: the warm-up the batch's filters need (split.cu split_fm_need_kernel), in A.warm; *ok = false when
// the form does not apply (a filter that does not forget fast enough: the warm-up would not be a small fraction of a
// segment).  Synchronises the stream once.
int prepare_split_fm(tb_program* p, uint64_t seg, bool* ok) {
    *ok = false;
    tb_split_args& A = p->split_args;
    int rc = ensure_warm_buffers(p, (uint64_t)A.n_real * A.n_seg);
    if (rc) return rc;
    A.vw = p->d_vw;
    A.snap = p->d_snap;
    A.warm_need = p->d_warm_need;
    A.fm_carrier = A.fm_filter = -1;
    for (size_t k = 0; k < p->low.split.size(); k++) {
        if (p->low.split[k].kind == SP_SINE_VAR) A.fm_carrier = (int32_t)k;
        if (p->low.split[k].kind == SP_FILTER) A.fm_filter = (int32_t)k;
    }
    if (A.fm_carrier < 0 || A.fm_filter < 0) return TB_OK;
    CU(cudaMemsetAsync(p->d_warm_need, 0, 4, p->stream));
    cudaError_t e = tb_split_fm_need(&A, p->stream);
    if (e != cudaSuccess) return cuda_fail(e, "tb_split_fm_need");
    uint32_t need = 0;
    CU(cudaMemcpyAsync(&need, p->d_warm_need, 4, cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    p->launches++;
    const uint64_t warm = ((uint64_t)need + TB_LS - 1) / TB_LS * TB_LS;
    if (need == 0xffffffffu || warm * 4 > seg) return TB_OK;
    A.warm = warm;
    *ok = true;
    return TB_OK;
}
// Returns TB_OK and *done = true when the round was rendered; *done = false: not applicable after all, nothing was changed.
int render_split_fm(tb_program* p, const tb_launch& L, const SplitPlan& plan, uint64_t pos, bool* done) {
    *done = false;
    int rc = split_begin(p, L, plan);
    if (rc) return rc;
    tb_split_args& A = p->split_args;
    bool ok = false;
    if ((rc = prepare_split_fm(p, plan.seg, &ok))) return rc;
    if (!ok) return TB_OK;
    const uint64_t warm = A.warm;
    cudaError_t e = cudaSuccess;
    // 1: phase sums (and the snapshot where the next segment's warm-up starts), then the exact starts
    PassOpt sums;
    sums.fm_sums = true;
    sums.snap_at = plan.seg - warm;
    if ((rc = split_pass(p, L, 1, 0, plan.n_seg - 1, nullptr, pos, sums))) return rc;  // nobody starts behind the last segment
    if ((rc = split_fix(p, 1))) return rc;
    // 2: the filters' histories by warm-up
    e = tb_split_fm_warm_seed(&A, p->stream);
    if (e != cudaSuccess) return cuda_fail(e, "tb_split_fm_warm_seed");
    PassOpt wp;
    wp.states = p->d_vw;
    wp.samples = warm;
    if ((rc = split_pass(p, L, 2, 0, plan.n_seg, nullptr, pos, wp))) return rc;
    e = tb_split_fm_adopt(&A, p->stream);
    if (e != cudaSuccess) return cuda_fail(e, "tb_split_fm_adopt");
    p->launches += 2;
    // 3: the samples
    if ((rc = split_pass(p, L, 3, 0, plan.n_seg, L.out, pos))) return rc;
    if ((rc = split_finish(p, L.state, L.out_len, plan.seg * plan.n_seg, L.accumulate != 0))) return rc;
    p->split_rounds++;
    p->split_fm_rounds++;
    p->split_last_warm = warm;
    *done = true;
    return TB_OK;
}

// A generate launch: split in time when that pays (rounds of S segments until what is left is short), else
// — and for the head tile of a stream and the rest — the serial form.
int launch_generate(tb_program* p, const tb_launch& L, uint64_t pos);
// A root sequence, part by part: part k is a stream of its own that starts at sample start_k = len_0 + .. + len_(k-1) of
// the voice's stream (the second arm of an Append starts from Initial state, generator.rs:169-188), so the samples
// [pos, pos + n) of the call belong to the parts they overlap, each rendered by its own (much smaller) program from
// its own carried state.  `pos` is the same for every voice: the lengths are analytic and voice-independent.
int launch_sequence(tb_program* p, const tb_launch& L, uint64_t pos) {
    const uint64_t v0 = (uint64_t)(L.state - p->d_state) / p->low.state_words;  // first voice of this launch in the batch
    // the parts this call overlaps
    struct Piece { size_t k; uint64_t a, b, start; };
    std::vector<Piece> pieces;
    uint64_t start = 0;
    for (size_t k = 0; k < p->part_prog.size(); k++) {
        const uint64_t len = p->part_len[k];
        const uint64_t end = len == ~0ull ? ~0ull : start + len;
        const uint64_t a = std::max(pos, start), b = std::min(pos + L.n_samples, end);
        if (a < b) pieces.push_back(Piece{k, a, b, start});
        if (end == ~0ull || end >= pos + L.n_samples) break;
        start = end;
    }
    if (pieces.empty()) return TB_OK;
    int rc = TB_OK;
    // Every part but the last one of the call fills its share whole (its Fin cannot end early), so out_len is what lies in
    // front of the last one plus what that one generates: set the former here, let the last part add.
    if (L.out_len) {
        cudaError_t e = tb_len_set(L.out_len, L.n_voices, pieces.back().a - pos, L.accumulate ? 1 : 0, p->stream);
        if (e != cudaSuccess) return cuda_fail(e, "tb_len_set");
        p->launches++;
    }
    CU(cudaEventRecord(p->seq_fork, p->stream));  // parameters staged, earlier renders done
    bool used[tb_program::kSeqStreams] = {};
    for (const Piece& pc : pieces) {
        tb_program* q = p->part_prog[pc.k];
        const int si = (int)(pc.k % tb_program::kSeqStreams);
        q->noise_seed = p->noise_seed;
        q->noise_first_voice = p->noise_first_voice;
        if (!used[si]) CU(cudaStreamWaitEvent(q->stream, p->seq_fork, 0));
        used[si] = true;
        if ((rc = ensure_voices(q, p->n_voices))) return rc;
        tb_launch Q;
        fill_launch(q, &Q);
        Q.params = L.params;
        Q.n_params = L.n_params;
        Q.n_voices = L.n_voices;
        Q.voice_base = L.voice_base;
        Q.state = q->d_state + v0 * q->low.state_words;
        Q.out = L.out ? L.out + (pc.a - pos) : nullptr;
        Q.out_stride = L.out_stride;
        Q.out_len = &pc == &pieces.back() ? L.out_len : nullptr;
        Q.done = L.done;
        Q.mode = 0;
        Q.n_samples = pc.b - pc.a;
        Q.accumulate = 1;
        Q.mid_call = (L.mid_call && pc.a == pos) ? 1u : 0u;
        Q.call_pos = L.call_pos + (pc.a - pos);
        if ((rc = launch_generate(q, Q, pc.a - pc.start))) return rc;
    }
    for (int si = 0; si < tb_program::kSeqStreams; si++) {
        if (!used[si]) continue;
        CU(cudaEventRecord(p->seq_join[si], p->seq_stream[si]));
        CU(cudaStreamWaitEvent(p->stream, p->seq_join[si], 0));
    }
    p->seq_renders++;
    return TB_OK;
}

int launch_generate(tb_program* p, const tb_launch& L, uint64_t pos) {
    if (!p->part_prog.empty() && p->pos_known && L.mode == 0) return launch_sequence(p, L, pos);
    if (p->seq_only) return set_error(TB_ERR_UNSUPPORTED, "this tree renders only part by part");
    if (p->low.split_passes == 0 || !p->pos_known) return launch_generate_seq(p, L, pos);
    // The first tile of a stream stays on the serial form: filter pre-reads (generator.rs:234-252).  (A program
    // without filters is steady from its first sample, and so are its segments — except under a Reset: its trigger
    // starts at phase 0 exactly, where the sign decides the first restart (generator.rs:296) and must come from the
    // exact phase, not from a rotated sin / cos pair.)
    const bool need_head = pos < (uint64_t)TB_TILE && (!p->low.filt.empty() || p->low.lane_clk);
    const uint64_t head = need_head ? (uint64_t)TB_TILE : 0;
    if (L.n_samples <= head) return launch_generate_seq(p, L, pos);
    SplitPlan plan;
    const bool fm = plan_split_fm(p, L, L.n_samples - head, &plan);
    if (!fm && !plan_split(p, L, L.n_samples - head, &plan)) return launch_generate_seq(p, L, pos);
    tb_launch R = L;  // what is left of the call
    int rc = TB_OK;
    auto advance = [&](uint64_t n) {
        R.out += n;
        R.n_samples -= n;
        R.accumulate = 1;
        R.mid_call = 1;
        R.call_pos += n;
        pos += n;
    };
    if (need_head) {
        tb_launch H = L;
        H.n_samples = head;
        if ((rc = launch_generate_seq(p, H, pos))) return rc;
        advance(head);
    }
    bool first = true;
    auto note = [&]() {  // tb_program_info: the round that covers most of the call
        if (first) {
            p->split_last_segments = plan.n_seg;
            p->split_last_seg_samples = plan.seg;
            first = false;
        }
    };
    if (fm) {
        bool done = false;
        if ((rc = render_split_fm(p, R, plan, pos, &done))) return rc;
        if (done) {
            note();
            advance(plan.seg * plan.n_seg);
        }
        return R.n_samples > 0 ? launch_generate_seq(p, R, pos) : TB_OK;
    }
    while (R.n_samples > 0 && plan_split(p, R, R.n_samples, &plan)) {
        note();
        if ((rc = render_split_round(p, R, plan, pos))) return rc;
        advance(plan.seg * plan.n_seg);
    }
    if (R.n_samples > 0) return launch_generate_seq(p, R, pos);
    return TB_OK;
}

// Mixdown of a large steady batch without rows (tb_render_mix, TB_NO_VOICE_OUT): the head of the call is
// rendered by the general kernel into a small staging block and mixed in voice order; everything after
// it is summed over the 32 voices of each warp inside the lane kernel (lanes.cuh mix_tile) and the
// per-warp partial rows are added in warp order.  No voice row of the lane part ever reaches memory.
int render_mix_lanes(tb_program* p, const tb_launch& L, float* d_mix, uint64_t pos) {
    // The fused-FM-voice kernel starts a stream itself and takes the samples that do not fill a tile (as for rows,
    // launch_generate_seq): one lane launch for the whole call, no general head.
    const bool fm_whole = fm_kernel_applies(p, L.n_voices) && p->pos_known && L.n_samples >= TB_LS;
    uint64_t head = fm_whole ? 0 : ((p->pos_known && pos >= (uint64_t)TB_TILE) ? 0 : TB_TILE);  // see launch_generate_seq
    if (!fm_whole) head += (L.n_samples - head) % TB_LS;
    const uint64_t bulk = L.n_samples - head;
    const uint64_t stride = (bulk + TB_LS - 1) / TB_LS * TB_LS;  // partial rows hold whole tiles
    const uint64_t n_warps = (L.n_voices + 31) / 32;
    const size_t need = std::max<size_t>((size_t)L.n_voices * head, (size_t)n_warps * stride);
    if (need > p->stage_cap) {
        for (int i = 0; i < 2; i++) {
            cudaFree(p->d_stage[i]);
            p->d_stage[i] = nullptr;
        }
        p->stage_cap = 0;
        for (int i = 0; i < 2; i++) CU(cudaMalloc(reinterpret_cast<void**>(&p->d_stage[i]), need * 4));
        p->stage_cap = need;
    }
    int rc = TB_OK;
    cudaError_t e = cudaSuccess;
    if (head) {
        tb_launch H = L;
        H.out = p->d_stage[0];
        H.out_stride = head;
        H.n_samples = head;
        H.exact_fb = 1;
        if ((rc = launch(p, H))) return rc;
        e = tb_mix_launch(p->d_stage[0], head, L.out_len, L.n_voices, head, 0, d_mix, 0, p->stream);
        if (e != cudaSuccess) return cuda_fail(e, "tb_mix_kernel launch");
        p->launches++;
    }
    tb_launch B = L;
    B.out = nullptr;
    B.out_stride = 0;
    B.n_samples = bulk;
    B.accumulate = head ? 1 : L.accumulate;
    B.done = nullptr;
    B.mix_partial = p->d_stage[1];
    B.mix_stride = stride;
    if ((rc = launch_lanes(p, B))) return rc;
    e = tb_mix_launch(p->d_stage[1], stride, nullptr, (uint32_t)n_warps, bulk, 0, d_mix + head, 0, p->stream);
    if (e != cudaSuccess) return cuda_fail(e, "tb_mix_kernel launch");
    p->launches++;
    return TB_OK;
}

int check_fault(tb_program* p) {
    if (p->h_fault && *p->h_fault != 0u) {
        *p->h_fault = 0u;
        return set_error(TB_ERR_STATE, "lane kernel met a voice without complete filter history");
    }
    return TB_OK;
}

}  // namespace

extern "C" {

uint32_t tb_abi_version(void) { return TB_ABI_VERSION; }
const char* tb_last_error(void) { return g_error.c_str(); }

static int create_program(const tb_node* nodes, uint32_t n_nodes, const int32_t* lists, uint32_t n_lists,
                          const float* fixed_pool, uint64_t fixed_len, uint32_t sample_rate, int device,
                          tb_program** out_program, const uint32_t* noise_ids, bool allow_sequence);

int tb_program_create(const tb_node* nodes, uint32_t n_nodes, const int32_t* lists, uint32_t n_lists,
                      const float* fixed_pool, uint64_t fixed_len, uint32_t sample_rate, int device,
                      tb_program** out_program) {
    return create_program(nodes, n_nodes, lists, n_lists, fixed_pool, fixed_len, sample_rate, device, out_program, nullptr,
                          true);
}

// The subtree rooted at `root` as an op list of its own (children before parents, Filter lists re-based); ids[i] =
// index the new node i had in the whole list (the number of its Noise stream).
static void extract_subtree(const std::vector<tb_node>& nodes, const std::vector<int32_t>& lists, int root,
                            std::vector<tb_node>& sub, std::vector<int32_t>& sub_lists, std::vector<uint32_t>& ids) {
    std::vector<char> in(nodes.size(), 0);
    std::vector<int> stack{root};
    while (!stack.empty()) {
        const int i = stack.back();
        stack.pop_back();
        if (in[i]) continue;
        in[i] = 1;
        const tb_node& n = nodes[i];
        for (int c : {n.a, n.b, n.c})
            if (c >= 0 && n.kind != TB_CONST && n.kind != TB_TIME && n.kind != TB_NOISE && n.kind != TB_FIXED) stack.push_back(c);
        if (n.kind == TB_FILTER)
            for (uint32_t j = 0; j < n.ff_count + n.fb_count; j++) stack.push_back(lists[n.list_off + j]);
    }
    std::vector<int> map(nodes.size(), -1);
    sub.clear();
    sub_lists.clear();
    ids.clear();
    for (int i = 0; i <= root; i++) {
        if (!in[i]) continue;
        tb_node n = nodes[i];
        const bool leaf = n.kind == TB_CONST || n.kind == TB_TIME || n.kind == TB_NOISE || n.kind == TB_FIXED;
        if (!leaf) {
            if (n.a >= 0) n.a = map[n.a];
            if (n.b >= 0) n.b = map[n.b];
            if (n.c >= 0) n.c = map[n.c];
        }
        if (n.kind == TB_FILTER) {
            const uint32_t off = (uint32_t)sub_lists.size();
            for (uint32_t j = 0; j < n.ff_count + n.fb_count; j++) sub_lists.push_back(map[lists[n.list_off + j]]);
            n.list_off = off;
        }
        map[i] = (int)sub.size();
        sub.push_back(n);
        ids.push_back((uint32_t)i);
    }
}

static int create_program(const tb_node* nodes, uint32_t n_nodes, const int32_t* lists, uint32_t n_lists,
                          const float* fixed_pool, uint64_t fixed_len, uint32_t sample_rate, int device,
                          tb_program** out_program, const uint32_t* noise_ids, bool allow_sequence) {
    if (!out_program) return set_error(TB_ERR_INVALID, "out_program is NULL");
    *out_program = nullptr;
    if (!nodes || n_nodes == 0) return set_error(TB_ERR_INVALID, "empty op list");
    if (sample_rate == 0 || sample_rate >= 0x80000000u) return set_error(TB_ERR_INVALID, "sample_rate must be > 0 and below 2^31");
    if (fixed_len > 0 && !fixed_pool) return set_error(TB_ERR_INVALID, "fixed_pool is NULL");
    tb_program* p = new (std::nothrow) tb_program();
    if (!p) return set_error(TB_ERR_NOMEM, "out of host memory");
    const char* fs = std::getenv("TUUN_B200_FAST_SINES");
    const bool fast = !(fs && fs[0] == '0');
    int rc = tb::lower(nodes, n_nodes, lists, n_lists, fixed_len, fast, p->low, noise_ids, sample_rate);
    std::string whole_error;
    if (rc == TB_ERR_UNSUPPORTED && allow_sequence) {
        // A long tune nests deeper than the interpreter's control stack as ONE program, but lowers part by part: the
        // program then has no whole-tree form at all (`low` becomes a stand-in with one state word).
        std::vector<tb::SeqPart> probe;
        const char* sq0 = std::getenv("TUUN_B200_SEQ");
        if (!(sq0 && sq0[0] == '0') && tb::sequence_parts(nodes, n_nodes, lists, n_lists, fixed_len, sample_rate, probe)) {
            whole_error = p->low.error;
            tb_node zero;
            std::memset(&zero, 0, sizeof(zero));
            zero.kind = TB_CONST;
            zero.a = zero.b = zero.c = -1;
            zero.param_slot = -1;
            rc = tb::lower(&zero, 1, nullptr, 0, 0, fast, p->low, nullptr);
            p->seq_only = rc == TB_OK;
        }
    }
    if (rc != TB_OK) {
        std::string msg = whole_error.empty() ? p->low.error : whole_error;
        delete p;
        return set_error(rc, msg);
    }
    if (noise_ids) p->noise_ids.assign(noise_ids, noise_ids + n_nodes);
    p->nodes.assign(nodes, nodes + n_nodes);
    if (lists && n_lists) p->lists.assign(lists, lists + n_lists);
    p->fixed_len = fixed_len;
    p->fast_sines = fast;
    if (std::getenv("TUUN_B200_DEBUG")) {
        std::fprintf(stderr, "[tuun_b200] lane_ok %u: W %u words, Q %u units, %u slots\n", p->low.lane_ok,
                     p->low.lane_w_words, p->low.lane_q_units, p->low.lane_slots);
        for (const tb_insn& i : p->low.lane_code)
            std::fprintf(stderr, "  op %3u q/class %3u np %2u hi %3u   a %d b %d c %d\n", i.op & 0xffu, (i.op >> 8) & 0xffu,
                         (i.op >> 16) & 0xffu, i.op >> 24, i.a, i.b, i.c);
        for (const tb_lane_aux& x : p->low.lane_aux)
            std::fprintf(stderr, "  aux kind %u a %d b %d c %d -> W %u Q %u\n", x.kind, x.a, x.b, x.c, x.w_off, x.q_off);
    }
    p->sample_rate = sample_rate;
    p->fast_mode = (fs && fs[0] == '1') ? 1u : 2u;  // "0" all exact, "1" f32 polynomial, default MUFU
    if (const char* se = std::getenv("TUUN_B200_STEADY"))
        if (se[0] == '0') {  // diagnostics: force the general interpreter (and with it the warp-per-voice kernel)
            p->low.steady_ok = 0;
            p->low.lane_ok = 0;
            p->low.lane_fin_goe = -1;
            p->low.lane_clk = 0;
        }
    // Everything below needs a device: no CPU path exists.
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        delete p;
        return set_error(TB_ERR_CUDA, std::string("no usable CUDA device: ") +
                                          (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    }
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= count) {
        delete p;
        return set_error(TB_ERR_INVALID, "device ordinal out of range");
    }
    p->device = device;
    auto bail = [&](int code) {
        delete p;
        return code;
    };
    if (cudaSetDevice(device) != cudaSuccess) return bail(set_error(TB_ERR_CUDA, "cudaSetDevice failed"));
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(set_error(TB_ERR_CUDA, "cudaStreamCreate failed"));
    p->own_stream = true;
    if (cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking) != cudaSuccess)
        return bail(set_error(TB_ERR_CUDA, "cudaStreamCreate failed"));
    for (int i = 0; i < 2; i++) {
        if (cudaEventCreateWithFlags(&p->ev_render[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&p->ev_copy[i], cudaEventDisableTiming) != cudaSuccess)
            return bail(set_error(TB_ERR_CUDA, "cudaEventCreate failed"));
    }
    std::vector<float> pool(fixed_pool, fixed_pool + fixed_len);
    if ((rc = upload(p->low.code, &p->d_code)) || (rc = upload(p->low.cexpr, &p->d_cexpr)) ||
        (rc = upload(p->low.aux, &p->d_aux)) || (rc = upload(p->low.goe, &p->d_goe)) ||
        (rc = upload(p->low.goe_steps, &p->d_goe_steps)) || (rc = upload(p->low.filt, &p->d_filt)) ||
        (rc = upload(p->low.fixed, &p->d_fixed)) || (rc = upload(pool, &p->d_pool)))
        return bail(rc);
    if (!size_cta(p->low, &p->warps, &p->smem))
        return bail(set_error(TB_ERR_UNSUPPORTED, "program needs more shared memory than one CTA has"));
    if (p->low.split_passes > 0 && (rc = upload(p->low.split, &p->d_split))) return bail(rc);
    // The lane-per-voice kernels (lanes.cuh): for batches large enough that one thread per voice fills
    // the device.  TUUN_B200_LANES=0 disables it; TUUN_B200_LANE_MIN_VOICES moves the threshold.
    const char* le = std::getenv("TUUN_B200_LANES");
    if (p->low.lane_ok && !(le && le[0] == '0')) {
        const size_t ls = tb_lanes_smem_bytes((uint32_t)p->low.lane_code.size(), p->low.lane_w_words,
                                              p->low.lane_q_units, p->low.lane_slots);
        int bps = 0, n_sm = 0;
        if (ls <= 220 * 1024 && tb_lanes_occupancy(ls, 0, &bps, &n_sm) == cudaSuccess && bps > 0) {
            if ((rc = upload(p->low.lane_code, &p->d_lane_code)) || (rc = upload(p->low.lane_aux, &p->d_lane_aux)))
                return bail(rc);
            if (cudaHostAlloc(reinterpret_cast<void**>(&p->h_fault), 4, cudaHostAllocMapped) == cudaSuccess) {
                *p->h_fault = 0u;
                if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&p->d_fault), p->h_fault, 0) != cudaSuccess)
                    p->d_fault = nullptr;
            }
            p->lane_smem = ls;
            p->lane_capacity = (uint32_t)(bps * n_sm);
            // A program that is one fused FM voice with a FAST carrier has its own kernel (lanes_fm.cu).
            const std::vector<tb_insn>& lc = p->low.lane_code;
            const char* fe = std::getenv("TUUN_B200_LANE_FM_KERNEL");  // diagnostics: "0" keeps it on the interpreter kernels
            if (lc.size() == 3 && (lc[0].op & 0xffu) == LN_FM && ((lc[0].op >> 16) & 0xffu) == 0 &&
                (lc[0].op >> 24) == TB_SINE_FAST && p->fast_mode == 2 && !(fe && fe[0] == '0')) {
                int fb = 0, fs = 0;
                if (tb_lanes_occupancy(ls, 2, &fb, &fs) == cudaSuccess && fb > 0) p->lane_fm_capacity = (uint32_t)(fb * fs);
                // Two warps a voice (lanes_fm_ws.cu): a filter tail; its ring lies over 8 Q units.
                const char* we = std::getenv("TUUN_B200_FM_WS");  // diagnostics: "0" keeps one thread a voice
                const size_t ws = tb_lanes_fm_ws_smem_bytes((uint32_t)lc.size(), p->low.lane_w_words, p->low.lane_q_units,
                                                            p->low.lane_slots);
                int wb = 0;
                if (p->lane_fm_capacity != 0 && lc[1].c >= 0 && p->low.filt.size() == 1 &&
                    p->low.lane_q_units + 4 * p->low.lane_slots >= 8 && ws <= 48 * 1024 && !(we && we[0] == '0') &&
                    tb_lanes_fm_ws_occupancy(ws, &wb) == cudaSuccess && wb > 0) {
                    p->lane_fm_ws_capacity = (uint32_t)(wb * fs);
                    p->lane_fm_ws_smem = ws;
                    int sb = 0;
                    if (tb_lanes_fm_ws_split_occupancy(ws, &sb) == cudaSuccess && sb > 0)
                        p->lane_fm_ws_split_capacity = (uint32_t)(sb * fs);
                    const char* wm = std::getenv("TUUN_B200_FM_WS_MIN_VOICES");
                    p->lane_fm_ws_min_voices = wm ? (uint32_t)std::strtoul(wm, nullptr, 10) : 0u;
                }
            }
            // Default threshold: a little over one CTA per SM.  Measured on config 5: the lane kernel takes
            // the same time for 9,472 and 18,944 voices (one warp per scheduler, latency bound: 3.7e11 and
            // 7.4e11 voice-samples/s) against 4.0e11 for the warp-per-voice kernel at any batch size.
            const char* mv = std::getenv("TUUN_B200_LANE_MIN_VOICES");
            p->lane_min_voices = mv ? (uint32_t)std::strtoul(mv, nullptr, 10)
                                    : (uint32_t)std::max(1, n_sm * TB_LANE_THREADS * 9 / 8);
            // Clocked words (a Reset or a timeline in the steady stream) run on the lane kernels only: the other
            // choice is the general interpreter.  Measured on config 2 (harmonica notes): level at 4,096 voices
            // (23.7 against 25.7 ms), the lane kernels ahead from there on (profiles/r2c_cfg2_lanes.txt).
            if (!mv && p->low.lane_clk) p->lane_min_voices = 4096;
        } else {
            cudaGetLastError();
        }
    }
    // A root sequence: one program per part (TUUN_B200_SEQ=0: the whole tree as one program, as for any other tree).
    const char* sq = std::getenv("TUUN_B200_SEQ");
    std::vector<tb::SeqPart> parts;
    if (allow_sequence && !(sq && sq[0] == '0') &&
        tb::sequence_parts(nodes, n_nodes, lists, n_lists, fixed_len, sample_rate, parts)) {
        bool ok = true;
        for (const tb::SeqPart& sp : parts) {
            std::vector<tb_node> sub;
            std::vector<int32_t> sub_lists;
            std::vector<uint32_t> ids;
            extract_subtree(p->nodes, p->lists, sp.root, sub, sub_lists, ids);
            tb_program* q = nullptr;
            if (create_program(sub.data(), (uint32_t)sub.size(), sub_lists.data(), (uint32_t)sub_lists.size(), fixed_pool,
                               fixed_len, sample_rate, device, &q, ids.data(), false) != TB_OK) {
                ok = false;
                break;
            }
            cudaStreamSynchronize(q->stream);
            if (q->own_stream) cudaStreamDestroy(q->stream);
            q->own_stream = false;
            const int si = (int)(p->part_prog.size() % tb_program::kSeqStreams);
            if (!p->seq_stream[si] &&
                (cudaStreamCreateWithFlags(&p->seq_stream[si], cudaStreamNonBlocking) != cudaSuccess ||
                 cudaEventCreateWithFlags(&p->seq_join[si], cudaEventDisableTiming) != cudaSuccess)) {
                delete q;
                ok = false;
                break;
            }
            q->stream = p->seq_stream[si];  // parts are independent streams: the ones a call overlaps render side by side
            p->part_prog.push_back(q);
            p->part_len.push_back(sp.len);
        }
        if (ok && cudaEventCreateWithFlags(&p->seq_fork, cudaEventDisableTiming) != cudaSuccess) ok = false;
        if (!ok) {  // some part the device path does not take: the whole tree as one program, like before
            for (tb_program* q : p->part_prog) delete q;
            p->part_prog.clear();
            p->part_len.clear();
            if (!p->seq_only) g_error.clear();
        }
    }
    if (p->seq_only && p->part_prog.empty()) {  // no whole-tree form and no parts: what the whole tree said
        delete p;
        return set_error(TB_ERR_UNSUPPORTED, whole_error);
    }
    *out_program = p;
    return TB_OK;
}

void tb_program_destroy(tb_program* p) { delete p; }

int tb_lower_check(const tb_node* nodes, uint32_t n_nodes, const int32_t* lists, uint32_t n_lists,
                   uint64_t fixed_len, tb_program_info* info) {
    if (!nodes || n_nodes == 0) return set_error(TB_ERR_INVALID, "empty op list");
    tb::Lowered low;
    const char* fs = std::getenv("TUUN_B200_FAST_SINES");
    // (lengths in samples — a timeline in the steady stream, lower.cpp — are decided at the default rate of the
    //  reference's tracker, 44,100 Hz: this entry point has no rate argument)
    int rc = tb::lower(nodes, n_nodes, lists, n_lists, fixed_len, !(fs && fs[0] == '0'), low, nullptr, 44100u);
    if (rc == TB_ERR_UNSUPPORTED) {
        // part by part (create_program): every part must lower; the geometry reported is that of the largest one
        std::vector<tb::SeqPart> parts;
        const char* sq0 = std::getenv("TUUN_B200_SEQ");
        std::string whole = low.error;
        if (!(sq0 && sq0[0] == '0') && tb::sequence_parts(nodes, n_nodes, lists, n_lists, fixed_len, 44100, parts)) {
            std::vector<tb_node> all(nodes, nodes + n_nodes);
            std::vector<int32_t> all_lists(lists, lists + (lists ? n_lists : 0));
            tb::Lowered best;
            bool ok = true;
            for (const tb::SeqPart& sp : parts) {
                std::vector<tb_node> sub;
                std::vector<int32_t> sub_lists;
                std::vector<uint32_t> ids;
                extract_subtree(all, all_lists, sp.root, sub, sub_lists, ids);
                tb::Lowered one;
                if (tb::lower(sub.data(), (uint32_t)sub.size(), sub_lists.data(), (uint32_t)sub_lists.size(), fixed_len,
                              !(fs && fs[0] == '0'), one, ids.data(), 44100u) != TB_OK) {
                    ok = false;
                    break;
                }
                if (one.code.size() >= best.code.size()) best = one;
            }
            if (ok) {
                low = best;
                low.n_nodes = n_nodes;
                rc = TB_OK;
            } else {
                low.error = whole;
            }
        }
    }
    if (rc != TB_OK) return set_error(rc, low.error);
    uint32_t warps = 0;
    size_t smem = 0;
    if (!size_cta(low, &warps, &smem))
        return set_error(TB_ERR_UNSUPPORTED, "program needs more shared memory than one CTA has");
    if (info) {
        info->n_nodes = low.n_nodes;
        info->n_code_words = (uint32_t)low.code.size() * 4;
        info->n_slots = low.n_slots;
        info->state_words = low.state_words;
        info->tile = low.steady_ok ? TB_TILE_S : TB_TILE;
        info->threads = 32 * warps;
        info->smem_bytes = (uint32_t)smem;
        info->n_params = low.n_params;
        info->kernel_launches = 0;
        info->lane_launches = 0;
        info->lane_smem_bytes = low.lane_ok ? (uint32_t)tb_lanes_smem_bytes((uint32_t)low.lane_code.size(), low.lane_w_words,
                                                                          low.lane_q_units, low.lane_slots)
                                            : 0u;
        info->lane_min_voices = 0;
        info->lane_capacity = 0;
        info->lane_fm_capacity = 0;
        info->lane_fm_ws_capacity = 0;
        info->fm_ws_launches = 0;
        info->split_passes = low.split_passes;
        info->split_segments = 0;
        info->split_seg_samples = 0;
        info->split_rounds = 0;
        std::vector<tb::SeqPart> parts;  // (lengths at 44.1 kHz: whether the root is a sequence does not depend on the rate)
        info->sequence_parts = tb::sequence_parts(nodes, n_nodes, lists, n_lists, fixed_len, 44100, parts) ? (uint32_t)parts.size() : 0u;
        info->split_fm_rounds = 0;
        info->sequence_renders = 0;
    }
    return TB_OK;
}

int tb_program_get_info(const tb_program* p, tb_program_info* info) {
    if (!p || !info) return set_error(TB_ERR_INVALID, "NULL argument");
    info->n_nodes = p->low.n_nodes;
    info->n_code_words = (uint32_t)p->low.code.size() * 4;
    info->n_slots = p->low.n_slots;
    info->state_words = p->low.state_words;
    info->tile = p->low.steady_ok ? TB_TILE_S : TB_TILE;
    info->threads = 32 * p->warps;
    info->smem_bytes = (uint32_t)p->smem;
    info->n_params = p->low.n_params;
    info->kernel_launches = p->launches;
    info->lane_launches = p->lane_launches;
    for (const tb_program* q : p->part_prog) {  // a sequence: what its parts launched
        info->kernel_launches += q->launches;
        info->lane_launches += q->lane_launches;
    }
    info->lane_smem_bytes = (uint32_t)p->lane_smem;
    info->lane_min_voices = p->lane_min_voices;
    info->lane_capacity = p->lane_capacity;
    info->lane_fm_capacity = p->lane_fm_capacity;
    info->lane_fm_ws_capacity = p->lane_fm_ws_capacity;
    info->fm_ws_launches = p->fm_ws_launches;
    info->split_passes = p->d_split ? p->low.split_passes : 0;
    info->split_segments = p->split_last_segments;
    info->split_seg_samples = p->split_last_seg_samples;
    info->split_rounds = p->split_rounds;
    info->sequence_parts = (uint32_t)p->part_prog.size();
    info->split_fm_rounds = (uint32_t)p->split_fm_rounds;
    info->sequence_renders = p->seq_renders;
    for (const tb_program* q : p->part_prog) info->split_rounds += q->split_rounds;
    return TB_OK;
}

void* tb_stream(tb_program* p) { return p ? (void*)p->stream : nullptr; }

int tb_set_stream(tb_program* p, void* cuda_stream) {
    if (!p) return set_error(TB_ERR_INVALID, "NULL program");
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->stream);
    if (p->own_stream) cudaStreamDestroy(p->stream);
    p->stream = (cudaStream_t)cuda_stream;
    p->own_stream = false;
    return TB_OK;  // (the parts of a sequence keep their own streams, forked from and joined to this one)
}

int tb_seed_noise(tb_program* p, uint64_t seed, uint64_t first_voice) {
    if (!p) return set_error(TB_ERR_INVALID, "NULL program");
    p->noise_seed = seed;
    p->noise_first_voice = first_voice;
    return TB_OK;
}

int tb_substitute(tb_program* p, uint32_t mark_id, float value, uint32_t* n_replaced) {
    if (!p) return set_error(TB_ERR_INVALID, "NULL program");
    if (n_replaced) *n_replaced = 0;
    std::vector<tb_node> nodes = p->nodes;
    uint32_t hits = 0;
    for (const tb_node& m : p->nodes) {
        if (m.kind != TB_MARKED || m.mark_id != mark_id) continue;
        if (m.a < 0 || (size_t)m.a >= nodes.size() || nodes[m.a].kind != TB_CONST)
            return set_error(TB_ERR_UNSUPPORTED, "tb_substitute: the Marked node holds a waveform that is not a Const");
        nodes[m.a].value = value;
        nodes[m.a].param_slot = -1;  // the new waveform is the same constant for every voice
        hits++;
    }
    if (n_replaced) *n_replaced = hits;
    if (hits == 0) return TB_OK;
    if (p->seq_only) {  // no whole-tree program to lower again: the parts hold the subtrees
        p->nodes.swap(nodes);
        for (tb_program* q : p->part_prog) {
            uint32_t n = 0;
            int rcq = tb_substitute(q, mark_id, value, &n);
            if (rcq) return rcq;
        }
        return TB_OK;
    }
    // Same tree shape, another literal: lower again and require the same program around the constant table
    // (a literal can decide the lowering only through is_const's Append(c, c) arm, generator.rs:597-603).
    tb::Lowered low;
    int rc = tb::lower(nodes.data(), (uint32_t)nodes.size(), p->lists.data(), (uint32_t)p->lists.size(), p->fixed_len,
                       p->fast_sines, low, p->noise_ids.empty() ? nullptr : p->noise_ids.data(), p->sample_rate);
    if (rc != TB_OK) return set_error(rc, low.error);
    const tb::Lowered& o = p->low;
    const bool same = low.code.size() == o.code.size() && low.cexpr.size() == o.cexpr.size() &&
                      low.state_words == o.state_words && low.aux.size() == o.aux.size() && low.n_slots == o.n_slots &&
                      low.lane_code.size() == o.lane_code.size() && low.lane_aux.size() == o.lane_aux.size() &&
                      low.split.size() == o.split.size() && low.goe.size() == o.goe.size() &&
                      low.goe_steps.size() == o.goe_steps.size() && low.filt.size() == o.filt.size() &&
                      (low.code.empty() || std::memcmp(low.code.data(), o.code.data(), low.code.size() * sizeof(tb_insn)) == 0) &&
                      (low.lane_code.empty() ||
                       std::memcmp(low.lane_code.data(), o.lane_code.data(), low.lane_code.size() * sizeof(tb_insn)) == 0) &&
                      (low.lane_aux.empty() ||
                       std::memcmp(low.lane_aux.data(), o.lane_aux.data(), low.lane_aux.size() * sizeof(tb_lane_aux)) == 0);
    if (!same) return set_error(TB_ERR_UNSUPPORTED, "tb_substitute: the new value changes how the tree lowers");
    CU(cudaSetDevice(p->device));
    // Renders enqueued so far read the old table: the copy is ordered behind them on the program's stream.
    CU(cudaStreamSynchronize(p->stream));
    if (!low.cexpr.empty())
        CU(cudaMemcpy(p->d_cexpr, low.cexpr.data(), low.cexpr.size() * sizeof(tb_cexpr), cudaMemcpyHostToDevice));
    // keep what tb_program_create adjusted after lowering (diagnostic switches)
    low.steady_ok = o.steady_ok;
    low.lane_ok = o.lane_ok;
    low.lane_fin_goe = o.lane_fin_goe;
    low.lane_clk = o.lane_clk;
    p->low = std::move(low);
    p->nodes.swap(nodes);
    for (tb_program* q : p->part_prog) {  // the parts of a sequence hold copies of their subtrees
        uint32_t n = 0;
        if ((rc = tb_substitute(q, mark_id, value, &n))) return rc;
    }
    return TB_OK;
}

// ---- time-segment sharding across GPUs (include/tuun_b200.h) ------------------------------------------
int tb_segments_begin(tb_program* p, const float* params, uint32_t n_params, uint32_t n_voices, uint32_t n_segments,
                      uint64_t seg_samples, uint32_t flags, uint32_t* n_passes) {
    if (!p) return set_error(TB_ERR_INVALID, "NULL program");
    if (p->low.split_passes == 0 || !p->d_split)
        return set_error(TB_ERR_UNSUPPORTED, "tb_segments_begin: not a steady program (state has no associative form)");
    if (n_voices == 0 || n_segments < 1 || seg_samples == 0 || seg_samples % TB_TILE_S != 0)
        return set_error(TB_ERR_INVALID, "tb_segments_begin: seg_samples must be a positive multiple of 512");
    if ((uint64_t)n_voices * n_segments > 0x7fffffffull) return set_error(TB_ERR_INVALID, "too many segments");
    CU(cudaSetDevice(p->device));
    int rc = ensure_voices(p, n_voices);
    if (rc) return rc;
    if (!p->pos_known || (!p->low.filt.empty() && p->stream_pos < (uint64_t)TB_TILE))
        return set_error(TB_ERR_STATE, "tb_segments_begin: render the first 256 samples of the stream with tb_render first "
                                       "(filters read ahead on their first call, generator.rs:234-252)");
    const float* d_params = nullptr;
    if ((rc = stage_params(p, params, n_params, n_voices, flags, &d_params))) return rc;
    tb_launch L;
    fill_launch(p, &L);
    L.params = d_params;
    L.n_params = n_params;
    L.n_voices = n_voices;
    SplitPlan plan;
    plan.n_seg = n_segments;
    plan.seg = seg_samples;
    if ((rc = split_begin(p, L, plan))) return rc;
    // A batch of fused FM voices with a biquad takes the cheaper form of render_split_fm: pass 1 = phase sums only,
    // pass 2 = the filters' warm-ups (local to every rank), pass 3 = the samples; same protocol for the caller.
    p->seg_fm = false;
    const char* fe = std::getenv("TUUN_B200_SPLIT_FM");
    if (p->low.split_passes == 3 && p->lane_fm_capacity != 0 && p->low.filt.size() == 1 && !(fe && fe[0] == '0') &&
        !std::getenv("TUUN_B200_SPLIT") && ((fe && fe[0] != '0') || (uint64_t)n_voices * n_segments >= 16384)) {
        bool ok = false;
        if ((rc = prepare_split_fm(p, seg_samples, &ok))) return rc;
        p->seg_fm = ok;
    }
    p->seg_voices = n_voices;
    p->seg_params = d_params;
    p->seg_n_params = n_params;
    if (n_passes) *n_passes = p->low.split_passes;
    return TB_OK;
}

int tb_segments_pass(tb_program* p, uint32_t pass, uint32_t seg_lo, uint32_t seg_hi, float* out, uint64_t out_stride,
                     uint32_t flags) {
    if (!p || p->seg_voices == 0) return set_error(TB_ERR_STATE, "tb_segments_pass without tb_segments_begin");
    const tb_split_args& A = p->split_args;
    if (pass < 1 || pass > p->low.split_passes || seg_lo >= seg_hi || seg_hi > A.n_seg)
        return set_error(TB_ERR_INVALID, "tb_segments_pass: bad pass or segment range");
    const bool last = pass == p->low.split_passes;
    if (last && (!out || !(flags & TB_OUT_DEVICE)))
        return set_error(TB_ERR_INVALID, "tb_segments_pass: the last pass needs device rows (TB_OUT_DEVICE)");
    if (last && out_stride < (uint64_t)(seg_hi - seg_lo) * A.seg) return set_error(TB_ERR_INVALID, "out_stride too small");
    CU(cudaSetDevice(p->device));
    tb_launch L;
    fill_launch(p, &L);
    L.params = p->seg_params;
    L.n_params = p->seg_n_params;
    L.n_voices = p->seg_voices;
    L.out_stride = out_stride;
    // rows are addressed by the segment's number within the voice: segment seg_lo starts at out[v * stride]
    float* base = last ? out - (size_t)seg_lo * A.seg : nullptr;
    if (p->seg_fm && pass == 1) {  // phase sums (the snapshot rides in the state blocks: lanes.cuh run_fm_sums)
        PassOpt sums;
        sums.fm_sums = true;
        sums.snap_at = A.seg - A.warm;
        return split_pass(p, L, 1, seg_lo, seg_hi, nullptr, p->stream_pos, sums);
    }
    if (p->seg_fm && pass == 2) {  // the warm-ups of this rank's segments, from the states tb_segments_fix(1) made
        PassOpt wp;
        wp.states = p->d_vw;
        wp.samples = A.warm;
        return split_pass(p, L, 2, seg_lo, seg_hi, nullptr, p->stream_pos, wp);
    }
    return split_pass(p, L, pass, seg_lo, seg_hi, base, p->stream_pos);
}

int tb_segments_states(tb_program* p, void** states, uint64_t* bytes_per_segment) {
    if (!p || p->seg_voices == 0) return set_error(TB_ERR_STATE, "tb_segments_states without tb_segments_begin");
    if (states) *states = p->d_vs;
    if (bytes_per_segment) *bytes_per_segment = (uint64_t)p->low.state_words * 4;
    return TB_OK;
}

int tb_segments_fix(tb_program* p, uint32_t pass) {
    if (!p || p->seg_voices == 0) return set_error(TB_ERR_STATE, "tb_segments_fix without tb_segments_begin");
    if (pass < 1 || pass >= p->low.split_passes) return set_error(TB_ERR_INVALID, "tb_segments_fix: no such summary pass");
    CU(cudaSetDevice(p->device));
    if (p->seg_fm) {
        cudaError_t e = cudaSuccess;
        if (pass == 1) {  // exact carrier starts, then the states every warm-up starts from
            int rc = split_fix(p, 1);
            if (rc) return rc;
            e = tb_split_fm_warm_seed(&p->split_args, p->stream);
        } else {          // the warm-ups' final filter histories are the segments' initial ones
            e = tb_split_fm_adopt(&p->split_args, p->stream);
        }
        if (e != cudaSuccess) return cuda_fail(e, "tb_segments_fix");
        p->launches++;
        return TB_OK;
    }
    return split_fix(p, pass);
}

int tb_segments_end(tb_program* p) {
    if (!p || p->seg_voices == 0) return set_error(TB_ERR_STATE, "tb_segments_end without tb_segments_begin");
    CU(cudaSetDevice(p->device));
    const uint64_t n = p->split_args.seg * p->split_args.n_seg;
    int rc = split_finish(p, p->d_state, nullptr, n, false);
    if (rc) return rc;
    p->stream_pos += n;
    p->seg_voices = 0;
    p->split_rounds++;
    if (p->seg_fm) p->split_fm_rounds++;
    p->split_last_segments = p->split_args.n_seg;
    p->split_last_seg_samples = p->split_args.seg;
    return TB_OK;
}

int tb_reset(tb_program* p) {
    if (!p) return set_error(TB_ERR_INVALID, "NULL program");
    CU(cudaSetDevice(p->device));
    // Every node's Initial state is the all-zero block, so a reset is one memset on the stream.
    if (p->d_state)
        CU(cudaMemsetAsync(p->d_state, 0, (size_t)p->n_voices * p->low.state_words * 4, p->stream));
    p->fresh = true;
    p->stream_pos = 0;
    p->pos_known = true;
    for (tb_program* q : p->part_prog) {
        int rc = tb_reset(q);
        if (rc) return rc;
    }
    return TB_OK;
}

}  // extern "C"

namespace {

// Shared body of tb_render / tb_render_mix.
//  * device rows:   one launch over all voices straight into `out`.
//  * host rows / no rows: voices are rendered group by group into two device staging buffers;
//    a group's rows are contiguous, so each group leaves with ONE cudaMemcpyAsync on a second
//    stream (2-D copies run at a few GB/s on this platform, 1-D ones at ~56 GB/s) while the next
//    group renders.  Very long renders are additionally cut in time; state carries over.
//  * mix: after a group (or the whole batch) is rendered, tb_mix_kernel adds its rows in voice
//    order — the tracker's serial `out[j] += tmp[j]` (tracker.rs:617-619).
int render_impl(tb_program* p, const float* params, uint32_t n_params, uint32_t n_voices, uint64_t n_samples,
                float* out, uint64_t out_stride, uint64_t* out_len, float* mix, bool want_mix, uint32_t flags) {
    if (!p) return set_error(TB_ERR_INVALID, "NULL program");
    if (n_voices == 0) return TB_OK;
    const bool dev_out = (flags & TB_OUT_DEVICE) != 0;
    const bool no_rows = want_mix && (flags & TB_NO_VOICE_OUT);
    if (!no_rows && !out) return set_error(TB_ERR_INVALID, "out is NULL");
    if (!no_rows && out_stride < n_samples) return set_error(TB_ERR_INVALID, "out_stride < n_samples");
    if (want_mix && !mix) return set_error(TB_ERR_INVALID, "mix is NULL");
    CU(cudaSetDevice(p->device));
    int rc = ensure_voices(p, n_voices);
    if (rc) return rc;
    const float* d_params = nullptr;
    if ((rc = stage_params(p, params, n_params, n_voices, flags, &d_params))) return rc;
    if (n_samples == 0) {
        if (out_len) std::fill(out_len, out_len + n_voices, 0ull);
        return TB_OK;
    }
    tb_launch L;
    fill_launch(p, &L);
    L.n_params = n_params;
    L.mode = 0;
    float* d_mix = nullptr;
    if (want_mix) {
        if (dev_out) d_mix = mix;
        else {
            if (n_samples > p->mix_cap) {
                cudaFree(p->d_mix);
                p->d_mix = nullptr;
                p->mix_cap = 0;
                CU(cudaMalloc(reinterpret_cast<void**>(&p->d_mix), n_samples * 4));
                p->mix_cap = n_samples;
            }
            d_mix = p->d_mix;
        }
    }
    const bool primed = p->pos_known && p->stream_pos >= (uint64_t)TB_TILE;
    if (no_rows && p->lane_smem != 0 && p->low.lane_fin_goe < 0 && n_voices >= p->lane_min_voices &&
        n_samples >= (primed ? 0 : (uint64_t)TB_TILE) + 2 * TB_LS) {
        L.params = d_params;
        L.n_voices = n_voices;
        L.n_samples = n_samples;
        L.out_len = p->d_len;
        if ((rc = render_mix_lanes(p, L, d_mix, p->stream_pos))) return rc;
    } else if (dev_out && !no_rows) {
        L.params = d_params;
        L.n_voices = n_voices;
        L.n_samples = n_samples;
        L.out = out;
        L.out_stride = out_stride;
        L.out_len = p->d_len;
        if ((rc = launch_generate(p, L, p->stream_pos))) return rc;
        if (want_mix) {
            cudaError_t e = tb_mix_launch(out, out_stride, p->d_len, n_voices, n_samples, 0, d_mix, 0, p->stream);
            if (e != cudaSuccess) return cuda_fail(e, "tb_mix_kernel launch");
            p->launches++;
        }
    } else {
        const char* env = std::getenv("TUUN_B200_STAGE_MB");
        uint64_t budget = (env ? std::strtoull(env, nullptr, 10) : 2048ull) << 18;  // floats per buffer
        if (budget < (uint64_t)TB_TILE) budget = TB_TILE;
        uint64_t t_chunk = n_samples;
        if (t_chunk > budget) t_chunk = budget / TB_TILE * TB_TILE;
        uint64_t G = budget / t_chunk;
        if (G < 1) G = 1;
        if (G > n_voices) G = n_voices;
        const size_t need = (size_t)G * t_chunk;
        if (need > p->stage_cap) {
            for (int i = 0; i < 2; i++) {
                cudaFree(p->d_stage[i]);
                p->d_stage[i] = nullptr;
            }
            p->stage_cap = 0;
            for (int i = 0; i < 2; i++) CU(cudaMalloc(reinterpret_cast<void**>(&p->d_stage[i]), need * 4));
            p->stage_cap = need;
        }
        const bool cut_time = t_chunk < n_samples;
        if (cut_time) CU(cudaMemsetAsync(p->d_done, 0, n_voices, p->stream));
        int k = 0;
        for (uint64_t v0 = 0; v0 < n_voices; v0 += G) {
            const uint32_t g = (uint32_t)std::min<uint64_t>(G, n_voices - v0);
            for (uint64_t t0 = 0; t0 < n_samples; t0 += t_chunk, k++) {
                const uint64_t len = std::min<uint64_t>(t_chunk, n_samples - t0);
                const int b = k & 1;
                if (k >= 2) CU(cudaStreamWaitEvent(p->stream, p->ev_copy[b], 0));
                L.params = d_params ? d_params + (size_t)v0 * n_params : nullptr;
                L.voice_base = p->noise_first_voice + v0;
                L.state = p->d_state + (size_t)v0 * p->low.state_words;
                L.out_len = p->d_len + v0;
                L.done = cut_time ? p->d_done + v0 : nullptr;
                L.accumulate = t0 > 0;
                L.mid_call = t0 > 0;
                L.call_pos = t0;
                L.n_voices = g;
                L.n_samples = len;
                L.out = p->d_stage[b];
                L.out_stride = len;
                if ((rc = launch_generate(p, L, p->stream_pos + t0))) return rc;
                if (want_mix) {
                    cudaError_t e = tb_mix_launch(p->d_stage[b], len, p->d_len + v0, g, len, t0, d_mix + t0,
                                                  v0 > 0 ? 1 : 0, p->stream);
                    if (e != cudaSuccess) return cuda_fail(e, "tb_mix_kernel launch");
                    p->launches++;
                }
                CU(cudaEventRecord(p->ev_render[b], p->stream));
                if (!no_rows) {
                    CU(cudaStreamWaitEvent(p->copy_stream, p->ev_render[b], 0));
                    if (!cut_time && out_stride == n_samples) {
                        CU(cudaMemcpyAsync(out + (size_t)v0 * out_stride, p->d_stage[b], (size_t)g * len * 4,
                                           cudaMemcpyDeviceToHost, p->copy_stream));
                    } else {
                        for (uint32_t r = 0; r < g; r++)
                            CU(cudaMemcpyAsync(out + (size_t)(v0 + r) * out_stride + t0, p->d_stage[b] + (size_t)r * len,
                                               len * 4, cudaMemcpyDeviceToHost, p->copy_stream));
                    }
                    CU(cudaEventRecord(p->ev_copy[b], p->copy_stream));
                } else {
                    CU(cudaEventRecord(p->ev_copy[b], p->stream));
                }
            }
        }
        if (!no_rows) CU(cudaStreamSynchronize(p->copy_stream));
    }
    p->stream_pos += n_samples;
    if (want_mix && !dev_out)
        CU(cudaMemcpyAsync(mix, d_mix, n_samples * 4, cudaMemcpyDeviceToHost, p->stream));
    if (out_len) {
        CU(cudaMemcpyAsync(out_len, p->d_len, (size_t)n_voices * 8, cudaMemcpyDeviceToHost, p->stream));
        CU(cudaStreamSynchronize(p->stream));
        return check_fault(p);
    } else if (!dev_out) {
        CU(cudaStreamSynchronize(p->stream));
        return check_fault(p);
    }
    return TB_OK;
}

}  // namespace

extern "C" {

int tb_render(tb_program* p, const float* params, uint32_t n_params, uint32_t n_voices, uint64_t n_samples,
              float* out, uint64_t out_stride, uint64_t* out_len, uint32_t flags) {
    return render_impl(p, params, n_params, n_voices, n_samples, out, out_stride, out_len, nullptr, false,
                       flags & ~TB_NO_VOICE_OUT);
}

int tb_render_mix(tb_program* p, const float* params, uint32_t n_params, uint32_t n_voices, uint64_t n_samples,
                  float* out, uint64_t out_stride, uint64_t* out_len, float* mix, uint32_t flags) {
    return render_impl(p, params, n_params, n_voices, n_samples, out, out_stride, out_len, mix, true, flags);
}

int tb_lane_kernel_times(tb_program* p, float* ms, uint32_t cap, uint32_t* n) {
    if (!p || !n) return set_error(TB_ERR_INVALID, "NULL argument");
    CU(cudaSetDevice(p->device));
    CU(cudaStreamSynchronize(p->stream));
    const uint64_t have = std::min<uint64_t>(p->lane_launches, tb_program::kLaneEvents);
    const uint32_t take = (uint32_t)std::min<uint64_t>(have, cap);
    for (uint32_t i = 0; i < take; i++) {  // oldest of the last `take` first
        const uint64_t k = p->lane_launches - take + i;
        cudaEvent_t* ev = p->lane_ev[k % tb_program::kLaneEvents];
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, ev[0], ev[1]));
        if (ms) ms[i] = t;
    }
    *n = take;
    return TB_OK;
}

int tb_length(tb_program* p, const float* params, uint32_t n_params, uint32_t n_voices, uint64_t max,
              uint64_t* len, uint32_t flags) {
    if (!p) return set_error(TB_ERR_INVALID, "NULL program");
    if (n_voices == 0) return TB_OK;
    if (!len) return set_error(TB_ERR_INVALID, "len is NULL");
    CU(cudaSetDevice(p->device));
    int rc = ensure_voices(p, n_voices);
    if (rc) return rc;
    const float* d_params = nullptr;
    if ((rc = stage_params(p, params, n_params, n_voices, flags, &d_params))) return rc;
    tb_launch L;
    fill_launch(p, &L);
    L.params = d_params;
    L.n_params = n_params;
    L.n_voices = n_voices;
    L.n_samples = max;
    L.out_len = p->d_len;
    L.mode = 1;
    if (!p->part_prog.empty() && p->pos_known) {
        // A root sequence (launch_sequence): length(Append(a, b), max) advances a, then b by what is left
        // (generator.rs:705-722) — part by part, like generate.
        uint64_t start = 0;
        const uint64_t pos = p->stream_pos;
        bool first = true;
        for (size_t k = 0; k < p->part_prog.size(); k++) {
            const uint64_t plen = p->part_len[k];
            const uint64_t end = plen == ~0ull ? ~0ull : start + plen;
            const uint64_t a = std::max(pos, start), b = std::min(pos + max, end);
            if (a < b) {
                tb_program* q = p->part_prog[k];
                if ((rc = ensure_voices(q, n_voices))) return rc;
                tb_launch Q;
                fill_launch(q, &Q);
                Q.params = d_params;
                Q.n_params = n_params;
                Q.n_voices = n_voices;
                Q.n_samples = b - a;
                Q.out_len = p->d_len;
                Q.mode = 1;
                Q.accumulate = first ? 0u : 1u;
                q->pos_known = false;
                // (on the part's own stream, ordered between what the program's stream did before and does next)
                CU(cudaEventRecord(p->seq_fork, p->stream));
                CU(cudaStreamWaitEvent(q->stream, p->seq_fork, 0));
                if ((rc = launch(q, Q))) return rc;
                const int si = (int)(k % tb_program::kSeqStreams);
                CU(cudaEventRecord(p->seq_join[si], q->stream));
                CU(cudaStreamWaitEvent(p->stream, p->seq_join[si], 0));
                first = false;
            }
            if (end == ~0ull || end >= pos + max) break;
            start = end;
        }
        p->stream_pos += max;
        CU(cudaMemcpyAsync(len, p->d_len, (size_t)n_voices * 8, cudaMemcpyDeviceToHost, p->stream));
        CU(cudaStreamSynchronize(p->stream));
        return TB_OK;
    }
    p->pos_known = false;  // length() advances positions by per-node amounts (generator.rs:620-782)
    if ((rc = launch(p, L))) return rc;
    CU(cudaMemcpyAsync(len, p->d_len, (size_t)n_voices * 8, cudaMemcpyDeviceToHost, p->stream));
    CU(cudaStreamSynchronize(p->stream));
    return TB_OK;
}

}  // extern "C"
