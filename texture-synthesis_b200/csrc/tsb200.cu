// tsb200.cu -- host side of the C ABI in include/tsb200.h: device memory, the exact wave schedule
// (stage plan, pixel order, dependency phases, rounds) and the read-outs.
//
// Reference call stack being replaced: Session::run (lib/src/session.rs:37-66) ->
// Generator::resolve_random_batch (ms.rs:427-445) + Generator::resolve (ms.rs:702-1052).
#include "../../include/tsb200.h"

#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "tsb_device.cuh"
#include "tsb_stream.cuh"

using namespace tsb;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return fail(TSB_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)
#define TRY(call)                 \
    do {                          \
        int r__ = (call);         \
        if (r__ != 0) return r__; \
    } while (0)

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    int ensure(size_t count) {
        if (count <= n) return 0;
        release();
        CU(cudaMalloc((void**)&p, std::max<size_t>(count, 1) * sizeof(T)));
        n = count;
        return 0;
    }
    int upload(const T* h, size_t count, cudaStream_t s) {
        TRY(ensure(count));
        if (count) CU(cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
        return 0;
    }
};

template <typename T>
struct PinnedBuf {
    T* p = nullptr;
    size_t n = 0;
    ~PinnedBuf() { if (p) cudaFreeHost(p); }
    int ensure(size_t count) {
        if (count <= n) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; n = 0;
        CU(cudaMallocHost((void**)&p, std::max<size_t>(count, 1) * sizeof(T)));
        n = count;
        return 0;
    }
};

// ---------------------------------------------------------------------------------------------
// image 0.23.12 imageops::resize tap tables (vertical_sample / horizontal_sample share the formula).
// ---------------------------------------------------------------------------------------------
constexpr int FILTER_BLUR = 3;  // image 0.23.12 `blur(sigma)`: kernel gaussian(x, sigma), support 2*sigma (utils.rs:115)

float kernel_eval(int f, float x, float sigma = 0.5f) {
    switch (f) {
    case FILTER_BLUR: {
        float norm = 1.0f / (sqrtf(2.0f * 3.14159265358979323846f) * sigma);
        return norm * expf(-(x * x) / (2.0f * (sigma * sigma)));
    }
    case TSB_FILTER_TRIANGLE: {
        float a = fabsf(x);
        return a < 1.0f ? 1.0f - a : 0.0f;
    }
    case TSB_FILTER_CATMULLROM: {  // bc_cubic_spline(x, b = 0, c = 0.5)
        const float b = 0.0f, c = 0.5f;
        float a = fabsf(x), k;
        if (a < 1.0f) k = (12.0f - 9.0f * b - 6.0f * c) * (a * a * a) + (-18.0f + 12.0f * b + 6.0f * c) * (a * a) + (6.0f - 2.0f * b);
        else if (a < 2.0f) k = (-b - 6.0f * c) * (a * a * a) + (6.0f * b + 30.0f * c) * (a * a) + (-12.0f * b - 48.0f * c) * a + (8.0f * b + 24.0f * c);
        else k = 0.0f;
        return k / 6.0f;
    }
    default: {  // gaussian(x, r = 0.5)
        const float r = 0.5f;
        float norm = 1.0f / (sqrtf(2.0f * 3.14159265358979323846f) * r);
        return norm * expf(-(x * x) / (2.0f * (r * r)));
    }
    }
}

struct HostTaps {
    std::vector<int> left, count, offset;
    std::vector<float> sum, weights;
};

void build_taps(int in_sz, int out_sz, int filter, HostTaps& t, float sigma = 0.5f) {
    t = HostTaps();
    const float support0 = filter == FILTER_BLUR ? 2.0f * sigma : (filter == TSB_FILTER_TRIANGLE ? 1.0f : (filter == TSB_FILTER_CATMULLROM ? 2.0f : 3.0f));
    const float ratio = (float)in_sz / (float)out_sz;
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float support = support0 * sratio;
    for (int o = 0; o < out_sz; ++o) {
        float inputx = ((float)o + 0.5f) * ratio;
        long long left = (long long)floorf(inputx - support);
        left = std::min<long long>(std::max<long long>(left, 0), (long long)in_sz - 1);
        long long right = (long long)ceilf(inputx + support);
        right = std::min<long long>(std::max<long long>(right, left + 1), (long long)in_sz);
        inputx = inputx - 0.5f;
        t.left.push_back((int)left);
        t.count.push_back((int)(right - left));
        t.offset.push_back((int)t.weights.size());
        float sum = 0.0f;
        for (long long i = left; i < right; ++i) {
            float w = kernel_eval(filter, ((float)i - inputx) / sratio, sigma);
            t.weights.push_back(w);
            sum += w;
        }
        t.sum.push_back(sum);
    }
}

struct DevTaps {
    DevBuf<int> left, count, offset;
    DevBuf<float> sum, weights;
    TapTable table() const { return TapTable{left.p, count.p, offset.p, sum.p, weights.p}; }
    int upload(const HostTaps& h, cudaStream_t s) {
        TRY(left.upload(h.left.data(), h.left.size(), s));
        TRY(count.upload(h.count.data(), h.count.size(), s));
        TRY(offset.upload(h.offset.data(), h.offset.size(), s));
        TRY(sum.upload(h.sum.data(), h.sum.size(), s));
        TRY(weights.upload(h.weights.data(), h.weights.size(), s));
        return 0;
    }
};

// resize on device buffers: src (w x h) -> dst (nw x nh); tmp must hold w*nh pixels.
int device_resize(const uint32_t* src, int w, int h, uint32_t* dst, int nw, int nh, int filter, uint32_t* tmp, cudaStream_t s, float sigma = 0.5f) {
    if (nw <= 0 || nh <= 0) return 0;
    HostTaps hv, hh;
    build_taps(h, nh, filter, hv, sigma);
    build_taps(w, nw, filter, hh, sigma);
    DevTaps dv, dh;
    TRY(dv.upload(hv, s));
    TRY(dh.upload(hh, s));
    dim3 b(128, 1);
    k_resample_v<<<dim3((w + 127) / 128, nh), b, 0, s>>>(src, w, h, tmp, nh, dv.table());
    k_resample_h<<<dim3((nw + 127) / 128, nh), b, 0, s>>>(tmp, w, nh, dst, nw, dh.table());
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s));  // tap tables are freed on return
    return 0;
}

struct SpiralHost {
    std::vector<short2> off;
    std::vector<uint32_t> cntLE;
    int RT2;
};

SpiralHost build_spiral(int RT) {
    SpiralHost sp;
    sp.RT2 = RT * RT;
    struct E { int d2, dy, dx; };
    std::vector<E> v;
    for (int dy = -RT; dy <= RT; ++dy)
        for (int dx = -RT; dx <= RT; ++dx) {
            int d2 = dx * dx + dy * dy;
            if (d2 <= sp.RT2) v.push_back({d2, dy, dx});
        }
    std::sort(v.begin(), v.end(), [](const E& a, const E& b) {
        if (a.d2 != b.d2) return a.d2 < b.d2;
        if (a.dy != b.dy) return a.dy < b.dy;
        return a.dx < b.dx;
    });
    sp.cntLE.assign(sp.RT2 + 1, 0);
    for (auto& e : v) {
        sp.off.push_back(make_short2((short)e.dx, (short)e.dy));
        sp.cntLE[e.d2] += 1;
    }
    for (int i = 1; i <= sp.RT2; ++i) sp.cntLE[i] += sp.cntLE[i - 1];
    return sp;
}

constexpr int SPIRAL_RT = 48;

struct StagePlan {
    int p_stage, level;
    bool recolour;
    uint64_t seed;
    size_t pixels_to_resolve, redo_count, n_redo, n_new, pick_base, resolved_before;
    float adaptive_alpha;
};

}  // namespace

// CUDA events are created once and reused by every run of a generator
struct EventPool {
    std::vector<cudaEvent_t> timing, plain;
    size_t nt = 0, np = 0;
    ~EventPool() { for (auto e : timing) cudaEventDestroy(e); for (auto e : plain) cudaEventDestroy(e); }
    void rewind() { nt = np = 0; }
    int get(cudaEvent_t* out, bool with_timing) {
        std::vector<cudaEvent_t>& v = with_timing ? timing : plain;
        size_t& n = with_timing ? nt : np;
        if (n == v.size()) {
            cudaEvent_t e;
            if (cudaEventCreateWithFlags(&e, with_timing ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess) return TSB_ERR_CUDA;
            v.push_back(e);
        }
        *out = v[n++];
        return 0;
    }
};

struct tsb_generator {
    int device = 0;
    cudaStream_t stream = nullptr;
    int W = 0, H = 0;
    bool inpaint = false;
    uint32_t inpaint_index = 0;
    // host mirrors of the order-defining state (ms.rs:212-215)
    std::vector<uint32_t> unresolved0;       // as created
    std::vector<uint32_t> resolved0;         // locked inpaint pixels, as created
    std::vector<uint32_t> unresolved;        // current
    std::vector<uint32_t> resolved_order;    // `resolved` flat coords in resolution order
    size_t locked = 0, inpaint_locked = 0;
    std::vector<int32_t> loaded_points;      // load_state: explicit tree points
    bool have_loaded_points = false;

    // device state
    DevBuf<uint4> d_state;
    DevBuf<float> d_score;
    DevBuf<uint32_t> d_mask, d_mask1;
    DevBuf<uint32_t> d_pmask, d_pmask1;            // the current stage's new pixels (+ mirror copies), geometry of d_mask
    DevBuf<uint32_t> d_inp_mask, d_inp_color;
    int mx = 0, my = 0, wpr = 0, mrows = 0, wpr1 = 0;
    DevBuf<short2> d_spiral;
    DevBuf<uint32_t> d_cntLE;
    DevBuf<double> d_divx, d_divy;                  // (double)(i - mx) / W and (double)(i - my) / H over the mask extent
    int spiralN = 0, RT2 = 0;

    // inputs (device resident)
    bool inputs_ready = false;
    int n_levels = 0, n_ex_all = 0, n_ex = 0;
    std::vector<int> ex_w, ex_h, ex_kind;           // all examples
    std::vector<int> filt;                          // filtered index -> all index
    std::vector<DevBuf<uint32_t>> d_ex;             // per (all) example: levels*w*h
    std::vector<DevBuf<uint32_t>> d_exf, d_exgf;    // framed copies (EX_PAD texels of the outside colour all round)
    DevBuf<uint32_t> d_alpha_flag;                  // [level] set when an input texel of that pyramid level has alpha != 255, [64] same for the state
    std::vector<uint8_t> level_opaque;              // per pyramid level: every input texel has alpha 255
    int pad_pitch = 0;                              // common row pitch of the framed copies, 0 when the sizes differ
    bool run_opaque = false, no_fast = false, luts_exact_mode = false;
    std::vector<DevBuf<uint8_t>> d_smask;           // per (all) example
    DevBuf<DevEx> d_exdesc;                         // [levels][n_ex] filtered
    bool guided = false;
    int tgw = 0, tgh = 0;
    DevBuf<uint32_t> d_tguide;                      // levels*tgw*tgh
    std::vector<DevBuf<uint32_t>> d_exg;
    std::vector<int> exg_w, exg_h;
    DevBuf<DevGuide> d_exgdesc;                     // [levels][n_ex_all]
    DevBuf<float> d_luts;                           // 512 floats
    DevBuf<unsigned long long> d_counters;

    // per-run buffers
    DevBuf<uint32_t> d_item_pixel, d_pick_idx, d_tmp_u32, d_read_color, d_read_coord, d_read_id;
    DevBuf<uint8_t> d_cub_temp, d_sort_temp;
    DevBuf<unsigned long long> d_keys, d_keys_sorted;
    DevBuf<uint32_t> d_v0, d_pick_first;
    cudaStream_t stream2 = nullptr;
    int max_ctas_radius = 0, max_ctas_stream = 0, max_ctas_stream_guided = 0;
    std::vector<void*> mg_opened;
    std::vector<std::pair<cudaIpcMemHandle_t, void*>> mg_blocks;
    uint32_t* h_ctrl = nullptr;  // pinned, 16 words
    PinnedBuf<uint32_t> h_items;  // host mirror of the pick array

    // in-order streaming scheduler (tsb_stream.cuh)
    DevBuf<uint4> d_state2;                         // second state buffer (redo phases write here, then the two swap)
    DevBuf<uint32_t> d_tmap;                        // pixel -> position in the pick array
    DevBuf<uint8_t> r_nbk, r_rand_map;              // ring of per-item lists written by the analysis stream
    DevBuf<short2> r_nb;
    DevBuf<float> r_g;
    DevBuf<uint4> r_low;
    DevBuf<uint32_t> r_rand_xy;
    DevBuf<float> d_luts_all;                       // [stage][512]
    PinnedBuf<float> h_luts_all;
    DevBuf<uint32_t> d_sctl;                        // [chunk][SC_WORDS] + abort flag
    EventPool events;
    DevBuf<uint32_t> d_live_color;                  // colour plane mirrored for progress snapshots
    PinnedBuf<uint32_t> h_snap;
    cudaStream_t stream3 = nullptr;                 // host copies that must not delay the analysis stream
    uint32_t* h_progress = nullptr;                 // mapped pinned word: work items claimed so far
    uint32_t* d_progress = nullptr;                 // its device alias
    // band-sharded execution on the streaming scheduler: one process per GPU, replicas linked through CUDA IPC
    bool mgs_on = false;
    int mgs_rank = 0, mgs_world = 1, mgs_band_h = 0;
    uint4* mgs_A[MG_MAX] = {nullptr};               // every rank's first / second state buffer (peer-mapped)
    uint4* mgs_B[MG_MAX] = {nullptr};
    float* mgs_score[MG_MAX] = {nullptr};
    uint32_t* mgs_sync[MG_MAX] = {nullptr};         // every rank's barrier flag block
    DevBuf<uint32_t> d_mgs_sync;                    // [MG_MAX] flags written by the peers
    DevBuf<uint32_t*> d_mgs_sync_ptrs;
    uint32_t mgs_seq = 0;                           // barrier sequence number (identical on all ranks)
    size_t mgs_shard_min = 32768;                   // smaller phases are executed redundantly by every rank
    DevBuf<uint8_t> d_own_flag;
    DevBuf<uint32_t> d_own_pos, d_own_t, d_own_pix, d_own_cnt, d_own_bidx, d_own_bpos, d_init_points;
    uint64_t mgs_sharded_chunks = 0;
    size_t l2_persist_bytes = 0;                    // persisting L2 carve-out set aside for the active example level (0: unavailable)
    bool state_init_opaque = true;                  // every colour in the state before the run has alpha 255
    bool inpaint_opaque = true;                     // the same for the locked inpaint pixels as created

    // trace
    bool trace = false;
    DevBuf<int32_t> d_tr_best, d_tr_ncand, d_tr_nneigh;
    DevBuf<float> d_tr_score;
    std::vector<uint32_t> tr_pixel;
    std::vector<int32_t> tr_fix_best;  // host-resolved (random) items: index list
    std::vector<uint64_t> tr_fix_idx;
    uint64_t trace_n = 0;

    tsb_stats stats{};
    int max_ctas = 0, n_sms = 0;

    ~tsb_generator() {
        if (h_ctrl) cudaFreeHost(h_ctrl);
        if (stream) cudaStreamDestroy(stream);
        if (stream2) cudaStreamDestroy(stream2);
        if (stream3) cudaStreamDestroy(stream3);
        if (h_progress) cudaFreeHost(h_progress);
    }
};

namespace {

int set_device(tsb_generator* g) { CU(cudaSetDevice(g->device)); return 0; }

void fill_stage_geometry(tsb_generator* g, StageDev& S, bool tiling) {
    memset(&S, 0, sizeof(S));
    S.state = g->d_state.p; S.mask = g->d_mask.p; S.mask1 = g->d_mask1.p; S.score = g->d_score.p;
    S.wpr1 = g->wpr1; S.n_points_max = 0xFFFFFFFFu;
    S.W = g->W; S.H = g->H;
    S.mx = g->mx; S.my = g->my; S.wpr = g->wpr; S.mrows = g->mrows;
    S.tiling = tiling ? 1 : 0;
    S.x_l = (int)((float)g->W * 0.05f); S.x_r = g->W - S.x_l;   // ms.rs:308-311
    S.y_b = (int)((float)g->H * 0.05f); S.y_t = g->H - S.y_b;
    S.spiral = g->d_spiral.p; S.cntLE = g->d_cntLE.p; S.spiralN = g->spiralN; S.RT2 = g->RT2;
    S.divx = g->d_divx.p; S.divy = g->d_divy.p;
    S.k = 1; S.m = 0; S.r2_hint = 16;
    S.counters = nullptr;
}

int init_state(tsb_generator* g) {
    const uint32_t n = (uint32_t)g->W * (uint32_t)g->H;
    cudaStream_t s = g->stream;
    if (g->inpaint)
        k_state_init_inpaint<<<(n + 255) / 256, 256, 0, s>>>(g->d_state.p, g->d_score.p, g->d_inp_mask.p, g->d_inp_color.p, g->W, n, g->inpaint_index);
    else
        k_state_init<<<(n + 255) / 256, 256, 0, s>>>(g->d_state.p, g->d_score.p, n);
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(g->d_mask.p, 0, (size_t)g->wpr * g->mrows * 4, s));
    CU(cudaMemsetAsync(g->d_mask1.p, 0, (size_t)g->wpr1 * g->mrows * 4, s));
    g->unresolved = g->unresolved0;
    g->resolved_order = g->resolved0;
    g->locked = g->inpaint_locked = g->resolved0.size();
    g->have_loaded_points = false;
    g->state_init_opaque = g->inpaint_opaque;
    return 0;
}

// upload one pyramid (levels*w*h RGBA) into a device buffer
int upload_pyramid(const tsb_pyramid& p, DevBuf<uint32_t>& d, cudaStream_t s) {
    size_t n = (size_t)p.n_levels * p.width * p.height;
    return d.upload((const uint32_t*)p.levels, n, s);
}
size_t framed_level_size(int w, int h) { return (size_t)(w + 2 * EX_PAD) * (size_t)(h + 2 * EX_PAD); }
// framed copy of an uploaded pyramid (all levels); also raises flag[0] when a texel is not fully opaque
int frame_pyramid(const DevBuf<uint32_t>& src, DevBuf<uint32_t>& dst, int w, int h, int levels, uint32_t* flag, cudaStream_t s) {
    const size_t n = framed_level_size(w, h) * (size_t)levels;
    TRY(dst.ensure(n));
    const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 64);
    k_frame_levels<<<grid, 256, 0, s>>>(src.p, dst.p, w, h, levels, flag);
    CU(cudaGetLastError());
    return 0;
}

int upload_inputs(tsb_generator* g, const tsb_pyramid* examples, uint32_t n_examples, const tsb_guides* guides, const tsb_sampling* sampling) {
    if (!examples || n_examples == 0) return fail(TSB_ERR_INVALID, "at least one example is required");
    cudaStream_t s = g->stream;
    g->inputs_ready = false;
    g->n_ex_all = (int)n_examples;
    g->n_levels = (int)examples[0].n_levels;
    g->ex_w.clear(); g->ex_h.clear(); g->ex_kind.clear(); g->filt.clear();
    if (g->d_ex.size() != n_examples) {
        g->d_ex.clear(); g->d_smask.clear(); g->d_exf.clear();
        g->d_ex.resize(n_examples); g->d_smask.resize(n_examples); g->d_exf.resize(n_examples);
    }
    if (g->n_levels > 64) return fail(TSB_ERR_UNSUPPORTED, "more than 64 pyramid levels");
    TRY(g->d_alpha_flag.ensure(65));
    CU(cudaMemsetAsync(g->d_alpha_flag.p, 0, 65 * 4, s));
    for (uint32_t e = 0; e < n_examples; ++e) {
        const tsb_pyramid& p = examples[e];
        if (!p.levels || p.width == 0 || p.height == 0 || p.n_levels == 0) return fail(TSB_ERR_INVALID, "example %u is empty", e);
        if ((int)p.n_levels != g->n_levels) return fail(TSB_ERR_INVALID, "all example pyramids must have the same number of levels");
        if (p.width > 32767 || p.height > 32767) return fail(TSB_ERR_UNSUPPORTED, "example dimensions above 32767 are not supported");
        int kind = sampling ? sampling[e].kind : TSB_SAMPLE_ALL;
        g->ex_w.push_back((int)p.width); g->ex_h.push_back((int)p.height); g->ex_kind.push_back(kind);
        TRY(upload_pyramid(p, g->d_ex[e], s));
        TRY(frame_pyramid(g->d_ex[e], g->d_exf[e], (int)p.width, (int)p.height, g->n_levels, g->d_alpha_flag.p, s));
        if (kind == TSB_SAMPLE_IMAGE) {
            if (!sampling[e].rgba) return fail(TSB_ERR_INVALID, "sampling mask %u is null", e);
            std::vector<uint8_t> r((size_t)p.width * p.height);
            bool any = false;
            for (size_t i = 0; i < r.size(); ++i) { r[i] = sampling[e].rgba[i * 4]; any |= r[i] != 0; }
            if (!any) return fail(TSB_ERR_INVALID, "sampling mask %u allows no pixel (the reference would loop forever, ms.rs:562-574)", e);
            TRY(g->d_smask[e].upload(r.data(), r.size(), s));
            CU(cudaStreamSynchronize(s));
        }
        if (kind != TSB_SAMPLE_IGNORE) g->filt.push_back((int)e);
    }
    g->n_ex = (int)g->filt.size();
    if (g->n_ex == 0) return fail(TSB_ERR_INVALID, "at least one example must not be ignored (session.rs:501-524)");
    if (g->n_ex > 255) return fail(TSB_ERR_UNSUPPORTED, "more than 255 examples");
    // descriptors per level, filtered (get_single_example_level, ms.rs:1552-1563)
    std::vector<DevEx> desc((size_t)g->n_levels * g->n_ex);
    for (int l = 0; l < g->n_levels; ++l)
        for (int f = 0; f < g->n_ex; ++f) {
            int e = g->filt[f];
            DevEx d;
            d.w = g->ex_w[e]; d.h = g->ex_h[e];
            d.px = g->d_ex[e].p + (size_t)l * d.w * d.h;
            d.pp = g->d_exf[e].p + (size_t)l * framed_level_size(d.w, d.h) + (size_t)EX_PAD * (d.w + 2 * EX_PAD) + EX_PAD;
            d.smask = g->ex_kind[e] == TSB_SAMPLE_IMAGE ? g->d_smask[e].p : nullptr;
            desc[(size_t)l * g->n_ex + f] = d;
        }
    TRY(g->d_exdesc.upload(desc.data(), desc.size(), s));
    g->pad_pitch = g->ex_w[g->filt[0]] + 2 * EX_PAD;
    for (int f = 1; f < g->n_ex; ++f) if (g->ex_w[g->filt[f]] + 2 * EX_PAD != g->pad_pitch) g->pad_pitch = 0;
    g->guided = guides != nullptr;
    g->exg_w.clear(); g->exg_h.clear();
    if (!guides) g->d_exg.clear();
    if (guides) {
        if (guides->n_examples != n_examples) return fail(TSB_ERR_INVALID, "guides must be given for all examples or none (session.rs:501-524)");
        if ((int)guides->target.n_levels != g->n_levels) return fail(TSB_ERR_INVALID, "target guide pyramid level count mismatch");
        g->tgw = (int)guides->target.width; g->tgh = (int)guides->target.height;
        TRY(upload_pyramid(guides->target, g->d_tguide, s));
        {
            const size_t n = (size_t)g->n_levels * g->tgw * g->tgh;
            k_alpha_check<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 64), 256, 0, s>>>(g->d_tguide.p, n, (size_t)g->tgw * g->tgh, g->d_alpha_flag.p);
        }
        if (g->d_exg.size() != n_examples) { g->d_exg.clear(); g->d_exgf.clear(); g->d_exg.resize(n_examples); g->d_exgf.resize(n_examples); }
        std::vector<DevGuide> gd((size_t)g->n_levels * n_examples);
        for (uint32_t e = 0; e < n_examples; ++e) {
            const tsb_pyramid& p = guides->examples[e];
            if ((int)p.n_levels != g->n_levels) return fail(TSB_ERR_INVALID, "example guide %u level count mismatch", e);
            TRY(upload_pyramid(p, g->d_exg[e], s));
            TRY(frame_pyramid(g->d_exg[e], g->d_exgf[e], (int)p.width, (int)p.height, g->n_levels, g->d_alpha_flag.p, s));
            g->exg_w.push_back((int)p.width); g->exg_h.push_back((int)p.height);
            // the framed (bounds-test-free) scoring path addresses example and guide with ONE offset: both must have the
            // same dimensions (a guide of another size keeps its own bounds, ms.rs:1265-1273)
            if ((int)p.width != g->ex_w[e] || (int)p.height != g->ex_h[e] || (int)p.width + 2 * EX_PAD != g->pad_pitch) g->pad_pitch = 0;
            for (int l = 0; l < g->n_levels; ++l) {
                DevGuide d;
                d.w = (int)p.width; d.h = (int)p.height;
                d.px = g->d_exg[e].p + (size_t)l * d.w * d.h;
                d.pp = g->d_exgf[e].p + (size_t)l * framed_level_size(d.w, d.h) + (size_t)EX_PAD * (d.w + 2 * EX_PAD) + EX_PAD;
                gd[(size_t)l * n_examples + e] = d;
            }
        }
        TRY(g->d_exgdesc.upload(gd.data(), gd.size(), s));
    }
    uint32_t flags[64];
    CU(cudaMemcpyAsync(flags, g->d_alpha_flag.p, (size_t)g->n_levels * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    g->level_opaque.assign((size_t)g->n_levels, 0);
    for (int l = 0; l < g->n_levels; ++l) g->level_opaque[l] = flags[l] == 0 ? 1 : 0;
    g->inputs_ready = true;
    return 0;
}

// Decides whether the stage working on pyramid level `level` may drop the alpha term: every input texel of that level
// and every colour currently in the synthesis state (re-coloured from that level at the start of the stage,
// ms.rs:687-700; random_init, inpaint, loaded snapshots) must have alpha 255.  (Blurred levels usually hold a few 254s --
// the truncating resize -- so in practice this is the last two stages, which work on the unblurred level: 3/4 of the
// work.)  Runs on g->stream.
int decide_opaque(tsb_generator* g, int level) {
    g->run_opaque = false;
    g->no_fast = getenv("TSB_NO_FAST") != nullptr;  // debug: general scoring path only (bounds-tested reads, alpha term kept)
    if (g->no_fast || level < 0 || level >= (int)g->level_opaque.size() || !g->level_opaque[level]) return 0;
    StageDev S;
    fill_stage_geometry(g, S, false);
    CU(cudaMemsetAsync(g->d_alpha_flag.p + 64, 0, 4, g->stream));
    const size_t n = (size_t)g->W * g->H;
    k_state_alpha_check<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 64), 256, 0, g->stream>>>(S, g->have_loaded_points ? 1 : 0, g->d_alpha_flag.p + 64);
    CU(cudaGetLastError());
    uint32_t flag = 1;
    CU(cudaMemcpyAsync(&flag, g->d_alpha_flag.p + 64, 4, cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    g->run_opaque = flag == 0;
    return 0;
}

int check_params(const tsb_generator* g, const tsb_params* p) {
    // session.rs:450-499
    if (!(p->cauchy_dispersion >= 0.0f && p->cauchy_dispersion <= 1.0f)) return fail(TSB_ERR_INVALID, "cauchy_dispersion must be in [0,1]");
    if (!(p->p >= 0.0f && p->p <= 1.0f)) return fail(TSB_ERR_INVALID, "backtrack_percent must be in [0,1]");
    if (!(p->alpha >= 0.0f && p->alpha <= 1.0f)) return fail(TSB_ERR_INVALID, "guide_alpha must be in [0,1]");
    if (p->max_thread_count == 0) return fail(TSB_ERR_INVALID, "max_thread_count must be >= 1");
    if (p->random_sample_locations == 0) return fail(TSB_ERR_INVALID, "random_sample_locations must be >= 1");
    if (p->p_stages < 0) return fail(TSB_ERR_INVALID, "backtrack_stages must be >= 0");
    if (p->nearest_neighbors == 0 || p->nearest_neighbors > (uint32_t)KMAX) return fail(TSB_ERR_UNSUPPORTED, "nearest_neighbors must be in [1,%d]", KMAX);
    if (p->nearest_neighbors + p->random_sample_locations > (uint64_t)CANDMAX)
        return fail(TSB_ERR_UNSUPPORTED, "nearest_neighbors + random_sample_locations must be <= %d", CANDMAX);
    if (p->random_sample_locations == 0) return fail(TSB_ERR_INVALID, "random_sample_locations must be at least 1 (session.rs:489-496; without random candidates a pixel can be left with no candidate at all)");
    int need_levels = p->p_stages == 0 ? 1 : p->p_stages;
    if (g->n_levels < need_levels) return fail(TSB_ERR_INVALID, "example pyramids have %d levels, %d needed", g->n_levels, need_levels);
    return 0;
}

void stage_inputs(tsb_generator* g, StageDev& S, int level, const tsb_params* p) {
    S.ex = g->d_exdesc.p + (size_t)level * g->n_ex;
    S.n_ex = g->n_ex;
    S.exg = g->guided ? g->d_exgdesc.p + (size_t)level * g->n_ex_all : nullptr;
    S.n_exg = g->guided ? g->n_ex_all : 0;
    S.tguide = g->guided ? g->d_tguide.p + (size_t)level * g->tgw * g->tgh : nullptr;
    S.tgw = g->tgw; S.tgh = g->tgh;
    S.lut_my = g->d_luts.p; S.lut_guide = g->d_luts.p + 256;
    S.k = (int)p->nearest_neighbors; S.m = (int)p->random_sample_locations;
    S.pad_pitch = g->no_fast ? 0 : g->pad_pitch;
    S.opaque = g->run_opaque ? 1 : 0;
}

// PrerenderedU8Function tables (ms.rs:739-742, 853-858, 1110-1120) reduced to |a-b| (256 entries)
// are pruning and early-outs result neutral for these cost tables?  (all entries finite and >= 0)
bool luts_well_behaved(const float* h) {
    for (int i = 0; i < 512; ++i) if (!(h[i] >= 0.0f) || std::isinf(h[i])) return false;
    return true;
}
void fill_luts(const tsb_generator* g, const tsb_params* p, float adaptive_alpha, float* h) {
    float sig2 = p->cauchy_dispersion * p->cauchy_dispersion;
    for (int d = 0; d < 256; ++d) {
        float x = (float)d / 255.0f;
        float x2 = x * x;
        float cauchy = log1pf(x2 / sig2);
        if (g->guided) { h[d] = (1.0f - adaptive_alpha) * cauchy; h[256 + d] = adaptive_alpha * x2; }
        else { h[d] = cauchy; h[256 + d] = 0.0f; }
    }
}
int upload_luts(tsb_generator* g, const tsb_params* p, float adaptive_alpha) {
    float h[512];
    fill_luts(g, p, adaptive_alpha, h);
    g->luts_exact_mode = !luts_well_behaved(h);
    CU(cudaMemcpyAsync(g->d_luts.p, h, sizeof(h), cudaMemcpyHostToDevice, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    return 0;
}

uint32_t r2_hint_for(const tsb_generator* g, size_t resolved_now, uint32_t k) {
    double area = (double)g->W * (double)g->H;
    double r2 = 1.5 * (double)k * area / (3.14159265358979 * (double)std::max<size_t>(resolved_now, 1));
    double cap = 4.0e9;
    return (uint32_t)std::min(std::max(r2, 8.0), cap);
}

// Stage plan (ms.rs:786-812): levels, seeds and work-item counts of every stage follow from the parameters alone.
void build_plan(const tsb_generator* g, const tsb_params* prm, std::vector<StagePlan>& plan, size_t& n_picks, size_t& max_stage_items,
                size_t& max_phase, size_t& total_items) {
    const size_t total = g->unresolved.size();  // ms.rs:710
    size_t resolved_n = g->resolved_order.size(), unresolved_n = total;
    n_picks = 0; max_stage_items = 1; max_phase = 1; total_items = 0;
    int pyramid_level = 0;
    for (int p_stage = prm->p_stages; p_stage >= 0; --p_stage) {
        StagePlan sp;
        sp.p_stage = p_stage;
        sp.level = pyramid_level;
        sp.recolour = pyramid_level > 0;
        pyramid_level = std::min(pyramid_level + 1, prm->p_stages - 1);
        sp.seed = (uint64_t)Pcg32::seed_from_u64(prm->seed + (uint64_t)p_stage).next_u32();
        float fp = powf(prm->p, (float)p_stage) * (float)total;
        sp.pixels_to_resolve = fp <= 0.0f ? 0 : (size_t)fp;
        sp.redo_count = resolved_n - g->locked;
        sp.n_redo = std::min(sp.redo_count, sp.pixels_to_resolve);
        sp.n_new = std::min(sp.pixels_to_resolve - sp.n_redo, unresolved_n);
        sp.pick_base = n_picks;
        sp.resolved_before = resolved_n;
        sp.adaptive_alpha = 0.0f;
        if (g->guided && p_stage > 0) {  // ms.rs:846-851
            float v = prm->alpha * (1.0f - ((float)resolved_n / (float)total));
            sp.adaptive_alpha = v * (v * v);
        }
        n_picks += sp.n_new;
        unresolved_n -= sp.n_new;
        resolved_n += sp.n_new;
        max_stage_items = std::max(max_stage_items, sp.n_redo + sp.n_new);
        max_phase = std::max(max_phase, std::max(sp.n_redo, sp.n_new));
        total_items += sp.n_redo + sp.n_new;
        plan.push_back(sp);
    }
}

// ---- pixel order: pick_random_unresolved (ms.rs:380-389) for every new pixel of every stage, entirely on the
// device: index draws (k_pick_indices), then the swap_remove chain resolved in parallel (k_resolve_picks).
// Every non-locked resolved pixel is a pick, and redo items follow the resolution order (ms.rs:905-907), so
// the work items of every stage are simply a PREFIX of the pick array: item i of a stage = picks[i].
// Leaves the picks in g->d_item_pixel (device), their first 4096 in g->h_items (host; the rest arrives on stream2) and
// the pixels never picked in g->unresolved.
// TSB_DEBUG_PLAN=1: wall-clock attribution of the planning steps (synchronises after each, debugging only)
void plan_mark(tsb_generator* g, const char* what) {
    static thread_local double last = 0.0;
    if (!getenv("TSB_DEBUG_PLAN")) return;
    cudaStreamSynchronize(g->stream);
    const double t = now_ms();
    if (what) fprintf(stderr, "[tsb plan] %-28s %8.3f ms\n", what, t - last);
    last = t;
}

int plan_pixel_order(tsb_generator* g, const std::vector<StagePlan>& plan, size_t n_picks, size_t total, size_t npix) {
    cudaStream_t s = g->stream;
    plan_mark(g, nullptr);
    TRY(g->h_items.ensure(std::max<size_t>(n_picks, 1)));
    TRY(g->d_item_pixel.ensure(std::max<size_t>(n_picks, 1)));
    {
        TRY(g->d_pick_idx.ensure(std::max<size_t>(n_picks, 1)));
        size_t un = total;
        for (auto& sp : plan) {
            if (sp.n_new) {
                uint32_t nn = (uint32_t)sp.n_new;
                k_pick_indices<<<(nn + 255) / 256, 256, 0, s>>>(sp.seed + (uint64_t)sp.redo_count, (uint64_t)un, nn, g->d_pick_idx.p + sp.pick_base);
                CU(cudaGetLastError());
                g->stats.kernel_launches++;
            }
            un -= sp.n_new;
        }
        if (n_picks) {
            const uint32_t T = (uint32_t)n_picks;
            TRY(g->d_keys.ensure(n_picks)); TRY(g->d_keys_sorted.ensure(n_picks));
            // v0: the unresolved list as it is now (identity unless inpaint / random_init removed entries)
            bool identity = g->unresolved.size() == npix;
            if (identity && (g->inpaint || g->locked)) identity = false;
            const uint32_t* v0 = nullptr;
            if (!identity) { TRY(g->d_v0.upload(g->unresolved.data(), g->unresolved.size(), s)); v0 = g->d_v0.p; }
            plan_mark(g, "pick indices");
            k_pick_keys<<<(T + 255) / 256, 256, 0, s>>>(g->d_pick_idx.p, T, g->d_keys.p);
            int end_bit = 33;
            while (end_bit < 64 && (total >> (end_bit - 32)) != 0) ++end_bit;
            size_t tb = 0;
            CU(cub::DeviceRadixSort::SortKeys(nullptr, tb, g->d_keys.p, g->d_keys_sorted.p, (int)T, 0, end_bit, s));
            TRY(g->d_sort_temp.ensure(tb + 256));
            tb = g->d_sort_temp.n;
            CU(cub::DeviceRadixSort::SortKeys(g->d_sort_temp.p, tb, g->d_keys.p, g->d_keys_sorted.p, (int)T, 0, end_bit, s));
            plan_mark(g, "sort");
            // directory of the runs of equal position: histogram of the drawn positions + exclusive scan
            TRY(g->d_pick_first.ensure(total + 2));
            CU(cudaMemsetAsync(g->d_pick_first.p, 0, (total + 2) * 4, s));
            k_pick_histogram<<<(T + 255) / 256, 256, 0, s>>>(g->d_pick_idx.p, T, g->d_pick_first.p);
            {
                size_t sb = 0;
                CU(cub::DeviceScan::ExclusiveSum(nullptr, sb, g->d_pick_first.p, g->d_pick_first.p, (int)(total + 1), s));
                TRY(g->d_sort_temp.ensure(sb + 256));
                sb = g->d_sort_temp.n;
                CU(cub::DeviceScan::ExclusiveSum(g->d_sort_temp.p, sb, g->d_pick_first.p, g->d_pick_first.p, (int)(total + 1), s));
            }
            plan_mark(g, "run directory");
            k_resolve_picks<<<(T + 255) / 256, 256, 0, s>>>(g->d_keys_sorted.p, g->d_pick_first.p, T, (uint64_t)total, v0, g->d_pick_idx.p, g->d_item_pixel.p);
            CU(cudaGetLastError());
            plan_mark(g, "resolve picks");
            g->stats.kernel_launches += 4;
            // host mirrors: the whole order is needed only after the run (resolved list, trace); the first few picks
            // are needed right away (first random pixel, serial-prefix bookkeeping)
            const size_t head = std::min<size_t>(n_picks, 4096);
            CU(cudaMemcpyAsync(g->h_items.p, g->d_item_pixel.p, head * 4, cudaMemcpyDeviceToHost, s));
            CU(cudaStreamSynchronize(s));
            // (the rest of the host mirror is copied by copy_picks_tail once nothing small has to cross PCIe any more)
            const size_t left = total - n_picks;
            if (left) {
                TRY(g->d_tmp_u32.ensure(left));
                k_resolve_leftover<<<(uint32_t)((left + 255) / 256), 256, 0, s>>>(g->d_keys_sorted.p, g->d_pick_first.p, T, (uint64_t)total, v0, (uint32_t)left, g->d_tmp_u32.p);
                CU(cudaGetLastError());
                std::vector<uint32_t> rest(left);
                CU(cudaMemcpyAsync(rest.data(), g->d_tmp_u32.p, left * 4, cudaMemcpyDeviceToHost, s));
                CU(cudaStreamSynchronize(s));
                g->unresolved.swap(rest);
            } else {
                g->unresolved.clear();
            }
        }
    }
    return 0;
}

// The host mirror of the whole pick array is needed only after the run (resolved list, trace): a large device->host copy on the
// copy stream.  Issued AFTER the planning's own small read-backs -- a 4-byte copy queued behind 268 MB on the same copy engine
// waits 10-20 ms (measured on 8 GPUs sharing the host's PCIe).
int copy_picks_tail(tsb_generator* g, size_t n_picks, cudaEvent_t picks_ready) {
    const size_t head = std::min<size_t>(n_picks, 4096);
    if (n_picks <= head) return 0;
    CU(cudaStreamWaitEvent(g->stream3, picks_ready, 0));
    CU(cudaMemcpyAsync(g->h_items.p + head, g->d_item_pixel.p + head, (n_picks - head) * 4, cudaMemcpyDeviceToHost, g->stream3));
    return 0;
}

// =============================================================================================
// In-order streaming scheduler (default).  See tsb_stream.cuh for the device side.
// Host: ONE pass that enqueues everything -- the analysis of every chunk on stream2 (it depends on (seed, size, parameters)
// only and therefore runs ahead of the synthesis, limited by the size of the list ring), the resolve kernels on the main
// stream, linked by events -- and a single synchronisation at the end (plus the progress poll when a callback is given).
// =============================================================================================
struct ChunkPlan {
    int stage;              // index into the stage plan
    bool redo;
    size_t first, n;        // stage work-item range
    size_t slot;            // first item slot in the list ring
    bool phase_first, phase_last;
    bool sharded = false;   // band-sharded run: this rank resolves only the items of its band, [own_lo, own_lo + own_n) of its own list
    size_t own_lo = 0, own_n = 0;
    size_t items() const { return sharded ? own_n : n; }
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr, ev_t0 = nullptr, ev_a0 = nullptr;
};

template <bool REDO, bool MG>
int launch_stream(tsb_generator* g, int grid, const StageDev& S, const ChunkDev& C, const StreamDev& D) {
    cudaStream_t s = g->stream;
    const bool op = S.opaque != 0;
    if (g->guided) { if (op) k_stream<true, true, REDO, MG><<<grid, CTA_THREADS, sizeof(StreamSmem), s>>>(S, C, D); else k_stream<true, false, REDO, MG><<<grid, CTA_THREADS, sizeof(StreamSmem), s>>>(S, C, D); }
    else { if (op) k_stream<false, true, REDO, MG><<<grid, CTA_THREADS, sizeof(StreamSmem), s>>>(S, C, D); else k_stream<false, false, REDO, MG><<<grid, CTA_THREADS, sizeof(StreamSmem), s>>>(S, C, D); }
    CU(cudaGetLastError());
    return 0;
}

int resolve_stream(tsb_generator* g, const tsb_params* prm, tsb_progress_fn cb, void* user) {
    if (!g->inputs_ready) return fail(TSB_ERR_INVALID, "inputs have not been uploaded");
    TRY(check_params(g, prm));
    TRY(set_device(g));
    const double t_start = now_ms();
    cudaStream_t s = g->stream, s2 = g->stream2;
    memset(&g->stats, 0, sizeof(g->stats));
    EventPool& events = g->events;
    events.rewind();
    cudaEvent_t ev_begin, ev_end, ev_a0, ev_a1, ev_picks;
    TRY(events.get(&ev_begin, true)); TRY(events.get(&ev_end, true)); TRY(events.get(&ev_a0, true)); TRY(events.get(&ev_a1, true));
    TRY(events.get(&ev_picks, false));
    CU(cudaEventRecord(ev_begin, s));
    const bool tiling = prm->tiling_mode != 0;
    const uint32_t k = prm->nearest_neighbors;
    const int m = (int)prm->random_sample_locations;
    const size_t total = g->unresolved.size();  // ms.rs:710
    const size_t npix = (size_t)g->W * g->H;

    // ---- stage plan (ms.rs:786-812) and pixel order ----
    const double t_plan0 = now_ms();
    std::vector<StagePlan> plan;
    size_t n_picks = 0, max_stage_items = 1, max_phase = 1, total_items = 0;
    build_plan(g, prm, plan, n_picks, max_stage_items, max_phase, total_items);
    if (max_stage_items > 0xFFFFFFF0ull) return fail(TSB_ERR_UNSUPPORTED, "output too large");
    if (plan.size() > 120) return fail(TSB_ERR_UNSUPPORTED, "more than 119 backtrack stages");
    TRY(plan_pixel_order(g, plan, n_picks, total, npix));
    const uint32_t* stage_pixels = g->h_items.p;  // host mirror (only its head is valid before the run ends)
    TRY(g->d_tmap.ensure(npix));
    CU(cudaMemsetAsync(g->d_tmap.p, 0xFF, npix * 4, s));
    if (n_picks) k_tmap_fill<<<(uint32_t)((n_picks + 255) / 256), 256, 0, s>>>(g->d_item_pixel.p, (uint32_t)n_picks, g->d_tmap.p);
    CU(cudaGetLastError());
    CU(cudaEventRecord(ev_picks, s));
    plan_mark(g, "head copy + tmap");
    g->stats.host_ms_schedule = now_ms() - t_plan0;

    // ---- cost tables of every stage (ms.rs:739-742, 853-858), one upload ----
    const size_t n_stages = plan.size();
    TRY(g->h_luts_all.ensure(n_stages * 512)); TRY(g->d_luts_all.ensure(n_stages * 512));
    for (size_t si = 0; si < n_stages; ++si) fill_luts(g, prm, plan[si].adaptive_alpha, g->h_luts_all.p + si * 512);
    CU(cudaMemcpyAsync(g->d_luts_all.p, g->h_luts_all.p, n_stages * 512 * sizeof(float), cudaMemcpyHostToDevice, s));

    // ---- chunks: runs of consecutive work items of one phase ----
    // Chunk size = granularity of the analysis -> resolve pipeline: the resolve kernel of a chunk starts when its lists are
    // complete, so the first chunks of a run are small (the synthesis starts early) and no chunk is so large that the last
    // one's resolve kernel -- which nothing overlaps -- matters.
    size_t chunk_max = 2u << 20;  // (512 Ki-item chunks were tried: the shorter tail does not pay for 2.5x as many launches)
    if (const char* e = getenv("TSB_CHUNK")) chunk_max = std::max<size_t>(16, (size_t)strtoull(e, nullptr, 10));
    std::vector<ChunkPlan> chunks;
    for (size_t si = 0; si < n_stages; ++si) {
        const StagePlan& sp = plan[si];
        auto add_phase = [&](bool redo, size_t a, size_t b) {
            // (a ramp of small first chunks -- 4 Ki, 16 Ki, ... -- starts the synthesis 1.4 ms earlier but every chunk boundary
            // drains the dependency pipeline of the sparse first phase: 51.0 instead of 49.8 ms per 2048^2 step)
            // (cutting a large sparse first phase -- 8192^2: 2 Mi items -- into eighths so that its analysis overlaps its own
            // resolution was tried as well: 696 -> 706 ms)
            const size_t ramp = chunk_max;
            for (size_t c0 = a; c0 < b;) {
                ChunkPlan c;
                c.stage = (int)si; c.redo = redo; c.first = c0; c.n = std::min(std::min(ramp, chunk_max), b - c0); c.slot = 0;
                c.phase_first = c0 == a; c.phase_last = c0 + c.n == b;
                chunks.push_back(c);
                c0 += c.n;
            }
        };
        if (sp.n_redo) add_phase(true, 0, sp.n_redo);
        size_t a = sp.n_redo;
        if (sp.n_new && sp.resolved_before == 0) a += 1;  // the very first pixel of a fresh run has no neighbour: resolve_at_random (ms.rs:1002-1009), see begin_stage
        if (sp.n_redo + sp.n_new > a) add_phase(false, a, sp.n_redo + sp.n_new);
    }
    // ---- band-sharded run: which phases are sharded, and this rank's items of every sharded chunk ----
    if (g->mgs_on && g->trace) return fail(TSB_ERR_UNSUPPORTED, "per-item trace is not available in multi-GPU mode");
    if (g->mgs_on && n_picks) {
        // A phase is sharded once its k-NN discs are small against the band height (the estimate r^2 = k * area / (pi *
        // resolved) shrinks monotonically) and it has enough items; from then on every phase is (a replica only keeps its
        // own band up to date).  Sparse phases are dependency bound anyway: every rank executes them on its own replica.
        bool latched = false;
        const double area = (double)g->W * (double)g->H;
        // resolved pixels needed for r * 8 <= band height
        double rfactor = 8.0;
        if (const char* e = getenv("TSB_MG_RFACTOR")) rfactor = std::max(1.0, atof(e));
        const double rmax = std::max(1.0, (double)g->mgs_band_h / rfactor);
        const size_t need = (size_t)std::ceil((double)k * area / (3.14159265358979 * rmax * rmax));
        std::vector<ChunkPlan> out;
        for (auto& c : chunks) {
            const StagePlan& sp = plan[c.stage];
            if (latched) { c.sharded = true; out.push_back(c); continue; }
            const size_t before = sp.resolved_before + (c.redo ? 0 : c.first - sp.n_redo);  // resolved pixels when the chunk starts
            if (c.redo) {  // the resolved set is static during a redo phase
                if (before >= need && c.n >= g->mgs_shard_min) { latched = true; c.sharded = true; }
                out.push_back(c);
                continue;
            }
            if (before >= need) {
                if (c.n >= g->mgs_shard_min) { latched = true; c.sharded = true; }
                out.push_back(c);
            } else if (before + c.n > need && before + c.n - need >= g->mgs_shard_min) {
                // the threshold is crossed inside this chunk: replicated head, sharded tail
                ChunkPlan a = c, b = c;
                a.n = need - before; a.phase_last = false;
                b.first = c.first + a.n; b.n = c.n - a.n; b.phase_first = false; b.sharded = true;
                latched = true;
                out.push_back(a); out.push_back(b);
            } else out.push_back(c);
        }
        chunks.swap(out);
        const uint32_t T = (uint32_t)n_picks;
        TRY(g->d_own_flag.ensure(n_picks + 1)); TRY(g->d_own_pos.ensure(n_picks + 1)); TRY(g->d_own_t.ensure(n_picks)); TRY(g->d_own_pix.ensure(n_picks));
        TRY(g->d_own_cnt.ensure(4));
        CU(cudaMemsetAsync(g->d_own_flag.p + n_picks, 0, 1, s));
        k_own_flags<<<(T + 255) / 256, 256, 0, s>>>(g->d_item_pixel.p, T, g->W, g->mgs_band_h, g->mgs_world, g->mgs_rank, g->d_own_flag.p);
        CU(cudaGetLastError());
        size_t tb1 = 0, tb2 = 0;
        CU(cub::DeviceScan::ExclusiveSum(nullptr, tb1, g->d_own_flag.p, g->d_own_pos.p, (int)(n_picks + 1), s));
        cub::CountingInputIterator<uint32_t> counting(0u);
        CU(cub::DeviceSelect::Flagged(nullptr, tb2, counting, g->d_own_flag.p, g->d_own_t.p, g->d_own_cnt.p, (int)n_picks, s));
        TRY(g->d_cub_temp.ensure(std::max(tb1, tb2) + 256));
        tb1 = tb2 = g->d_cub_temp.n;
        CU(cub::DeviceScan::ExclusiveSum(g->d_cub_temp.p, tb1, g->d_own_flag.p, g->d_own_pos.p, (int)(n_picks + 1), s));
        CU(cub::DeviceSelect::Flagged(g->d_cub_temp.p, tb2, counting, g->d_own_flag.p, g->d_own_t.p, g->d_own_cnt.p, (int)n_picks, s));
        // own counts at the chunk boundaries
        std::vector<uint32_t> bidx;
        for (auto& c : chunks) { bidx.push_back((uint32_t)c.first); bidx.push_back((uint32_t)(c.first + c.n)); }
        // (no allocation in the steady state: with peer mappings enabled every cudaMalloc / cudaFree is a cross-process affair)
        DevBuf<uint32_t>& d_bidx = g->d_own_bidx;
        DevBuf<uint32_t>& d_bpos = g->d_own_bpos;
        std::vector<uint32_t> bpos(bidx.size());
        TRY(d_bidx.upload(bidx.data(), bidx.size(), s)); TRY(d_bpos.ensure(bidx.size()));
        k_gather_u32<<<(uint32_t)((bidx.size() + 255) / 256), 256, 0, s>>>(g->d_own_pos.p, d_bidx.p, (uint32_t)bidx.size(), d_bpos.p);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(bpos.data(), d_bpos.p, bidx.size() * 4, cudaMemcpyDeviceToHost, s));
        uint32_t n_own_total = 0;
        CU(cudaMemcpyAsync(&n_own_total, g->d_own_cnt.p, 4, cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        if (n_own_total) k_gather_u32<<<(n_own_total + 255) / 256, 256, 0, s>>>(g->d_item_pixel.p, g->d_own_t.p, n_own_total, g->d_own_pix.p);
        CU(cudaGetLastError());
        CU(cudaEventRecord(ev_picks, s));  // the analysis stream also needs the own-item lists
        plan_mark(g, "own-item lists");
        for (size_t i = 0; i < chunks.size(); ++i) { chunks[i].own_lo = bpos[2 * i]; chunks[i].own_n = bpos[2 * i + 1] - bpos[2 * i]; }
    }
    TRY(copy_picks_tail(g, n_picks, ev_picks));
    // ---- list ring ----
    const size_t bytes_per_item = 1 + (size_t)k * 8 + 16 + (size_t)m * 5;
    size_t ring_mb = 24576;
    if (const char* e = getenv("TSB_RING_MB")) ring_mb = std::max<size_t>(64, (size_t)strtoull(e, nullptr, 10));
    size_t ring_items = std::min(std::max<size_t>(total_items, 1), ring_mb * (1u << 20) / bytes_per_item);
    if (g->r_nbk.n < ring_items) {  // only when the ring has to grow (a repeated run allocates nothing)
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            size_t have = g->r_nb.n * sizeof(short2) + g->r_g.n * 4 + g->r_low.n * 16 + g->r_rand_xy.n * 4 + g->r_rand_map.n + g->r_nbk.n;
            size_t cap = (free_b + have) / 2 / bytes_per_item;  // never take more than half of what is free
            ring_items = std::min(ring_items, std::max<size_t>(cap, 1));
        }
    }
    if (const char* e = getenv("TSB_RING_ITEMS")) ring_items = std::max<size_t>(16, (size_t)strtoull(e, nullptr, 10));  // tests: force the ring to wrap
    size_t largest = 1;
    for (auto& c : chunks) largest = std::max(largest, c.items());
    if (g->mgs_on) ring_items = std::max(ring_items, 2 * largest);  // (own-item ranges are per chunk: no re-splitting)
    if (ring_items < 2 * largest) {  // smaller chunks so that two of them fit the ring
        const size_t cm = std::max<size_t>(8, ring_items / 3);
        std::vector<ChunkPlan> split;
        for (auto& c : chunks)
            for (size_t o = 0; o < c.n; o += cm) {
                ChunkPlan d = c;
                d.first = c.first + o; d.n = std::min(cm, c.n - o);
                d.phase_first = c.phase_first && o == 0; d.phase_last = c.phase_last && o + d.n == c.n;
                split.push_back(d);
            }
        chunks.swap(split);
        largest = std::min(largest, cm);
        ring_items = std::max(ring_items, 2 * largest);
    }
    TRY(g->r_nbk.ensure(ring_items)); TRY(g->r_nb.ensure(ring_items * k)); TRY(g->r_g.ensure(ring_items * k)); TRY(g->r_low.ensure(ring_items));
    TRY(g->r_rand_xy.ensure(ring_items * (size_t)m)); TRY(g->r_rand_map.ensure(ring_items * (size_t)m));
    TRY(g->d_sctl.ensure((chunks.size() + 1) * SC_WORDS));
    CU(cudaMemsetAsync(g->d_sctl.p, 0, (chunks.size() + 1) * SC_WORDS * 4, s));
    uint32_t* abort_flag = g->d_sctl.p + chunks.size() * SC_WORDS;
    TRY(g->d_counters.ensure(ST_COUNT));
    CU(cudaMemsetAsync(g->d_counters.p, 0, ST_COUNT * sizeof(unsigned long long), s));
    TRY(g->d_pmask.ensure((size_t)g->wpr * g->mrows));
    TRY(g->d_pmask1.ensure((size_t)g->wpr1 * g->mrows));
    TRY(g->d_state2.ensure(npix));
    g->trace_n = 0;
    g->tr_pixel.clear(); g->tr_fix_idx.clear();
    if (g->trace) {
        TRY(g->d_tr_best.ensure(total_items)); TRY(g->d_tr_ncand.ensure(total_items));
        TRY(g->d_tr_nneigh.ensure(total_items)); TRY(g->d_tr_score.ensure(total_items));
    }
    plan_mark(g, "ring + buffers");
    for (auto& c : chunks) { TRY(events.get(&c.ev_ready, true)); TRY(events.get(&c.ev_done, true)); TRY(events.get(&c.ev_t0, true)); TRY(events.get(&c.ev_a0, true)); }
    uint32_t watchdog_ms = 20000;
    if (const char* e = getenv("TSB_WATCHDOG_MS")) watchdog_ms = (uint32_t)std::max(1, atoi(e));
    *g->h_progress = 0;

    // ---- analysis stream: resolved-set mask with the tiling mirrors (ms.rs:747-779), then chunk after chunk ----
    CU(cudaStreamWaitEvent(s2, ev_picks, 0));
    CU(cudaEventRecord(ev_a0, s2));
    StageDev A;  // geometry + analysis-owned mask
    fill_stage_geometry(g, A, tiling);
    A.k = (int)k; A.m = m;
    CU(cudaMemsetAsync(g->d_mask.p, 0, (size_t)g->wpr * g->mrows * 4, s2));
    CU(cudaMemsetAsync(g->d_mask1.p, 0, (size_t)g->wpr1 * g->mrows * 4, s2));
    DevBuf<uint32_t>& d_init_points = g->d_init_points;
    if (g->have_loaded_points) {
        TRY(d_init_points.upload((const uint32_t*)g->loaded_points.data(), g->loaded_points.size(), s2));
        uint32_t np = (uint32_t)(g->loaded_points.size() / 2);
        if (np) k_mask_insert_points<<<(np + 255) / 256, 256, 0, s2>>>(A, (const int32_t*)d_init_points.p, np);
    } else if (!g->resolved_order.empty()) {
        TRY(d_init_points.upload(g->resolved_order.data(), g->resolved_order.size(), s2));
        uint32_t np = (uint32_t)g->resolved_order.size();
        k_mask_insert_flat<<<(np + 255) / 256, 256, 0, s2>>>(A, d_init_points.p, np, tiling ? 1 : 0);
    }
    CU(cudaGetLastError());

    std::deque<size_t> live;       // chunks whose lists occupy the ring (analysis enqueued, slots not yet reclaimed)
    size_t ring_head = 0;
    size_t a_next = 0, r_next = 0;  // next chunk to analyse / to resolve
    // Exclusive batches: beside a throughput-bound resolve chunk the analysis kernels only take turns with it -- both get slower
    // (a 2048^2 step: resolve launches +5 ms, the analysis 38 instead of 16 ms) -- so analysis and resolve alternate in batches
    // as large as the list ring allows; only the dependency-bound first phase, which leaves most of the machine idle, runs
    // beside the analysis.  TSB_EXCLUSIVE=0 restores free overlap.
    static const bool exclusive = !(getenv("TSB_EXCLUSIVE") && atoi(getenv("TSB_EXCLUSIVE")) == 0);
    bool batch_open = false;  // an analysis batch is being enqueued (its first chunk has waited for the resolve stream)
    auto chunk_dev = [&](const ChunkPlan& c) {
        ChunkDev C;
        C.pixel = g->d_item_pixel.p + c.first;
        C.nbk = g->r_nbk.p + c.slot; C.nb = g->r_nb.p + c.slot * k; C.g = g->r_g.p + c.slot * k; C.low = g->r_low.p + c.slot;
        C.rand_xy = g->r_rand_xy.p + c.slot * (size_t)m; C.rand_map = g->r_rand_map.p + c.slot * (size_t)m;
        C.n = (uint32_t)c.n; C.first = (uint32_t)c.first; C.tidx = nullptr;
        if (c.sharded) { C.pixel = g->d_own_pix.p + c.own_lo; C.tidx = g->d_own_t.p + c.own_lo; C.n = (uint32_t)c.own_n; }
        return C;
    };
    // returns 1 when the ring has no room until more resolve kernels have been enqueued
    auto enqueue_analysis = [&](ChunkPlan& c) -> int {
        size_t off = ring_head;
        if (off + c.items() > ring_items) off = 0;
        // slots are reclaimed in FIFO order: everything up to the LAST live chunk that overlaps the new range must be done
        size_t reclaim = 0;
        for (size_t i = 0; i < live.size(); ++i) {
            const ChunkPlan& f = chunks[live[i]];
            if (off < f.slot + f.items() && f.slot < off + c.items()) reclaim = i + 1;
        }
        for (size_t i = 0; i < reclaim; ++i)
            if (live[i] >= r_next) return 1;  // its resolve kernel is not enqueued yet: no event to wait for
        for (size_t i = 0; i < reclaim; ++i) {
            CU(cudaStreamWaitEvent(s2, chunks[live.front()].ev_done, 0));
            live.pop_front();
        }
        if (exclusive && !batch_open && r_next > 0) CU(cudaStreamWaitEvent(s2, chunks[r_next - 1].ev_done, 0));  // after everything resolved so far
        batch_open = true;
        c.slot = off;
        ring_head = off + c.items();
        CU(cudaEventRecord(c.ev_a0, s2));
        const StagePlan& sp = plan[c.stage];
        StageDev S = A;
        S.ex = g->d_exdesc.p + (size_t)sp.level * g->n_ex; S.n_ex = g->n_ex;
        const ChunkDev C = chunk_dev(c);
        const uint32_t n = C.n;  // this rank's items of the chunk
        const int gr = std::max(1, std::min((int)((n + WARPS_PER_CTA - 1) / WARPS_PER_CTA), g->max_ctas_radius));
        const size_t first_new = sp.n_redo + ((sp.n_new && sp.resolved_before == 0) ? 1 : 0);
        const size_t n_new_phase = sp.n_redo + sp.n_new - first_new;
        TimeFilter T;
        memset(&T, 0, sizeof(T));
        if (c.redo) {
            S.r2_hint = r2_hint_for(g, sp.resolved_before, k);
            S.n_points_max = (uint32_t)std::min<size_t>((tiling ? 3 : 1) * sp.resolved_before, 0xFFFFFFFFull);
            if (n) k_lists_chunk<true><<<gr, CTA_THREADS, sizeof(KnnScratch) * WARPS_PER_CTA, s2>>>(S, C, T, g->d_tmap.p, 0u, 0u);
        } else {
            const size_t resolved_now = sp.resolved_before + (first_new - sp.n_redo);
            if (c.phase_first) {
                if (first_new > sp.n_redo) {  // the randomly resolved first pixel joins the set WITHOUT mirror copies (ms.rs:473)
                    k_mask_insert_flat<<<1, 32, 0, s2>>>(A, g->d_item_pixel.p + sp.n_redo, 1u, 0);
                }
                CU(cudaMemsetAsync(g->d_pmask.p, 0, (size_t)g->wpr * g->mrows * 4, s2));
                CU(cudaMemsetAsync(g->d_pmask1.p, 0, (size_t)g->wpr1 * g->mrows * 4, s2));
                const uint32_t nn = (uint32_t)n_new_phase;
                k_mask_insert_flat_at<<<(nn + 255) / 256, 256, 0, s2>>>(A, g->d_pmask.p, g->d_pmask1.p, g->d_item_pixel.p + first_new, nn, tiling ? 1 : 0);
            }
            S.r2_hint = r2_hint_for(g, resolved_now, k);
            S.n_points_max = (uint32_t)std::min<size_t>((tiling ? 3 : 1) * (resolved_now + n_new_phase), 0xFFFFFFFFull);
            T.pend = g->d_pmask.p; T.pend1 = g->d_pmask1.p; T.pmap = g->d_tmap.p;
            T.item_pixel = g->d_item_pixel.p + first_new;
            T.brute_below = (uint32_t)std::min<double>(sqrt((double)n_new_phase * (double)k), 65536.0);
            if (const char* e = getenv("TSB_BRUTE_BELOW")) T.brute_below = (uint32_t)atoi(e);
            if (n) k_lists_chunk<false><<<gr, CTA_THREADS, sizeof(KnnScratch) * WARPS_PER_CTA, s2>>>(S, C, T, g->d_tmap.p, (uint32_t)first_new,
                                                                                                      (uint32_t)std::min<size_t>(resolved_now, 0xFFFFFFFFull));
        }
        CU(cudaGetLastError());
        if (n) {
            const unsigned wgroups = (n + KW_ITEMS - 1) / KW_ITEMS;
            k_weights<<<std::max(1u, std::min((wgroups + KW_WARPS - 1) / KW_WARPS, (unsigned)g->n_sms * 16u)), KW_WARPS * 32, (size_t)KW_WARPS * k * (KW_ITEMS + 1) * sizeof(double), s2>>>(S, C);
            const unsigned rb = m <= 64 ? 128u : 32u;  // items per block: the staging area is rb * m * 5 bytes (< 48 KB)
            if (S.n_ex == 1)
                k_rand_candidates<true><<<(n + rb - 1) / rb, rb, (size_t)rb * m * 5 + rb, s2>>>(S.ex, S.n_ex, m, sp.seed + 1ull + (c.sharded ? 0ull : (uint64_t)c.first), n,
                                                                                             C.rand_xy, C.rand_map, C.tidx);
            else
                k_rand_candidates<false><<<(n + rb - 1) / rb, rb, (size_t)rb * m * 5 + rb, s2>>>(S.ex, S.n_ex, m, sp.seed + 1ull + (c.sharded ? 0ull : (uint64_t)c.first), n,
                                                                                              C.rand_xy, C.rand_map, C.tidx);
            CU(cudaGetLastError());
        }
        if (!c.redo && c.phase_last) {  // the stage's new pixels join the resolved set of the next stage's analysis
            const uint32_t nn = (uint32_t)n_new_phase;
            k_mask_insert_flat<<<(nn + 255) / 256, 256, 0, s2>>>(A, g->d_item_pixel.p + first_new, nn, tiling ? 1 : 0);
            CU(cudaGetLastError());
        }
        CU(cudaEventRecord(c.ev_ready, s2));
        g->stats.kernel_launches += 3;
        return 0;
    };

    // ---- resolve stream ----
    StageDev S;
    fill_stage_geometry(g, S, tiling);
    S.counters = g->d_counters.p;
    bool state_opaque = g->state_init_opaque;  // every resolved colour in the state has alpha 255
    int cur_stage = -1;
    std::vector<uint64_t> stage_trace_base(n_stages, 0), stage_progress_base(n_stages, 0);
    std::vector<const uint4*> stage_buf(n_stages, nullptr);  // the state buffer each stage writes (progress snapshots)
    {
        uint64_t tb = 0, pb = 0;
        for (size_t si = 0; si < n_stages; ++si) { stage_trace_base[si] = tb; stage_progress_base[si] = pb; tb += plan[si].n_redo + plan[si].n_new; pb += plan[si].pixels_to_resolve; }
    }
    int grid_full = g->guided ? g->max_ctas_stream_guided : g->max_ctas_stream;
    if (const char* e = getenv("TSB_STREAM_OCC")) grid_full = std::min(grid_full, g->n_sms * std::max(1, atoi(e)));
    bool stage_prologue_pending = false;  // a recolour happened since the last kernel of this rank
    auto mg_barrier = [&]() -> int {
        const uint32_t seq = ++g->mgs_seq;
        k_mg_signal<<<1, 32, 0, s>>>(g->d_mgs_sync_ptrs.p, g->mgs_world, g->mgs_rank, seq);
        k_mg_wait<<<1, 32, 0, s>>>(g->d_mgs_sync.p, g->mgs_world, seq, watchdog_ms, abort_flag);
        CU(cudaGetLastError());
        return 0;
    };
    auto begin_stage = [&](int si) -> int {
        const StagePlan& sp = plan[si];
        stage_prologue_pending = stage_prologue_pending || sp.recolour;
        stage_inputs(g, S, sp.level, prm);
        S.state = g->d_state.p;
        S.lut_my = g->d_luts_all.p + (size_t)si * 512; S.lut_guide = S.lut_my + 256;
        if (g->l2_persist_bytes) {
            // L2 persisting window over the level every candidate of this stage gathers from (the framed copy of the first
            // example): hits stay resident, everything else the resolve stream touches is ordinary / streaming traffic
            const int e0 = g->filt[0];
            const size_t lvl = framed_level_size(g->ex_w[e0], g->ex_h[e0]) * sizeof(uint32_t);
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof(av));
            av.accessPolicyWindow.base_ptr = (void*)(g->d_exf[e0].p + (size_t)sp.level * framed_level_size(g->ex_w[e0], g->ex_h[e0]));
            av.accessPolicyWindow.num_bytes = std::min(lvl, g->l2_persist_bytes);
            av.accessPolicyWindow.hitRatio = 1.0f;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
            if (cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av) != cudaSuccess) { cudaGetLastError(); g->l2_persist_bytes = 0; }
        }
        if (sp.recolour) {  // next_pyramid_level, ms.rs:687-700
            k_recolour<<<(uint32_t)((npix + 255) / 256), 256, 0, s>>>(S);
            if (cb) k_snapshot_color<<<(unsigned)std::min<size_t>((npix + 255) / 256, 4096), 256, 0, s>>>(g->d_state.p, (uint32_t)npix, g->d_live_color.p);
            CU(cudaGetLastError());
            g->stats.kernel_launches++;
            // recoloured pixels take the level's alphas; a pixel whose source is out of range keeps its colour
            const bool may_keep = g->inpaint || g->locked > 0 || g->have_loaded_points;
            state_opaque = (may_keep ? state_opaque : true) && g->level_opaque[sp.level];
        }
        g->no_fast = getenv("TSB_NO_FAST") != nullptr;
        S.seq_exact = luts_well_behaved(g->h_luts_all.p + (size_t)si * 512) ? 0 : 1;  // q13 / q14: literal candidate loop
        g->run_opaque = !g->no_fast && !S.seq_exact && state_opaque && sp.level < (int)g->level_opaque.size() && g->level_opaque[sp.level];
        S.opaque = g->run_opaque ? 1 : 0;
        S.pad_pitch = g->no_fast ? 0 : g->pad_pitch;
        state_opaque = state_opaque && g->level_opaque[sp.level];  // this stage commits texels of this level
        if (sp.n_new && sp.resolved_before == 0) {
            // no resolved neighbour at all: resolve_at_random(seed = p_stage_seed), ms.rs:1002-1009 -> 447-475
            uint32_t flat = stage_pixels[sp.n_redo];
            uint32_t rmap = (uint32_t)Pcg32::seed_from_u64(sp.seed).gen_range_usize((uint64_t)g->n_ex);
            int e = g->filt[rmap];
            uint32_t rx = Pcg32::seed_from_u64(sp.seed).gen_range_u32((uint32_t)g->ex_w[e]);
            uint32_t ry = Pcg32::seed_from_u64(sp.seed).gen_range_u32((uint32_t)g->ex_h[e]);
            uint32_t* item = g->h_ctrl;  // pinned
            item[0] = flat; item[1] = rx; item[2] = ry; item[3] = rmap;
            TRY(g->d_tmp_u32.ensure(4));
            CU(cudaMemcpyAsync(g->d_tmp_u32.p, item, 16, cudaMemcpyHostToDevice, s));
            k_commit_fixed<<<1, 32, 0, s>>>(S, S.ex, g->d_tmp_u32.p, 1, 0);
            if (cb) k_snapshot_color<<<(unsigned)std::min<size_t>((npix + 255) / 256, 4096), 256, 0, s>>>(g->d_state.p, (uint32_t)npix, g->d_live_color.p);
            CU(cudaGetLastError());
            g->stats.kernel_launches++;
            if (g->trace) g->tr_fix_idx.push_back(stage_trace_base[si] + sp.n_redo);
        }
        return 0;
    };
    auto enqueue_resolve = [&](ChunkPlan& c) -> int {
        const StagePlan& sp = plan[c.stage];
        const size_t ci0 = (size_t)(&c - chunks.data());
        const bool shard_begins = c.sharded && ci0 > 0 && !chunks[ci0 - 1].sharded;  // the peers' replicas are read from here on
        if (c.sharded && (c.phase_first || shard_begins)) TRY(mg_barrier());  // every rank has finished the previous phase: nobody reads what the prologue changes
        if (c.stage != cur_stage) {
            for (int si = cur_stage + 1; si <= c.stage; ++si) TRY(begin_stage(si));  // stages without work items still recolour
            cur_stage = c.stage;
        }
        StreamDev D;
        memset(&D, 0, sizeof(D));
        if (c.redo && c.phase_first) {
            // redo results go to the other buffer: pixels that are not re-resolved must be there too
            size_t o = 0, cnt = npix;
            if (c.sharded) {  // only this rank's band is ever read from this replica
                const int y0 = g->mgs_rank * g->mgs_band_h, y1 = g->mgs_rank == g->mgs_world - 1 ? g->H : std::min(g->H, y0 + g->mgs_band_h);
                o = (size_t)std::min(y0, g->H) * g->W; cnt = (size_t)std::max(0, y1 - y0) * g->W;
            }
            if (cnt) CU(cudaMemcpyAsync(g->d_state2.p + o, g->d_state.p + o, cnt * sizeof(uint4), cudaMemcpyDeviceToDevice, s));
        }
        if (c.redo) { D.prev = g->d_state.p; D.cur = g->d_state2.p; }
        else { D.prev = nullptr; D.cur = g->d_state.p; }
        const size_t ci = (size_t)(&c - chunks.data());
        stage_buf[c.stage] = D.cur;
        D.ctl = g->d_sctl.p + ci * SC_WORDS;
        D.abort_flag = abort_flag;
        D.progress = cb ? g->d_progress : nullptr;
        D.live_color = cb ? g->d_live_color.p : nullptr;
        D.progress_base = (uint32_t)std::min<uint64_t>(stage_progress_base[c.stage] + c.first, 0xFFFFFFFFull);
        D.tag = (uint32_t)(2 * c.stage + (c.redo ? 1 : 2));
        D.watchdog_ms = watchdog_ms;
        D.profile = getenv("TSB_DEBUG_PHASES") ? 1u : 0u;
        if (g->trace) { D.tr_best = g->d_tr_best.p; D.tr_ncand = g->d_tr_ncand.p; D.tr_nneigh = g->d_tr_nneigh.p; D.tr_score = g->d_tr_score.p; }
        D.trace_base = stage_trace_base[c.stage];
        StageDev Sc = S;
        Sc.state = D.cur;
        const ChunkDev C = chunk_dev(c);
        int grid = std::max(1, std::min((int)((C.n + WARPS_PER_CTA - 1) / WARPS_PER_CTA), grid_full));
        // a phase that multiplies the resolved set is bound by its dependency depth, not by throughput: a small grid
        // leaves the rest of the machine to the analysis stream
        const bool dependency_bound = !c.redo && std::max<size_t>(sp.resolved_before, 1) * 4 < sp.n_new;
        // (its dependency chains keep only a few hundred items in flight: one CTA on every other SM resolves it as fast as
        // one per SM -- measured 10.6 vs 9.5 ms -- and costs the analysis beside it less: 43.5 vs 44.4 ms per 2048^2 step)
        static const int dep_grid = getenv("TSB_DEP_GRID") ? atoi(getenv("TSB_DEP_GRID")) : 0;
        if (dependency_bound) grid = std::min(grid, dep_grid > 0 ? dep_grid : std::max(1, g->n_sms / 2));
        if (c.sharded) {
            if (c.phase_first && (c.redo || stage_prologue_pending)) TRY(mg_barrier());  // ... and every replica is through its own prologue
            D.world = g->mgs_world; D.rank = g->mgs_rank; D.band_h = g->mgs_band_h;
            D.y0 = g->mgs_rank * g->mgs_band_h; D.y1 = g->mgs_rank == g->mgs_world - 1 ? g->H : std::min(g->H, D.y0 + g->mgs_band_h);
            const bool flip = g->d_state.p != g->mgs_A[g->mgs_rank];  // the same on every rank
            for (int r = 0; r < g->mgs_world; ++r) {
                const uint4* a = flip ? g->mgs_B[r] : g->mgs_A[r];  // what every rank calls d_state right now
                const uint4* b = flip ? g->mgs_A[r] : g->mgs_B[r];
                D.prev_r[r] = c.redo ? a : nullptr;
                D.cur_r[r] = c.redo ? b : a;
            }
            g->mgs_sharded_chunks++;
        }
        stage_prologue_pending = false;
        CU(cudaStreamWaitEvent(s, c.ev_ready, 0));
        if (exclusive && !dependency_bound && a_next > 0) CU(cudaStreamWaitEvent(s, chunks[a_next - 1].ev_ready, 0));  // after the whole analysis batch
        CU(cudaEventRecord(c.ev_t0, s));
        if (C.n) {
            if (c.sharded) { if (c.redo) TRY((launch_stream<true, true>(g, grid, Sc, C, D))); else TRY((launch_stream<false, true>(g, grid, Sc, C, D))); }
            else { if (c.redo) TRY((launch_stream<true, false>(g, grid, Sc, C, D))); else TRY((launch_stream<false, false>(g, grid, Sc, C, D))); }
        }
        CU(cudaEventRecord(c.ev_done, s));
        g->stats.kernel_launches++;
        g->stats.phases += c.phase_first ? 1 : 0;
        g->stats.rounds++;
        if (c.redo && c.phase_last) std::swap(g->d_state.p, g->d_state2.p);  // both hold npix entries
        return 0;
    };

    // ---- progress (ProgressNotifier, ms.rs:1054-1107): the calling thread polls the claim counter, as the reference's main
    // thread does (ms.rs:1026-1034), and reports every change of the integer percentage with a snapshot of the colours.
    // Polled between the enqueue calls too: on a cold start (lazy kernel loading) the device runs while the host still enqueues.
    uint64_t overall_total = 0;
    for (auto& sp : plan) overall_total += sp.pixels_to_resolve;
    uint32_t last_pcnt = 0;
    if (cb) {
        // A plain colour plane mirrors the commits while a callback is registered (the counterpart of the reference's shared
        // color_map): a snapshot is then ONE copy-engine transfer into pinned memory -- no kernel has to squeeze in beside the
        // persistent resolve grid, nothing is stalled.
        TRY(g->d_live_color.ensure(npix)); TRY(g->h_snap.ensure(npix));
        k_snapshot_color<<<(unsigned)std::min<size_t>((npix + 255) / 256, 4096), 256, 0, s>>>(g->d_state.p, (uint32_t)npix, g->d_live_color.p);
        CU(cudaGetLastError());
    }
    auto report = [&](uint64_t cur_total) -> int {
        if (!cb || overall_total == 0) return 0;
        const uint32_t pcnt = (uint32_t)lroundf((float)cur_total / (float)overall_total * 100.0f);
        if (pcnt == last_pcnt) return 0;
        size_t si = 0;
        while (si + 1 < n_stages && cur_total >= stage_progress_base[si + 1]) ++si;
        while (si > 0 && !stage_buf[si]) --si;
        if (!stage_buf[si]) return 0;  // nothing of that stage is enqueued yet
        last_pcnt = pcnt;
        // snapshot on the copy stream: racy by design, like the reference's read of the shared colour map
        CU(cudaMemcpyAsync(g->h_snap.p, g->d_live_color.p, npix * 4, cudaMemcpyDeviceToHost, g->stream3));
        CU(cudaStreamSynchronize(g->stream3));
        cb(user, (const uint8_t*)g->h_snap.p, (uint32_t)g->W, (uint32_t)g->H, cur_total, overall_total, cur_total - stage_progress_base[si], plan[si].pixels_to_resolve);
        return 0;
    };

    while (r_next < chunks.size()) {
        if (cb) TRY(report(*(volatile uint32_t*)g->h_progress));
        while (a_next < chunks.size()) {
            int rc = enqueue_analysis(chunks[a_next]);
            if (rc == 1) break;
            if (rc != 0) return rc;
            live.push_back(a_next);
            ++a_next;
            if (cb) TRY(report(*(volatile uint32_t*)g->h_progress));
        }
        if (a_next <= r_next) return fail(TSB_ERR_INTERNAL, "list ring too small for chunk %zu", r_next);
        batch_open = false;
        const size_t batch_end = exclusive ? a_next : r_next + 1;
        while (r_next < batch_end) {
            TRY(enqueue_resolve(chunks[r_next]));
            ++r_next;
            if (cb) TRY(report(*(volatile uint32_t*)g->h_progress));
        }
    }
    for (int si = cur_stage + 1; si < (int)n_stages; ++si) TRY(begin_stage(si));
    if (g->mgs_on && !chunks.empty() && chunks.back().sharded) {
        // all-gather of the bands: every replica ends up complete (state and first-resolution scores)
        TRY(mg_barrier());
        const bool flip = g->d_state.p != g->mgs_A[g->mgs_rank];
        for (int r = 0; r < g->mgs_world; ++r) {
            if (r == g->mgs_rank) continue;
            const int y0 = r * g->mgs_band_h, y1 = r == g->mgs_world - 1 ? g->H : std::min(g->H, y0 + g->mgs_band_h);
            if (y1 <= y0) continue;
            const size_t o = (size_t)y0 * g->W, cnt = (size_t)(y1 - y0) * g->W;
            const uint4* src = flip ? g->mgs_B[r] : g->mgs_A[r];
            CU(cudaMemcpyAsync(g->d_state.p + o, src + o, cnt * sizeof(uint4), cudaMemcpyDeviceToDevice, s));
            CU(cudaMemcpyAsync(g->d_score.p + o, g->mgs_score[r] + o, cnt * sizeof(float), cudaMemcpyDeviceToDevice, s));
        }
        TRY(mg_barrier());  // nobody changes its band before everybody has copied it
    }
    if (g->l2_persist_bytes) {  // release the window
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof(av));
        cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &av);
    }
    CU(cudaEventRecord(ev_a1, s2));
    const double t_enqueued = now_ms();
    CU(cudaStreamWaitEvent(s, ev_a1, 0));
    CU(cudaMemcpyAsync(g->h_ctrl + 8, abort_flag, 4, cudaMemcpyDeviceToHost, s));
    CU(cudaEventRecord(ev_end, s));

    // ---- progress (ProgressNotifier, ms.rs:1054-1107): the calling thread polls the claim counter, as the reference's main
    // thread does (ms.rs:1026-1034), and reports every change of the integer percentage with a snapshot of the colours ----
    if (cb) {
        // busy poll of the mapped counter, like the reference's main thread (a plain load per iteration); the event is
        // queried every ~20 us only
        double t_query = now_ms();
        for (;;) {
            TRY(report(*(volatile uint32_t*)g->h_progress));
            const double t = now_ms();
            if (t - t_query < 0.02) continue;
            t_query = t;
            if (cudaEventQuery(ev_end) != cudaErrorNotReady) break;
        }
        TRY(report(overall_total));
    }
    CU(cudaEventSynchronize(ev_end));
    CU(cudaStreamSynchronize(s2));
    if (g->h_ctrl[8]) return fail(TSB_ERR_INTERNAL, "resolve stalled: a work item waited more than %u ms for a neighbour (TSB_WATCHDOG_MS)", watchdog_ms);

    // host mirrors of the order: ms.rs:1043-1049 (newly resolved pixels join `resolved` in processing order)
    CU(cudaStreamSynchronize(g->stream3));
    g->resolved_order.insert(g->resolved_order.end(), g->h_items.p, g->h_items.p + n_picks);
    if (g->trace) for (auto& sp : plan) g->tr_pixel.insert(g->tr_pixel.end(), g->h_items.p, g->h_items.p + sp.n_redo + sp.n_new);
    g->trace_n = g->trace ? total_items : 0;
    unsigned long long cnt[ST_COUNT];
    CU(cudaMemcpy(cnt, g->d_counters.p, sizeof(cnt), cudaMemcpyDeviceToHost));
    g->stats.texels_fetched = cnt[ST_FETCHED]; g->stats.texels_nominal = cnt[ST_NOMINAL]; g->stats.candidates = cnt[ST_CANDS];
    const bool dbg = getenv("TSB_DEBUG_PHASES") != nullptr && g->mgs_rank == 0;
    for (auto& c : chunks) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c.ev_t0, c.ev_done);
        g->stats.gpu_ms_resolve += ms;
        if (dbg) {
            float t0 = 0.f, a0 = 0.f, a1 = 0.f;
            cudaEventElapsedTime(&t0, ev_begin, c.ev_t0);
            cudaEventElapsedTime(&a0, ev_begin, c.ev_a0);
            cudaEventElapsedTime(&a1, ev_begin, c.ev_ready);
            fprintf(stderr, "[tsb] chunk stage %d %s%s first=%zu n=%zu | analysis %.3f..%.3f ms | resolve %.3f..%.3f ms (%.3f)\n", c.stage, c.redo ? "redo" : "new", c.sharded ? " (sharded)" : "",
                    c.first, c.items(), a0, a1, t0, t0 + ms, ms);
        }
    }
    float ms_a = 0.f, ms_total = 0.f;
    cudaEventElapsedTime(&ms_a, ev_a0, ev_a1);
    cudaEventElapsedTime(&ms_total, ev_begin, ev_end);
    g->stats.gpu_ms_analysis = ms_a;  // overlapped with the resolve kernels
    if (dbg) {
        float tp = 0.f;
        cudaEventElapsedTime(&tp, ev_begin, ev_a0);
        fprintf(stderr, "[tsb] pixel order + plan: %.3f ms on the device, %.3f ms host wall until everything was enqueued\n", tp, t_enqueued - t_start);
    }
    if (dbg && cnt[ST_ITEMS])
        fprintf(stderr, "[tsb] analysis stream %.3f ms (overlapped), total %.3f ms | cycles/item: wait %.0f lists %.0f neigh %.0f rand %.0f score %.0f commit %.0f (items %llu)\n",
                ms_a, ms_total, (double)cnt[ST_CYC_READY] / cnt[ST_ITEMS], (double)cnt[ST_CYC_KNN] / cnt[ST_ITEMS], (double)cnt[ST_CYC_NEIGH] / cnt[ST_ITEMS],
                (double)cnt[ST_CYC_WEIGHT] / cnt[ST_ITEMS], (double)cnt[ST_CYC_SCORE] / cnt[ST_ITEMS], (double)cnt[ST_CYC_COMMIT] / cnt[ST_ITEMS], cnt[ST_ITEMS]);
    g->stats.work_items = total_items;
    g->stats.gpu_ms_total = ms_total;
    g->stats.wall_ms_total = now_ms() - t_start;
    g->stats.gpu_ms_other = ms_total - g->stats.gpu_ms_resolve;
    g->have_loaded_points = false;
    g->state_init_opaque = state_opaque;
    return 0;
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* tsb_last_error(void) { return g_err.c_str(); }

int tsb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int tsb_resize(const uint8_t* rgba, uint32_t w, uint32_t h, uint8_t* out, uint32_t nw, uint32_t nh, int filter) {
    if (!rgba || !out || w == 0 || h == 0) return fail(TSB_ERR_INVALID, "tsb_resize: empty image");
    if (filter < 0 || filter > 2) return fail(TSB_ERR_INVALID, "tsb_resize: unknown filter");
    if (nw == 0 || nh == 0) return 0;
    cudaStream_t s = nullptr;
    DevBuf<uint32_t> src, tmp, dst;
    TRY(src.upload((const uint32_t*)rgba, (size_t)w * h, s));
    TRY(tmp.ensure((size_t)w * nh));
    TRY(dst.ensure((size_t)nw * nh));
    TRY(device_resize(src.p, (int)w, (int)h, dst.p, (int)nw, (int)nh, filter, tmp.p, s));
    CU(cudaMemcpy(out, dst.p, (size_t)nw * nh * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// utils::transform_to_guide_map (utils.rs:101-116): blur(sigma) -> grayscale -> RGBA (l,l,l,255)
int tsb_guide_map(const uint8_t* rgba, uint32_t w, uint32_t h, float sigma, uint8_t* out) {
    if (!rgba || !out || w == 0 || h == 0) return fail(TSB_ERR_INVALID, "tsb_guide_map: empty image");
    if (sigma < 0.0f) sigma = 1.0f;
    cudaStream_t s = nullptr;
    const size_t n = (size_t)w * h;
    DevBuf<uint32_t> src, tmp, dst;
    TRY(src.upload((const uint32_t*)rgba, n, s));
    TRY(tmp.ensure(n)); TRY(dst.ensure(n));
    TRY(device_resize(src.p, (int)w, (int)h, dst.p, (int)w, (int)h, FILTER_BLUR, tmp.p, s, sigma));
    k_grayscale<<<(uint32_t)((n + 255) / 256), 256, 0, s>>>(dst.p, (uint32_t)n);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, dst.p, n * 4, cudaMemcpyDeviceToHost));
    return 0;
}

// utils::match_histograms (utils.rs:135-183): `source` is remapped so that its R-channel CDF follows the target's
int tsb_match_histograms(const uint8_t* source, uint32_t sw, uint32_t sh, const uint8_t* target, uint32_t tw, uint32_t th, uint8_t* out) {
    if (!source || !target || !out || !sw || !sh || !tw || !th) return fail(TSB_ERR_INVALID, "tsb_match_histograms: empty image");
    cudaStream_t s = nullptr;
    const size_t ns = (size_t)sw * sh, nt = (size_t)tw * th;
    DevBuf<uint32_t> src, tgt, hist, lut;
    TRY(src.upload((const uint32_t*)source, ns, s));
    TRY(tgt.upload((const uint32_t*)target, nt, s));
    TRY(hist.ensure(512)); TRY(lut.ensure(256));
    CU(cudaMemsetAsync(hist.p, 0, 512 * 4, s));
    k_histogram_r<<<(uint32_t)((nt + 255) / 256), 256, 0, s>>>(tgt.p, (uint32_t)nt, hist.p);
    k_histogram_r<<<(uint32_t)((ns + 255) / 256), 256, 0, s>>>(src.p, (uint32_t)ns, hist.p + 256);
    CU(cudaGetLastError());
    uint32_t h[512];
    CU(cudaMemcpy(h, hist.p, sizeof(h), cudaMemcpyDeviceToHost));
    float tc[256], sc[256];
    auto cdf = [](const uint32_t* hh, float* o) {  // get_cdf, utils.rs:165-183
        for (int i = 0; i < 256; ++i) o[i] = i ? o[i - 1] + (float)hh[i] : (float)hh[i];
        float mx = o[255];
        for (int i = 0; i < 256; ++i) o[i] /= mx;
    };
    cdf(h, tc);
    cdf(h + 256, sc);
    uint32_t l[256];
    for (int v = 0; v < 256; ++v) {
        int pos = -1;
        for (int i = 0; i < 256; ++i) if (tc[i] > sc[v]) { pos = i; break; }
        unsigned nv = pos >= 0 ? (unsigned)pos : (unsigned)(uint8_t)(v + 1);   // unwrap_or((pixel_value + 1) as usize) as u8
        l[v] = (uint32_t)(uint8_t)((uint8_t)nv - 1);                             // `- 1` in u8 (wraps in release builds)
    }
    CU(cudaMemcpy(lut.p, l, sizeof(l), cudaMemcpyHostToDevice));
    k_apply_lut_r<<<(uint32_t)((ns + 255) / 256), 256, 0, s>>>(src.p, (uint32_t)ns, lut.p);
    CU(cudaGetLastError());
    CU(cudaMemcpy(out, src.p, ns * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int tsb_pyramid_build(const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t levels, uint8_t* out) {
    if (!rgba || !out || w == 0 || h == 0) return fail(TSB_ERR_INVALID, "tsb_pyramid_build: empty image");
    if (levels == 0) levels = 1;  // (1..0) is empty: only the original is pushed (img_pyramid.rs:25,35)
    cudaStream_t s = nullptr;
    const size_t img = (size_t)w * h;
    DevBuf<uint32_t> src, small, tmp, lvl;
    TRY(src.upload((const uint32_t*)rgba, img, s));
    TRY(small.ensure(img)); TRY(tmp.ensure(img)); TRY(lvl.ensure(img));
    size_t n = 0;
    for (uint32_t i = levels - 1; i >= 1; --i) {
        if (i >= 32) return fail(TSB_ERR_INVALID, "too many pyramid levels");
        uint32_t p = 1u << i;
        int sw = (int)(w / p), sh = (int)(h / p);
        if (sw == 0 || sh == 0) return fail(TSB_ERR_INVALID, "image too small for %u pyramid levels (the reference would produce an empty image)", levels);
        TRY(device_resize(src.p, (int)w, (int)h, small.p, sw, sh, TSB_FILTER_GAUSSIAN, tmp.p, s));
        TRY(device_resize(small.p, sw, sh, lvl.p, (int)w, (int)h, TSB_FILTER_GAUSSIAN, tmp.p, s));
        CU(cudaMemcpy(out + n * img * 4, lvl.p, img * 4, cudaMemcpyDeviceToHost));
        ++n;
    }
    memcpy(out + n * img * 4, rgba, img * 4);
    return 0;
}

int tsb_generator_create(const tsb_generator_desc* desc, tsb_generator** out) {
    if (!desc || !out) return fail(TSB_ERR_INVALID, "null argument");
    if (desc->out_width == 0 || desc->out_height == 0) return fail(TSB_ERR_INVALID, "empty output size");
    if (desc->out_width > 16384 || desc->out_height > 16384) return fail(TSB_ERR_UNSUPPORTED, "output dimensions above 16384 are not supported");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (ndev == 0) return fail(TSB_ERR_CUDA, "no CUDA device");
    tsb_generator* g = new tsb_generator();
    int dev = desc->device;
    if (dev < 0) { if (cudaGetDevice(&dev) != cudaSuccess) dev = 0; }
    g->device = dev;
    auto bail = [&](int code) { delete g; return code; };
    if (cudaSetDevice(dev) != cudaSuccess) return bail(fail(TSB_ERR_CUDA, "cudaSetDevice(%d) failed", dev));
    // (stream priorities were measured -- resolve or analysis stream at the highest priority: 46.5 / 46.2 vs 46.1 ms per 2048^2
    // step; the persistent resolve CTAs keep their SMs either way)
    if (cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(TSB_ERR_CUDA, "stream creation failed"));
    if (cudaStreamCreateWithFlags(&g->stream2, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(TSB_ERR_CUDA, "stream creation failed"));
    if (cudaStreamCreateWithFlags(&g->stream3, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(TSB_ERR_CUDA, "stream creation failed"));
    if (cudaHostAlloc((void**)&g->h_progress, 64, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&g->d_progress, g->h_progress, 0) != cudaSuccess) return bail(fail(TSB_ERR_CUDA, "mapped allocation failed"));
    if (cudaMallocHost((void**)&g->h_ctrl, 64) != cudaSuccess) return bail(fail(TSB_ERR_CUDA, "pinned allocation failed"));
    g->W = (int)desc->out_width; g->H = (int)desc->out_height;
    const size_t npix = (size_t)g->W * g->H;
    // mask geometry: room for the tiling mirror copies on every side (ms.rs:308-327)
    int x_l = (int)((float)g->W * 0.05f), y_b = (int)((float)g->H * 0.05f);
    g->mx = ((x_l + 1 + 31) / 32) * 32; g->my = y_b + 1;
    g->wpr = (g->W + 2 * g->mx + 31) / 32; g->mrows = g->H + 2 * g->my;
    g->wpr1 = (g->wpr + 31) / 32;
    int rc = 0;
    if ((rc = g->d_state.ensure(npix)) || (rc = g->d_score.ensure(npix)) || (rc = g->d_mask.ensure((size_t)g->wpr * g->mrows)) ||
        (rc = g->d_mask1.ensure((size_t)g->wpr1 * g->mrows))) return bail(rc);
    {   // normalised coordinates of ms.rs:405-415 (an IEEE division each) as tables over every coordinate a point can have
        std::vector<double> dx((size_t)g->wpr * 32), dy((size_t)g->mrows);
        for (size_t i = 0; i < dx.size(); ++i) dx[i] = (double)((int)i - g->mx) / (double)g->W;
        for (size_t i = 0; i < dy.size(); ++i) dy[i] = (double)((int)i - g->my) / (double)g->H;
        if ((rc = g->d_divx.upload(dx.data(), dx.size(), g->stream)) || (rc = g->d_divy.upload(dy.data(), dy.size(), g->stream))) return bail(rc);
        if (cudaStreamSynchronize(g->stream) != cudaSuccess) return bail(fail(TSB_ERR_CUDA, "table upload failed"));
    }
    SpiralHost sp = build_spiral(SPIRAL_RT);
    g->spiralN = (int)sp.off.size(); g->RT2 = sp.RT2;
    if ((rc = g->d_spiral.upload(sp.off.data(), sp.off.size(), g->stream)) || (rc = g->d_cntLE.upload(sp.cntLE.data(), sp.cntLE.size(), g->stream))) return bail(rc);
    if (desc->inpaint_mask) {
        if (!desc->inpaint_color) return bail(fail(TSB_ERR_INVALID, "inpaint_color is required with inpaint_mask"));
        g->inpaint = true;
        g->inpaint_index = desc->inpaint_example_index;
        if ((rc = g->d_inp_mask.upload((const uint32_t*)desc->inpaint_mask, npix, g->stream)) ||
            (rc = g->d_inp_color.upload((const uint32_t*)desc->inpaint_color, npix, g->stream))) return bail(rc);
        for (size_t i = 0; i < npix; ++i) {  // ms.rs:271-279
            if (desc->inpaint_mask[i * 4] < 255) g->unresolved0.push_back((uint32_t)i);
            else { g->resolved0.push_back((uint32_t)i); if (desc->inpaint_color[i * 4 + 3] != 255) g->inpaint_opaque = false; }
        }
    } else {
        g->unresolved0.resize(npix);
        for (size_t i = 0; i < npix; ++i) g->unresolved0[i] = (uint32_t)i;
    }
    cudaFuncSetAttribute(k_eval_items<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtaSmem));
    cudaFuncSetAttribute(k_eval_items<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtaSmem));
#define TSB_STREAM_ATTR(G, O, R)                                                                                          \
    cudaFuncSetAttribute(k_stream<G, O, R, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StreamSmem)); \
    cudaFuncSetAttribute(k_stream<G, O, R, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StreamSmem))
    TSB_STREAM_ATTR(false, false, false); TSB_STREAM_ATTR(false, true, false); TSB_STREAM_ATTR(false, false, true); TSB_STREAM_ATTR(false, true, true);
    TSB_STREAM_ATTR(true, false, false); TSB_STREAM_ATTR(true, true, false); TSB_STREAM_ATTR(true, false, true); TSB_STREAM_ATTR(true, true, true);
#undef TSB_STREAM_ATTR
    cudaFuncSetAttribute(k_lists_chunk<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(KnnScratch) * WARPS_PER_CTA));
    cudaFuncSetAttribute(k_lists_chunk<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(KnnScratch) * WARPS_PER_CTA));
    cudaFuncSetAttribute(k_weights, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(KW_WARPS * KMAX * (KW_ITEMS + 1) * sizeof(double)));
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return bail(fail(TSB_ERR_CUDA, "cudaGetDeviceProperties failed"));
    g->n_sms = prop.multiProcessorCount;
    if (!getenv("TSB_NO_L2_WINDOW") && prop.persistingL2CacheMaxSize > 0) {
        // room for one framed example level (north_star: "keeps the active example level resident in an L2 persisting window")
        size_t want = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, 32u << 20);
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) g->l2_persist_bytes = want; else cudaGetLastError();
    }
    {
        int ps = 0, psg = 0, pr = 0, pe = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ps, k_stream<false, false, true>, CTA_THREADS, sizeof(StreamSmem)) != cudaSuccess || ps < 1) ps = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&psg, k_stream<true, false, true>, CTA_THREADS, sizeof(StreamSmem)) != cudaSuccess || psg < 1) psg = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pr, k_lists_chunk<false>, CTA_THREADS, sizeof(KnnScratch) * WARPS_PER_CTA) != cudaSuccess || pr < 1) pr = 4;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pe, k_eval_items<false>, CTA_THREADS, sizeof(CtaSmem)) != cudaSuccess || pe < 1) pe = 1;
        g->max_ctas_stream = prop.multiProcessorCount * ps;         // persistent grids: co-resident CTAs only
        g->max_ctas_stream_guided = prop.multiProcessorCount * psg;
        g->max_ctas_radius = prop.multiProcessorCount * pr;         // one wave of co-resident CTAs, grid-stride over the items
        g->max_ctas = prop.multiProcessorCount * pe;
    }
    if ((rc = init_state(g))) return bail(rc);
    if (cudaStreamSynchronize(g->stream) != cudaSuccess) return bail(fail(TSB_ERR_CUDA, "generator initialisation failed: %s", cudaGetErrorString(cudaGetLastError())));
    *out = g;
    return 0;
}

void tsb_generator_destroy(tsb_generator* g) {
    if (!g) return;
    cudaSetDevice(g->device);
    for (void* p : g->mg_opened) cudaIpcCloseMemHandle(p);
    delete g;
}

int tsb_generator_reset(tsb_generator* g) {
    if (!g) return fail(TSB_ERR_INVALID, "null generator");
    TRY(set_device(g));
    TRY(init_state(g));
    CU(cudaStreamSynchronize(g->stream));
    return 0;
}

int tsb_generator_random_init(tsb_generator* g, uint64_t count, const tsb_image* top, uint32_t n, uint64_t seed) {
    if (!g || !top || n == 0) return fail(TSB_ERR_INVALID, "null argument");
    TRY(set_device(g));
    cudaStream_t s = g->stream;
    // upload the images (pyramid[len-1] of EVERY example, session.rs:42-48)
    std::vector<DevBuf<uint32_t>> imgs(n);
    std::vector<DevEx> desc(n);
    for (uint32_t e = 0; e < n; ++e) {
        TRY(imgs[e].upload((const uint32_t*)top[e].rgba, (size_t)top[e].width * top[e].height, s));
        desc[e].px = imgs[e].p; desc[e].smask = nullptr; desc[e].w = (int)top[e].width; desc[e].h = (int)top[e].height;
    }
    DevBuf<DevEx> d_desc;
    TRY(d_desc.upload(desc.data(), n, s));
    std::vector<uint32_t> items;
    for (uint64_t i = 0; i < count; ++i) {  // ms.rs:433-443
        if (g->unresolved.empty()) continue;
        size_t idx = (size_t)Pcg32::seed_from_u64(seed + i).gen_range_usize((uint64_t)g->unresolved.size());
        uint32_t flat = g->unresolved[idx];
        g->unresolved[idx] = g->unresolved.back();
        g->unresolved.pop_back();
        uint64_t s2 = seed + i + (uint64_t)flat;
        uint32_t rmap = (uint32_t)Pcg32::seed_from_u64(s2).gen_range_usize((uint64_t)n);
        uint32_t rx = Pcg32::seed_from_u64(s2).gen_range_u32(top[rmap].width);
        uint32_t ry = Pcg32::seed_from_u64(s2).gen_range_u32(top[rmap].height);
        items.push_back(flat); items.push_back(rx); items.push_back(ry); items.push_back(rmap);
        if (top[rmap].rgba[((size_t)ry * top[rmap].width + rx) * 4 + 3] != 255) g->state_init_opaque = false;
        g->resolved_order.push_back(flat);
    }
    g->locked += (size_t)count;  // ms.rs:444
    if (!items.empty()) {
        StageDev S;
        fill_stage_geometry(g, S, false);
        DevBuf<uint32_t> d_items;
        TRY(d_items.upload(items.data(), items.size(), s));
        uint32_t ni = (uint32_t)(items.size() / 4);
        k_commit_fixed<<<(ni + 127) / 128, 128, 0, s>>>(S, d_desc.p, d_items.p, ni, 0);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(s));
    }
    return 0;
}

int tsb_generator_upload_inputs(tsb_generator* g, const tsb_pyramid* examples, uint32_t n_examples, const tsb_guides* guides, const tsb_sampling* sampling) {
    if (!g) return fail(TSB_ERR_INVALID, "null generator");
    TRY(set_device(g));
    return upload_inputs(g, examples, n_examples, guides, sampling);
}

static int resolve_dispatch(tsb_generator* g, const tsb_params* params, tsb_progress_fn cb, void* user) {
    return resolve_stream(g, params, cb, user);
}

int tsb_generator_resolve_resident(tsb_generator* g, const tsb_params* params, tsb_progress_fn cb, void* user) {
    if (!g || !params) return fail(TSB_ERR_INVALID, "null argument");
    return resolve_dispatch(g, params, cb, user);
}

int tsb_generator_resolve(tsb_generator* g, const tsb_params* params, const tsb_pyramid* examples, uint32_t n_examples,
                          const tsb_guides* guides, const tsb_sampling* sampling, tsb_progress_fn cb, void* user) {
    if (!g || !params) return fail(TSB_ERR_INVALID, "null argument");
    TRY(set_device(g));
    TRY(upload_inputs(g, examples, n_examples, guides, sampling));
    return resolve_dispatch(g, params, cb, user);
}

static int read_unpacked(tsb_generator* g, uint32_t* color, uint32_t* coord, uint32_t* idm) {
    TRY(set_device(g));
    const size_t npix = (size_t)g->W * g->H;
    DevBuf<uint32_t>& dc = g->d_read_color;
    DevBuf<uint32_t>& dco = g->d_read_coord;
    DevBuf<uint32_t>& di = g->d_read_id;
    if (color) TRY(dc.ensure(npix));
    if (coord) TRY(dco.ensure(npix * 3));
    if (idm) TRY(di.ensure(npix * 2));
    k_unpack_state<<<(uint32_t)((npix + 255) / 256), 256, 0, g->stream>>>(g->d_state.p, (uint32_t)npix, color ? dc.p : nullptr,
                                                                           coord ? dco.p : nullptr, idm ? di.p : nullptr);
    CU(cudaGetLastError());
    if (color) CU(cudaMemcpyAsync(color, dc.p, npix * 4, cudaMemcpyDeviceToHost, g->stream));
    if (coord) CU(cudaMemcpyAsync(coord, dco.p, npix * 12, cudaMemcpyDeviceToHost, g->stream));
    if (idm) CU(cudaMemcpyAsync(idm, di.p, npix * 8, cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    return 0;
}

int tsb_generator_read_color(tsb_generator* g, uint8_t* rgba) {
    if (!g || !rgba) return fail(TSB_ERR_INVALID, "null argument");
    return read_unpacked(g, (uint32_t*)rgba, nullptr, nullptr);
}
int tsb_generator_read_coord(tsb_generator* g, uint32_t* xym) {
    if (!g || !xym) return fail(TSB_ERR_INVALID, "null argument");
    return read_unpacked(g, nullptr, xym, nullptr);
}
int tsb_generator_read_id(tsb_generator* g, uint32_t* pm) {
    if (!g || !pm) return fail(TSB_ERR_INVALID, "null argument");
    return read_unpacked(g, nullptr, nullptr, pm);
}
int tsb_generator_resolved_count(tsb_generator* g, uint64_t* n, uint64_t* locked) {
    if (!g) return fail(TSB_ERR_INVALID, "null generator");
    if (n) *n = g->resolved_order.size();
    if (locked) *locked = g->locked;
    return 0;
}
int tsb_generator_read_resolved(tsb_generator* g, uint32_t* flat, float* score) {
    if (!g) return fail(TSB_ERR_INVALID, "null generator");
    TRY(set_device(g));
    size_t n = g->resolved_order.size();
    if (flat) memcpy(flat, g->resolved_order.data(), n * 4);
    if (score && n) {
        DevBuf<uint32_t> df;
        DevBuf<float> ds;
        TRY(df.upload(g->resolved_order.data(), n, g->stream));
        TRY(ds.ensure(n));
        k_gather_scores<<<(uint32_t)((n + 255) / 256), 256, 0, g->stream>>>(g->d_score.p, df.p, (uint32_t)n, ds.p);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(score, ds.p, n * 4, cudaMemcpyDeviceToHost, g->stream));
        CU(cudaStreamSynchronize(g->stream));
    }
    return 0;
}
int tsb_generator_read_uncertainty(tsb_generator* g, uint8_t* rgba) {
    if (!g || !rgba) return fail(TSB_ERR_INVALID, "null argument");
    TRY(set_device(g));
    const size_t npix = (size_t)g->W * g->H;
    StageDev S;
    fill_stage_geometry(g, S, false);
    DevBuf<uint32_t> d;
    TRY(d.ensure(npix));
    k_uncertainty<<<(uint32_t)((npix + 255) / 256), 256, 0, g->stream>>>(S, d.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(rgba, d.p, npix * 4, cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    return 0;
}
int tsb_generator_read_id_maps(tsb_generator* g, uint8_t* patch_rgba, uint8_t* map_rgba) {
    if (!g || !patch_rgba || !map_rgba) return fail(TSB_ERR_INVALID, "null argument");
    TRY(set_device(g));
    const size_t npix = (size_t)g->W * g->H;
    DevBuf<uint32_t> a, b;
    TRY(a.ensure(npix)); TRY(b.ensure(npix));
    k_id_maps<<<(uint32_t)((npix + 255) / 256), 256, 0, g->stream>>>(g->d_state.p, (uint32_t)npix, a.p, b.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(patch_rgba, a.p, npix * 4, cudaMemcpyDeviceToHost, g->stream));
    CU(cudaMemcpyAsync(map_rgba, b.p, npix * 4, cudaMemcpyDeviceToHost, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    return 0;
}
int tsb_generator_get_stats(tsb_generator* g, tsb_stats* out) {
    if (!g || !out) return fail(TSB_ERR_INVALID, "null argument");
    *out = g->stats;
    return 0;
}

// ---- test-only ---------------------------------------------------------------------------------
int tsb_generator_load_state(tsb_generator* g, const uint8_t* color, const uint32_t* coord, const uint32_t* idm,
                             const int32_t* tree_xy, uint64_t n_tree, const uint32_t* resolved_flat,
                             const float* resolved_score, uint64_t n_resolved, uint64_t locked) {
    if (!g || !color || !coord || !idm) return fail(TSB_ERR_INVALID, "null argument");
    TRY(set_device(g));
    cudaStream_t s = g->stream;
    const size_t npix = (size_t)g->W * g->H;
    DevBuf<uint32_t> dc, dco, di;
    TRY(dc.upload((const uint32_t*)color, npix, s));
    TRY(dco.upload(coord, npix * 3, s));
    TRY(di.upload(idm, npix * 2, s));
    k_pack_state<<<(uint32_t)((npix + 255) / 256), 256, 0, s>>>(g->d_state.p, (uint32_t)npix, dc.p, dco.p, di.p);
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(g->d_mask.p, 0, (size_t)g->wpr * g->mrows * 4, s));
    CU(cudaMemsetAsync(g->d_mask1.p, 0, (size_t)g->wpr1 * g->mrows * 4, s));
    StageDev S;
    fill_stage_geometry(g, S, false);
    DevBuf<uint32_t> dp;
    if (n_tree) {
        TRY(dp.upload((const uint32_t*)tree_xy, n_tree * 2, s));
        k_mask_insert_points<<<(uint32_t)((n_tree + 255) / 256), 256, 0, s>>>(S, (const int32_t*)dp.p, (uint32_t)n_tree);
        CU(cudaGetLastError());
    }
    g->loaded_points.assign(tree_xy, tree_xy + n_tree * 2);
    g->have_loaded_points = true;
    g->resolved_order.assign(resolved_flat, resolved_flat + n_resolved);
    g->locked = (size_t)locked;
    if (n_resolved && resolved_score) {
        DevBuf<uint32_t> df;
        DevBuf<float> ds;
        TRY(df.upload(resolved_flat, n_resolved, s));
        TRY(ds.upload(resolved_score, n_resolved, s));
        k_scatter_scores<<<(uint32_t)((n_resolved + 255) / 256), 256, 0, s>>>(g->d_score.p, df.p, ds.p, (uint32_t)n_resolved);
        CU(cudaGetLastError());
    }
    // unresolved = every pixel not in the resolved list, in ascending order (order only matters for later picks)
    std::vector<uint8_t> isres(npix, 0);
    g->state_init_opaque = true;
    for (uint64_t i = 0; i < n_resolved; ++i) { isres[resolved_flat[i]] = 1; if (color[(size_t)resolved_flat[i] * 4 + 3] != 255) g->state_init_opaque = false; }
    if (n_resolved) {
        DevBuf<uint32_t> df;
        TRY(df.upload(resolved_flat, n_resolved, s));
        k_tag_locked<<<(uint32_t)((n_resolved + 255) / 256), 256, 0, s>>>(g->d_state.p, df.p, (uint32_t)n_resolved);
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(s));
    }
    g->unresolved.clear();
    for (size_t i = 0; i < npix; ++i) if (!isres[i]) g->unresolved.push_back((uint32_t)i);
    CU(cudaStreamSynchronize(s));
    return 0;
}

int tsb_generator_eval_items(tsb_generator* g, const tsb_params* prm, int32_t level, float adaptive_alpha, uint64_t p_stage_seed,
                             uint32_t n, const uint32_t* pixel_flat, const uint64_t* loop_seed, int32_t* neigh, int32_t* res, float* score) {
    if (!g || !prm || !pixel_flat || !loop_seed || !neigh || !res || !score) return fail(TSB_ERR_INVALID, "null argument");
    if (!g->inputs_ready) return fail(TSB_ERR_INVALID, "inputs have not been uploaded");
    TRY(check_params(g, prm));
    TRY(set_device(g));
    (void)p_stage_seed;
    if (level < 0 || level >= g->n_levels) return fail(TSB_ERR_INVALID, "level out of range");
    cudaStream_t s = g->stream;
    const int k = (int)prm->nearest_neighbors, m = (int)prm->random_sample_locations;
    TRY(g->d_luts.ensure(512));
    TRY(decide_opaque(g, level));
    StageDev S;
    fill_stage_geometry(g, S, prm->tiling_mode != 0);
    stage_inputs(g, S, level, prm);
    TRY(upload_luts(g, prm, adaptive_alpha));
    S.seq_exact = g->luts_exact_mode ? 1 : 0;
    if (S.seq_exact) S.opaque = 0;
    S.r2_hint = r2_hint_for(g, g->resolved_order.size(), (uint32_t)k);
    DevBuf<uint32_t> dpix, dxy;
    DevBuf<uint8_t> dmap;
    DevBuf<int32_t> dneigh, dres;
    DevBuf<float> dscore;
    TRY(dpix.upload(pixel_flat, n, s));
    TRY(dxy.ensure((size_t)n * m)); TRY(dmap.ensure((size_t)n * m));
    TRY(dneigh.ensure((size_t)n * 2 * k)); TRY(dres.ensure((size_t)n * 8)); TRY(dscore.ensure(n));
    // items carry arbitrary loop seeds: generate their random candidates one launch per run of consecutive seeds
    uint32_t i = 0;
    while (i < n) {
        uint32_t j = i + 1;
        while (j < n && loop_seed[j] == loop_seed[j - 1] + 1) ++j;
        const unsigned rb = m <= 64 ? 128u : 32u;
        if (S.n_ex == 1) k_rand_candidates<true><<<(j - i + rb - 1) / rb, rb, (size_t)rb * m * 5 + rb, s>>>(S.ex, S.n_ex, m, loop_seed[i] + 1ull, j - i, dxy.p + (size_t)i * m, dmap.p + (size_t)i * m);
        else k_rand_candidates<false><<<(j - i + rb - 1) / rb, rb, (size_t)rb * m * 5 + rb, s>>>(S.ex, S.n_ex, m, loop_seed[i] + 1ull, j - i, dxy.p + (size_t)i * m, dmap.p + (size_t)i * m);
        CU(cudaGetLastError());
        i = j;
    }
    const int grid = std::max(1, std::min((int)((n + WARPS_PER_CTA - 1) / WARPS_PER_CTA), g->max_ctas));
    if (g->guided) k_eval_items<true><<<grid, CTA_THREADS, sizeof(CtaSmem), s>>>(S, n, dpix.p, dxy.p, dmap.p, dneigh.p, dres.p, dscore.p);
    else k_eval_items<false><<<grid, CTA_THREADS, sizeof(CtaSmem), s>>>(S, n, dpix.p, dxy.p, dmap.p, dneigh.p, dres.p, dscore.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(neigh, dneigh.p, (size_t)n * 2 * k * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(res, dres.p, (size_t)n * 8 * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaMemcpyAsync(score, dscore.p, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    // items without any resolved neighbour take the resolve_at_random path (ms.rs:1002-1009): host draws
    for (uint32_t t = 0; t < n; ++t) {
        int32_t* ro = res + (size_t)t * 8;
        if (ro[7]) {
            uint32_t rmap = (uint32_t)Pcg32::seed_from_u64(p_stage_seed).gen_range_usize((uint64_t)g->n_ex);
            int e = g->filt[rmap];
            ro[3] = (int32_t)Pcg32::seed_from_u64(p_stage_seed).gen_range_u32((uint32_t)g->ex_w[e]);
            ro[4] = (int32_t)Pcg32::seed_from_u64(p_stage_seed).gen_range_u32((uint32_t)g->ex_h[e]);
            ro[5] = (int32_t)rmap; ro[6] = (int32_t)pixel_flat[t]; ro[1] = 0; ro[2] = 0;
            score[t] = 0.f;
        }
    }
    return 0;
}

// ---- band-sharded multi-GPU execution: one process per GPU, replicas linked through CUDA IPC ----------------
// Shared with the peers: both state buffers (a neighbour outside this rank's band is read in its owner's replica), the
// first-resolution scores (gathered at the end of a run) and the barrier flag block.
enum { MG_STATE_A = 0, MG_STATE_B, MG_SCORE, MG_SYNC, MG_NBUF };
enum { MG_ENTRY = 80 };  // bytes per exported buffer: IPC handle (64) + offset inside the allocation block (8) + padding

static void* mg_local_ptr(tsb_generator* g, int which) {
    switch (which) {
    case MG_STATE_A: return g->d_state.p;
    case MG_STATE_B: return g->d_state2.p;
    case MG_SCORE: return g->d_score.p;
    default: return g->d_mgs_sync.p;
    }
}

int tsb_generator_mg_prepare(tsb_generator* g, const tsb_params* params, uint32_t* n_handles) {
    if (!g || !params || !n_handles) return fail(TSB_ERR_INVALID, "null argument");
    TRY(set_device(g));
    // final sizes: the IPC mappings must stay valid for the lifetime of the generator
    TRY(g->d_state2.ensure((size_t)g->W * g->H));
    TRY(g->d_mgs_sync.ensure(64));
    CU(cudaMemsetAsync(g->d_mgs_sync.p, 0, 64 * 4, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    g->mgs_seq = 0;
    *n_handles = MG_NBUF;
    return 0;
}

// An IPC handle names the whole underlying allocation block and cudaIpcOpenMemHandle returns the BASE of that
// block; cudaMalloc packs small buffers into shared blocks, so every exported entry carries the buffer's offset
// inside its block (driver API cuMemGetAddressRange, resolved at run time so that the library has no link-time
// dependency on libcuda).  Entry layout: 64-byte handle + 8-byte offset + 8 bytes padding = 80 bytes.
static int ipc_offset_of(void* ptr, uint64_t* off) {
    typedef int (*fn_t)(unsigned long long*, size_t*, unsigned long long);
    static fn_t fn = nullptr;
    if (!fn) {
        void* h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
        if (!h) return fail(TSB_ERR_CUDA, "libcuda.so.1 not found");
        fn = (fn_t)dlsym(h, "cuMemGetAddressRange_v2");
        if (!fn) return fail(TSB_ERR_CUDA, "cuMemGetAddressRange_v2 not found");
    }
    unsigned long long base = 0;
    size_t size = 0;
    int rc = fn(&base, &size, (unsigned long long)(uintptr_t)ptr);
    if (rc != 0) return fail(TSB_ERR_CUDA, "cuMemGetAddressRange failed (%d)", rc);
    *off = (uint64_t)((unsigned long long)(uintptr_t)ptr - base);
    return 0;
}

int tsb_generator_mg_export(tsb_generator* g, uint8_t* handles) {
    if (!g || !handles) return fail(TSB_ERR_INVALID, "null argument");
    TRY(set_device(g));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!g->d_state2.p || !g->d_mgs_sync.p) return fail(TSB_ERR_INVALID, "tsb_generator_mg_prepare has not been called");
    for (int i = 0; i < MG_NBUF; ++i) {
        uint64_t off = 0;
        TRY(ipc_offset_of(mg_local_ptr(g, i), &off));
        memcpy(handles + (size_t)i * MG_ENTRY + 64, &off, 8);
        cudaIpcMemHandle_t h;
        cudaError_t e = cudaIpcGetMemHandle(&h, mg_local_ptr(g, i));
        if (e != cudaSuccess) return fail(TSB_ERR_CUDA, "cudaIpcGetMemHandle failed for shared buffer %d (%p): %s", i, mg_local_ptr(g, i), cudaGetErrorString(e));
        memcpy(handles + (size_t)i * MG_ENTRY, &h, 64);
    }
    return 0;
}

int tsb_generator_mg_attach(tsb_generator* g, uint32_t rank, uint32_t world, const uint8_t* all_handles, tsb_barrier_fn barrier, void* user) {
    if (!g || !all_handles) return fail(TSB_ERR_INVALID, "null argument");
    if (world < 2 || world > (uint32_t)MG_MAX || rank >= world) return fail(TSB_ERR_INVALID, "world size must be in [2,%d]", MG_MAX);
    TRY(set_device(g));
    g->mgs_rank = (int)rank; g->mgs_world = (int)world;
    g->mgs_band_h = (g->H + (int)world - 1) / (int)world;
    if (g->mgs_band_h < 8) return fail(TSB_ERR_UNSUPPORTED, "output too small for %u bands", world);
    for (uint32_t r = 0; r < world; ++r) {
        void* ptrs[MG_NBUF];
        for (int i = 0; i < MG_NBUF; ++i) {
            if (r == rank) { ptrs[i] = mg_local_ptr(g, i); continue; }
            cudaIpcMemHandle_t h;
            uint64_t off = 0;
            const uint8_t* entry = all_handles + ((size_t)r * MG_NBUF + i) * MG_ENTRY;
            memcpy(&h, entry, 64);
            memcpy(&off, entry + 64, 8);
            // several buffers of a peer may live in one allocation block: open every block once
            void* base = nullptr;
            for (auto& ob : g->mg_blocks) if (!memcmp(&ob.first, &h, 64)) base = ob.second;
            if (!base) {
                CU(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
                g->mg_blocks.push_back({h, base});
                g->mg_opened.push_back(base);
            }
            ptrs[i] = (uint8_t*)base + off;
        }
        g->mgs_A[r] = (uint4*)ptrs[MG_STATE_A]; g->mgs_B[r] = (uint4*)ptrs[MG_STATE_B];
        g->mgs_score[r] = (float*)ptrs[MG_SCORE]; g->mgs_sync[r] = (uint32_t*)ptrs[MG_SYNC];
    }
    TRY(g->d_mgs_sync_ptrs.upload(g->mgs_sync, MG_MAX, g->stream));
    CU(cudaStreamSynchronize(g->stream));
    (void)barrier; (void)user;  // the ranks synchronise on the device (k_mg_signal / k_mg_wait); kept for ABI compatibility
    if (getenv("TSB_MG_MIN_PHASE")) g->mgs_shard_min = (size_t)std::max(1, atoi(getenv("TSB_MG_MIN_PHASE")));
    g->mgs_on = true;
    return 0;
}

int tsb_generator_mg_phases(tsb_generator* g, uint64_t* n) {
    if (!g || !n) return fail(TSB_ERR_INVALID, "null argument");
    *n = g->mgs_sharded_chunks;
    return 0;
}

int tsb_generator_set_trace(tsb_generator* g, int on) {
    if (!g) return fail(TSB_ERR_INVALID, "null generator");
    g->trace = on != 0;
    return 0;
}
int tsb_generator_trace_count(tsb_generator* g, uint64_t* n) {
    if (!g || !n) return fail(TSB_ERR_INVALID, "null argument");
    *n = g->trace_n;
    return 0;
}
int tsb_generator_read_trace(tsb_generator* g, uint32_t* pixel, int32_t* best, int32_t* ncand, int32_t* nneigh, float* score) {
    if (!g) return fail(TSB_ERR_INVALID, "null generator");
    TRY(set_device(g));
    size_t n = (size_t)g->trace_n;
    if (!n) return 0;
    memcpy(pixel, g->tr_pixel.data(), n * 4);
    CU(cudaMemcpy(best, g->d_tr_best.p, n * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(ncand, g->d_tr_ncand.p, n * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(nneigh, g->d_tr_nneigh.p, n * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(score, g->d_tr_score.p, n * 4, cudaMemcpyDeviceToHost));
    for (uint64_t idx : g->tr_fix_idx) { best[idx] = -1; ncand[idx] = 0; nneigh[idx] = 0; score[idx] = 0.f; }
    return 0;
}

int tsb_microbench_gather(uint64_t bytes, int mode, int iters, double* useful_gbs, double* gathers_per_s) {
    int w = 512;
    int h = (int)(bytes / 4 / (uint64_t)w);
    if (h < 32) return fail(TSB_ERR_INVALID, "window too small");
    DevBuf<uint32_t> img, sink;
    std::vector<uint32_t> host((size_t)w * h);
    for (size_t i = 0; i < host.size(); ++i) host[i] = (uint32_t)(i * 2654435761u);
    TRY(img.upload(host.data(), host.size(), nullptr));
    TRY(sink.ensure(1));
    cudaDeviceProp prop;
    int dev = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaGetDeviceProperties(&prop, dev));
    const int threads = 256, blocks = prop.multiProcessorCount * 8 * 4, steps = 50;
    cudaTextureObject_t tex = 0;
    if (mode == 1) {
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypePitch2D;
        rd.res.pitch2D.devPtr = img.p;
        rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
        rd.res.pitch2D.width = (size_t)w; rd.res.pitch2D.height = (size_t)h; rd.res.pitch2D.pitchInBytes = (size_t)w * 4;
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        CU(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
    }
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    for (int it = -2; it < iters; ++it) {
        if (it == 0) CU(cudaEventRecord(e0));
        if (mode == 1) k_gather_bench_tex<<<blocks, threads>>>(tex, w, h, steps, (uint32_t)it, sink.p);
        else k_gather_bench<<<blocks, threads>>>(img.p, w, h, steps, (uint32_t)it, sink.p);
    }
    CU(cudaEventRecord(e1));
    CU(cudaEventSynchronize(e1));
    CU(cudaGetLastError());
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    double gathers = (double)iters * (double)blocks * threads * 8.0 * steps;
    if (gathers_per_s) *gathers_per_s = gathers / (ms * 1e-3);
    if (useful_gbs) *useful_gbs = gathers * 4.0 / (ms * 1e-3) / 1e9;
    if (tex) cudaDestroyTextureObject(tex);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}

}  // extern "C"
