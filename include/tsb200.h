/* tsb200.h -- C ABI of the B200-native texture-synthesis hot path.
 *
 * Drop-in boundary for EmbarkStudios/texture-synthesis (reference paths relative to its checkout):
 * the reference has no FFI of its own; the hot path sits behind three Rust call sites in
 * lib/src/session.rs -- `ImagePyramid::new` (lib.rs:570,575,581; session.rs:395),
 * `Generator::new/new_from_inpaint` (session.rs:426-434) and
 * `Generator::resolve_random_batch` + `Generator::resolve` (session.rs:50-61) -- and
 * `GeneratedImage` reads the finished maps back (lib.rs:378-455 -> ms.rs:605-684).
 * Each entry point below names the reference item it replaces.  INTEGRATION.md shows the Rust
 * `extern "C"` block and the patched call sites.
 *
 * Conventions: all images are tightly packed RGBA8, row major.  All pointers are HOST pointers
 * owned by the caller and borrowed for the duration of the call only.  Every function returns 0
 * on success or a negative tsb_status; tsb_last_error() returns a thread-local message.
 * There is no CPU fallback: without a usable CUDA device every compute entry point fails with
 * TSB_ERR_CUDA.
 */
#ifndef TSB200_H
#define TSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    TSB_OK = 0,
    TSB_ERR_INVALID = -1,     /* invalid parameter (same conditions as session.rs:450-524, plus size limits) */
    TSB_ERR_CUDA = -2,        /* CUDA runtime error / no device */
    TSB_ERR_UNSUPPORTED = -3, /* outside the limits of this implementation */
    TSB_ERR_INTERNAL = -4
} tsb_status;

/* GeneratorParams, ms.rs:18-42 (filled by Parameters::to_generator_params, lib.rs:362-374). */
typedef struct {
    uint32_t nearest_neighbors;       /* k, default 50 */
    uint32_t _pad0;
    uint64_t random_sample_locations; /* m, default 50 */
    float cauchy_dispersion;          /* default 1.0 */
    float p;                          /* backtrack_percent, default 0.5 */
    int32_t p_stages;                 /* backtrack_stages, default 5 */
    float alpha;                      /* guide_alpha, default 0.8 */
    uint64_t seed;
    uint64_t max_thread_count;        /* accepted for API compatibility; the schedule is always the 1-thread one */
    int32_t tiling_mode;
    int32_t _pad1;
} tsb_params;

/* One RgbaImage. */
typedef struct {
    const uint8_t* rgba;
    uint32_t width, height;
} tsb_image;

/* One ImagePyramid (img_pyramid.rs:2-4): n_levels images of identical size stored back to back,
 * level 0 = blurriest, level n_levels-1 = the original (`bottom()`). */
typedef struct {
    const uint8_t* levels;
    uint32_t width, height, n_levels;
} tsb_pyramid;

/* SamplingMethod, lib.rs:458-468.  For TSB_SAMPLE_IMAGE, `rgba` has the example's dimensions and a
 * pixel may be sampled iff its R channel != 0 (ms.rs:1546). */
typedef enum { TSB_SAMPLE_ALL = 0, TSB_SAMPLE_IGNORE = 1, TSB_SAMPLE_IMAGE = 2 } tsb_sample_kind;
typedef struct {
    int32_t kind;
    int32_t _pad;
    const uint8_t* rgba;
} tsb_sampling;

/* GuidesPyramidStruct, ms.rs:62-65: one target-guide pyramid at output size and one guide pyramid per
 * example (ALL examples, including ignored ones -- the reference does not filter them, ms.rs:67-81). */
typedef struct {
    tsb_pyramid target;
    const tsb_pyramid* examples;
    uint32_t n_examples;
    uint32_t _pad;
} tsb_guides;

/* Generator::new (ms.rs:220-234) when inpaint_mask == NULL, else Generator::new_from_inpaint
 * (ms.rs:236-293).  Mask and colour must already have the output size (the Triangle resize of
 * ms.rs:242-263 is available as tsb_resize). */
typedef struct {
    uint32_t out_width, out_height;
    const uint8_t* inpaint_mask;   /* RGBA; pixel kept (locked) iff R == 255 (ms.rs:272) */
    const uint8_t* inpaint_color;  /* RGBA initial colour map */
    uint32_t inpaint_example_index;
    int32_t device;                /* CUDA device ordinal, -1 = current device */
} tsb_generator_desc;

typedef struct tsb_generator tsb_generator;

/* GeneratorProgress::update (session.rs:528-558).  Called on the calling thread only, with a host
 * RGBA snapshot valid for the duration of the callback, when the integer percentage changes. */
typedef void (*tsb_progress_fn)(void* user, const uint8_t* rgba, uint32_t width, uint32_t height,
                                uint64_t total_current, uint64_t total_total,
                                uint64_t stage_current, uint64_t stage_total);

/* Run statistics of the last tsb_generator_resolve* call. */
typedef struct {
    uint64_t work_items;        /* pixel resolutions performed (sum over stages) */
    uint64_t candidates;        /* candidates scored */
    uint64_t texels_fetched;    /* example texels actually gathered by the scoring kernel (instrumented) */
    uint64_t texels_nominal;    /* sum over items of n_candidates * n_neighbours */
    uint64_t rounds;            /* resolve-kernel launches (dependency rounds) */
    uint64_t kernel_launches;   /* all kernels launched by the call */
    uint64_t phases;            /* dependency-analysis phases */
    double gpu_ms_resolve;      /* CUDA-event time of the resolve (round) kernels */
    double gpu_ms_analysis;     /* CUDA-event time of radius/dependency kernels */
    double gpu_ms_other;        /* candidate pre-generation, recolour, uploads */
    double host_ms_schedule;    /* host time spent planning the pixel order */
    double wall_ms_total;       /* wall time of the call */
    double gpu_ms_total;        /* CUDA-event time from the first to the last operation of the call on its stream */
} tsb_stats;

enum { TSB_FILTER_TRIANGLE = 0, TSB_FILTER_CATMULLROM = 1, TSB_FILTER_GAUSSIAN = 2 };

/* ImagePyramid::build_gaussian, img_pyramid.rs:20-37.  out: max(levels,1) * w * h * 4 bytes. */
int tsb_pyramid_build(const uint8_t* rgba, uint32_t w, uint32_t h, uint32_t levels, uint8_t* out);

/* image::imageops::resize as the reference calls it (img_pyramid.rs:27-32 Gaussian, ms.rs:244-260
 * Triangle, utils.rs:67-72 CatmullRom). */
int tsb_resize(const uint8_t* rgba, uint32_t w, uint32_t h, uint8_t* out, uint32_t nw, uint32_t nh, int filter);

/* Guide preprocessing for style transfer: utils::transform_to_guide_map (utils.rs:101-116: blur(sigma), grayscale,
 * RGBA (l,l,l,255)) and utils::match_histograms (utils.rs:135-183).  Callers: session.rs:389-393, lib.rs:577-581. */
int tsb_guide_map(const uint8_t* rgba, uint32_t w, uint32_t h, float sigma, uint8_t* out);
int tsb_match_histograms(const uint8_t* source, uint32_t sw, uint32_t sh, const uint8_t* target, uint32_t tw, uint32_t th, uint8_t* out);

int tsb_generator_create(const tsb_generator_desc* desc, tsb_generator** out);
void tsb_generator_destroy(tsb_generator* g);

/* Generator::resolve_random_batch, ms.rs:427-445.  top_levels = pyramid[len-1] of EVERY example
 * (session.rs:42-48). */
int tsb_generator_random_init(tsb_generator* g, uint64_t count, const tsb_image* top_levels, uint32_t n, uint64_t seed);

/* Generator::resolve, ms.rs:702-1052.  Blocking.  guides may be NULL.  sampling has n_examples entries. */
int tsb_generator_resolve(tsb_generator* g, const tsb_params* params, const tsb_pyramid* examples, uint32_t n_examples,
                          const tsb_guides* guides, const tsb_sampling* sampling, tsb_progress_fn cb, void* user);

/* Same computation split in two so that a benchmark can time the path with inputs already resident in
 * HBM: upload once, then resolve (repeatably after tsb_generator_reset). */
int tsb_generator_upload_inputs(tsb_generator* g, const tsb_pyramid* examples, uint32_t n_examples,
                                const tsb_guides* guides, const tsb_sampling* sampling);
int tsb_generator_resolve_resident(tsb_generator* g, const tsb_params* params, tsb_progress_fn cb, void* user);
/* Back to the state right after tsb_generator_create (keeps uploaded inputs). */
int tsb_generator_reset(tsb_generator* g);

/* Read-outs used by GeneratedImage (lib.rs:378-455) and ms.rs:605-684. */
int tsb_generator_read_color(tsb_generator* g, uint8_t* rgba);          /* color_map, w*h*4 */
int tsb_generator_read_coord(tsb_generator* g, uint32_t* xym);          /* coord_map as [x,y,map] triplets (ms.rs:655-684) */
int tsb_generator_read_id(tsb_generator* g, uint32_t* patch_map);       /* id_map as [patch,map] pairs */
int tsb_generator_resolved_count(tsb_generator* g, uint64_t* n, uint64_t* locked);
int tsb_generator_read_resolved(tsb_generator* g, uint32_t* flat, float* score); /* `resolved` in resolution order */
int tsb_generator_read_uncertainty(tsb_generator* g, uint8_t* rgba);    /* get_uncertainty_map, ms.rs:635-653 */
int tsb_generator_read_id_maps(tsb_generator* g, uint8_t* patch_rgba, uint8_t* map_rgba); /* get_id_maps, ms.rs:605-633 */
int tsb_generator_get_stats(tsb_generator* g, tsb_stats* out);

/* ---- band-sharded multi-GPU execution of ONE output (one process per GPU on one node) -------------------------
 * Every rank creates the same generator, uploads the same inputs and calls tsb_generator_resolve* with the same
 * parameters.  Before that the ranks link their replicas: prepare (allocates the shared buffers at their final
 * size) -> export (CUDA IPC handle + offset, 80 bytes per buffer) -> all-gather of the handles by the host application
 * (torch.distributed / MPI / files) -> attach.  `barrier` must block until every rank has called it (e.g.
 * torch.distributed.barrier); it is called a few times per dependency phase.  Phases with fewer than 32768 work
 * items (env TSB_MG_MIN_PHASE) are executed redundantly by every rank; larger ones are sharded by horizontal
 * band: commits are peer stores into every replica, successor notifications are system-scope atomics on the
 * owner's counters.  The result is identical to the single-GPU result and ends up on every rank. */
typedef void (*tsb_barrier_fn)(void* user);
int tsb_generator_mg_prepare(tsb_generator* g, const tsb_params* params, uint32_t* n_handles);
int tsb_generator_mg_export(tsb_generator* g, uint8_t* handles /* n_handles * 80 bytes: IPC handle + offset */);
int tsb_generator_mg_attach(tsb_generator* g, uint32_t rank, uint32_t world, const uint8_t* all_handles /* world * n_handles * 80 */,
                            tsb_barrier_fn barrier, void* user);
int tsb_generator_mg_phases(tsb_generator* g, uint64_t* n /* band-sharded phases executed so far */);

const char* tsb_last_error(void);
int tsb_device_count(void);

/* ---- test-only entry points (parity harness; not used by the Rust shim) ------------------------- */

/* Overwrite the synthesis state with a frozen snapshot: maps, the resolved point set exactly as the
 * reference's tree holds it (including tiling mirror copies), and the resolved list. */
int tsb_generator_load_state(tsb_generator* g, const uint8_t* color, const uint32_t* coord_xym, const uint32_t* id_patch_map,
                             const int32_t* tree_xy, uint64_t n_tree, const uint32_t* resolved_flat,
                             const float* resolved_score, uint64_t n_resolved, uint64_t locked);

/* One pixel resolution per item against the current state WITHOUT committing (ms.rs:917-986).
 * Needs tsb_generator_upload_inputs first.  neigh: n*k*2 int32 (x,y; unused = INT32_MIN);
 * res: n*8 int32 [n_neigh, n_cand, best_idx, best_x, best_y, best_map, best_patch, random]; score: n. */
int tsb_generator_eval_items(tsb_generator* g, const tsb_params* params, int32_t level, float adaptive_alpha,
                             uint64_t p_stage_seed, uint32_t n, const uint32_t* pixel_flat, const uint64_t* loop_seed,
                             int32_t* neigh, int32_t* res, float* score);

/* Per-work-item trace of the last resolve (enable before resolving): pixel, chosen candidate index
 * (-1 = resolve_at_random), number of candidates, number of neighbours, score. */
int tsb_generator_set_trace(tsb_generator* g, int on);
int tsb_generator_trace_count(tsb_generator* g, uint64_t* n);
int tsb_generator_read_trace(tsb_generator* g, uint32_t* pixel, int32_t* best, int32_t* ncand, int32_t* nneigh, float* score);

/* Microbenchmark used for the L2-gather roofline denominator: random 4-byte gathers (one per lane,
 * warp-coherent windows like the scoring kernel's) over a window of `bytes` bytes; returns GB/s of
 * useful bytes and lane-gathers per second.  mode 0 = LDG, 1 = texture object. */
int tsb_microbench_gather(uint64_t bytes, int mode, int iters, double* useful_gbs, double* gathers_per_s);

#ifdef __cplusplus
}
#endif
#endif /* TSB200_H */
