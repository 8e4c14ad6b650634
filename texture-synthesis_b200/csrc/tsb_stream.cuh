// tsb_stream.cuh -- the in-order streaming scheduler of the resolve path (single GPU and band-sharded).
//
// Everything that decides WHICH neighbours a work item sees -- the pixel order (ms.rs:380-389, 905-915), the resolved set
// at every serial time and therefore the k nearest resolved neighbours of every item (ms.rs:926-930) with their distance
// weights (ms.rs:405-425, 1198-1203), and the random candidates (ms.rs:549-599) -- depends on (seed, size, parameters)
// only, never on synthesis results.  The "analysis" kernels below compute all of it ahead of the resolve kernel, on a
// second stream, chunk by chunk into a ring of list buffers:
//   k_lists_chunk<false>  new pixels: neighbours "as of" the item's own serial time (resolved set + lower-index new pixels)
//   k_lists_chunk<true>   redo items: neighbours in the (static) resolved set + one bit per neighbour: re-resolved EARLIER
//                         in this phase (read its new state) or later / never (read the state as of the start of the stage)
//   k_weights             f64 distance chain of ms.rs:405-425 with one THREAD per item (the strictly sequential x4 sum costs
//                         one lane instead of a warp), exp in f64, cast to f32
// The resolve kernel k_stream then needs no search, no dependency graph and no queue: warps claim items in serial order;
// an item reads the 128-bit state of each neighbour with ONE relaxed 128-bit load that carries a phase tag (see st_tag), and
// simply re-reads until the tag says the neighbour has been committed.  Because items are claimed in order, the lowest
// unfinished item never waits on an unclaimed one, so the scheme cannot deadlock for any grid size.  A commit is one
// 128-bit store.  Redo phases write into a second state buffer, so write-after-read hazards do not exist.
#pragma once
#include "tsb_device.cuh"

namespace tsb {

// Single-copy-atomic 128-bit accesses (PTX .b128, LDG/STG.E.128.STRONG.GPU): the payload and its tag travel together,
// so neither side needs a fence -- all a consumer needs from a producer is inside the one word it polls.
__device__ __forceinline__ uint4 ld_state(const uint4* p) {
    uint4 v;
    asm volatile("{\n\t.reg .b128 t;\n\tld.relaxed.gpu.global.b128 t, [%4];\n\tmov.b128 {%0, %1, %2, %3}, t;\n\t}"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state(uint4* p, const uint4 v) {
    asm volatile("{\n\t.reg .b128 t;\n\tmov.b128 t, {%1, %2, %3, %4};\n\tst.relaxed.gpu.global.b128 [%0], t;\n\t}"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_state_sys(const uint4* p) {
    uint4 v;
    asm volatile("{\n\t.reg .b128 t;\n\tld.relaxed.sys.global.b128 t, [%4];\n\tmov.b128 {%0, %1, %2, %3}, t;\n\t}"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state_sys(uint4* p, const uint4 v) {
    asm volatile("{\n\t.reg .b128 t;\n\tmov.b128 t, {%1, %2, %3, %4};\n\tst.relaxed.sys.global.b128 [%0], t;\n\t}"
                 ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Lists of a run of consecutive work items of one phase ("chunk"), produced by the analysis, consumed by k_stream.
struct ChunkDev {
    const uint32_t* pixel;  // [n] flat output pixel of the item
    uint8_t* nbk;           // [n] neighbours in the list (< k only while fewer than k points exist)
    short2* nb;             // [n][k] neighbour offsets n_j - p in canonical order (d^2, dy, dx)
    float* g;               // [n][k] distance weights (f32) of find_best_match, ms.rs:1198-1203
    uint4* low;             // [n] redo phases: bit j = neighbour j is re-resolved EARLIER in this phase
    uint32_t* rand_xy;      // [n][m] random candidates (x | y << 16)
    uint8_t* rand_map;      // [n][m]
    uint32_t n;             // items in the chunk
    uint32_t first;         // stage work-item index of item 0 (= its position in the pick array)
    const uint32_t* tidx;   // band-sharded chunks hold this rank's items only: [n] stage work-item index of each (else nullptr)
};
__device__ __forceinline__ uint32_t chunk_item_index(const ChunkDev& C, uint32_t c) { return C.tidx ? __ldg(C.tidx + c) : C.first + c; }

// pixel -> position in the pick array (= the work-item index of its first resolution in its stage); NONE32 for pixels
// that are never picked (locked before the run)
__global__ void k_tmap_fill(const uint32_t* picks, uint32_t n, uint32_t* tmap) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) tmap[picks[t]] = t;
}

// Neighbour lists of one chunk.  S.mask / S.mask1 = the resolved set at the start of the phase (owned by the analysis
// stream); NEW phases add the stage's own new pixels below the item's index (TimeFilter, tmap holds pick positions).
template <bool REDO>
__global__ void __launch_bounds__(CTA_THREADS) k_lists_chunk(StageDev S, ChunkDev C, TimeFilter T0, const uint32_t* __restrict__ tmap,
                                                             uint32_t first_new, uint32_t n_before) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    KnnScratch* all_ws = reinterpret_cast<KnnScratch*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    KnnScratch& ws = all_ws[warp];
    const uint32_t nwarps = gridDim.x * WARPS_PER_CTA;
    const double area = (double)S.W * (double)S.H;
    for (uint32_t c = blockIdx.x * WARPS_PER_CTA + warp; c < C.n; c += nwarps) {
        const uint32_t i = chunk_item_index(C, c);
        const uint32_t flat = C.pixel[c];
        const int y = (int)(flat / (uint32_t)S.W), x = (int)(flat - (uint32_t)y * (uint32_t)S.W);
        uint32_t r2;
        int kk;
        if (REDO) kk = knn_search<true>(S, ws, lane, x, y, R2_INF, &r2);
        else {
            TimeFilter T = T0;
            T.idx = i - first_new; T.idx_cmp = i;
            const double npts = (double)n_before + (double)T.idx;
            T.hint = (uint32_t)fmin(fmax(1.5 * (double)S.k * area / (3.14159265358979 * fmax(npts, 1.0)), 8.0), 4.0e9);
            T.n_points_max = (uint32_t)fmin((S.tiling ? 3.0 : 1.0) * npts, 4.0e9);
            kk = knn_search<true>(S, ws, lane, x, y, R2_INF, &r2, &T);
        }
        short2* out = C.nb + (size_t)c * S.k;
        uint32_t lw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int j = b * 32 + lane;
            bool low = false;
            if (j < kk) {
                const short2 o = ws.off[j];
                __stcs(reinterpret_cast<uint32_t*>(out + j), (uint32_t)(uint16_t)o.x | ((uint32_t)(uint16_t)o.y << 16));  // written once, read once
                if (REDO) {
                    int qx = x + o.x, qy = y + o.y;
                    if (S.tiling) { qx = imod(qx, S.W); qy = imod(qy, S.H); }
                    low = __ldg(tmap + (size_t)qy * S.W + qx) < i;  // NONE32 (locked) is never below an index
                }
            }
            if (REDO) lw[b] = __ballot_sync(FULL, low);
        }
        if (lane == 0) {
            C.nbk[c] = (uint8_t)kk;
            if (REDO) C.low[c] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
        __syncwarp();
    }
}

// get_distances_to_k_neighs + the gaussians of find_best_match (ms.rs:405-425, 1198-1203) for every item of a chunk:
// d_j = fma(dx, dx, dy * dy) on the normalised coordinates (tables divx / divy hold the reference's divisions), mean over the
// x4-duplicated list as a strictly sequential f64 sum, g_j = (f32) exp(-(d_j / mean)).  Each warp takes 32 items at a time:
//   1. item by item, lanes = neighbours: the distances (coalesced list reads, table reads that share sectors) -> shared memory
//   2. lanes = items: the sequential sum -- 32 chains side by side, one lane each instead of one warp each
//   3. (item, neighbour) pairs flattened over the lanes: division, exp, cast, coalesced store
// items per warp and pass: 16 keeps the staging area at 6.6 KB per warp (k = 50); measured per 2048^2 step: 8 -> 4.29 ms,
// 16 -> 4.24 ms, 32 -> 6.8 ms
constexpr int KW_WARPS = 4, KW_ITEMS = 16;
__global__ void __launch_bounds__(KW_WARPS * 32) k_weights(StageDev S, ChunkDev C) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int k = S.k;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int P = KW_ITEMS + 1;  // padded pitch: (j, item) at j * P + item
    double* sd = reinterpret_cast<double*>(smem_raw) + (size_t)warp * (size_t)k * P;
    const uint32_t ngroups = (C.n + KW_ITEMS - 1) / KW_ITEMS;
    for (uint32_t grp = blockIdx.x * KW_WARPS + warp; grp < ngroups; grp += gridDim.x * KW_WARPS) {
        const uint32_t c0 = grp * KW_ITEMS;
        const int rows = (int)min((uint32_t)KW_ITEMS, C.n - c0);
        int my_kk = 0, my_x = 0, my_y = 0;
        double my_x2 = 0.0, my_y2 = 0.0;
        if (lane < rows) {  // lanes = items: the per-item set-up once, side by side
            my_kk = (int)C.nbk[c0 + lane];
            const uint32_t flat = C.pixel[c0 + lane];
            my_y = (int)(flat / (uint32_t)S.W); my_x = (int)(flat - (uint32_t)my_y * (uint32_t)S.W);
            my_x2 = __ldg(S.divx + my_x + S.mx); my_y2 = __ldg(S.divy + my_y + S.my);
        }
        for (int it = 0; it < rows; ++it) {
            const uint32_t c = c0 + (uint32_t)it;
            const int kk = __shfl_sync(FULL, my_kk, it);
            const int x = __shfl_sync(FULL, my_x, it), y = __shfl_sync(FULL, my_y, it);
            const double x2 = __shfl_sync(FULL, my_x2, it), y2 = __shfl_sync(FULL, my_y2, it);
            for (int j = lane; j < kk; j += 32) {
                const short2 o = C.nb[(size_t)c * k + j];
                const double ddx = __dsub_rn(__ldg(S.divx + x + o.x + S.mx), x2), ddy = __dsub_rn(__ldg(S.divy + y + o.y + S.my), y2);
                sd[(size_t)j * P + it] = __fma_rn(ddx, ddx, __dmul_rn(ddy, ddy));
            }
        }
        __syncwarp();
        double sum = 0.0;
        for (int j = 0; j < my_kk; ++j) {  // lanes >= rows have my_kk = 0
            const double d = sd[(size_t)j * P + lane];
            sum = __dadd_rn(sum, d); sum = __dadd_rn(sum, d); sum = __dadd_rn(sum, d); sum = __dadd_rn(sum, d);
        }
        const double my_avg = __ddiv_rn(sum, (double)(my_kk * 4));  // 0 / 0 = NaN for an empty list (never read)
        const int total = rows * k;
        int it = lane / k, j = lane - it * k;  // (item, neighbour) of element e, advanced without a division per trip
        for (int e = lane; e < ((total + 31) & ~31); e += 32, j += 32) {
            while (j >= k) { j -= k; ++it; }
            const double avg = __shfl_sync(FULL, my_avg, it & 31);
            const int kk = __shfl_sync(FULL, my_kk, it & 31);
            if (e < total) {
                float gv = 0.f;
                // avg == 0 (the pixel is its own only neighbour): 0 / 0 -> NaN weights, as in the reference
                if (j < kk) gv = (float)exp(-__ddiv_rn(sd[(size_t)j * P + it], avg));
                __stcs(C.g + (size_t)c0 * k + e, gv);
            }
        }
        __syncwarp();
    }
}

// Control block of one k_stream launch
enum { SC_NEXT = 0, SC_ABORT = 32, SC_WORDS = 64 };

struct StreamDev {
    const uint4* prev;        // redo phases: the state as of the start of the stage
    uint4* cur;               // the state this phase writes (and, for new-pixel phases, reads)
    uint32_t* ctl;            // SC_*
    uint32_t* abort_flag;     // sticky per-run abort flag (watchdog)
    volatile uint32_t* progress;  // mapped host word (or nullptr): work items claimed so far in this run
    uint32_t* live_color;     // progress callback registered: plain colour plane kept up to date for copy-engine snapshots
    uint32_t progress_base;
    uint32_t tag;             // phase id written with every commit
    uint32_t watchdog_ms;     // a single wait longer than this aborts the run
    uint32_t profile;         // 1: also accumulate the per-section cycle counters (TSB_DEBUG_PHASES)
    // per-item trace (tests)
    int32_t* tr_best; int32_t* tr_ncand; int32_t* tr_nneigh; float* tr_score;
    uint64_t trace_base;
    // band-sharded execution (MG): the state of a pixel lives with the rank that owns its row; neighbours outside this rank's
    // band are read (and polled) in the owner's replica over NVLink -- nothing is ever pushed
    int world, rank, band_h, y0, y1;      // this rank owns rows [y0, y1)
    const uint4* prev_r[MG_MAX];
    const uint4* cur_r[MG_MAX];
};

// Persistent in-order resolve kernel for one chunk.  REDO: re-resolution of already resolved pixels (ms.rs:905-907) --
// neighbours flagged "earlier" are read from `cur` once their tag is this phase's, all others from `prev`; the result goes
// to `cur`.  Otherwise new pixels (ms.rs:909-915): every neighbour is read from `cur` once its tag is non-zero.
struct __align__(16) StreamSmem {
    CtaSmem c;
    unsigned long long stat[ST_COUNT];
};
#ifndef TSB_STREAM_CTAS
#define TSB_STREAM_CTAS 3   /* 4 (64 registers) was tried: every item takes longer, the L1 gather path is the limit */
#endif
template <bool GUIDED, bool OPAQUE, bool REDO, bool MG = false>
__global__ void __launch_bounds__(CTA_THREADS, TSB_STREAM_CTAS) k_stream(StageDev S, ChunkDev C, StreamDev D) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    StreamSmem& rs = *reinterpret_cast<StreamSmem*>(smem_raw);
    CtaSmem& sm = rs.c;
    if (threadIdx.x < ST_COUNT) rs.stat[threadIdx.x] = 0ull;
    load_luts(S, sm.lut, sm.lutg);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    WarpScratch& ws = sm.ws[warp];
    if (lane < 16) ws.stat[lane] = 0ull;
    __syncwarp();
    const int k = S.k, W = S.W, H = S.H;
    // Items are claimed in order, one per atomic, and nothing is claimed ahead.  Measured and dropped (2048^2 step): claiming one
    // item ahead and pulling its lists into L2 while the current one is resolved -- safe (the lowest unfinished item is always
    // somebody's CURRENT item) but a held item cannot start elsewhere, which costs more in waiting than the prefetch saves
    // (50.7 -> 53.3 ms); claiming 2 / 4 consecutive items per atomic in the throughput phases -- 46.4 -> 49.3 / 64.6 ms, the held
    // items are exactly the ones other warps end up waiting for; an L2 prefetch of the lists of the item one generation of
    // resident warps ahead (no claim involved) -- 46.35 ms, the list latency is already hidden.
    uint32_t c = 0;
    if (lane == 0) c = atomicAdd(D.ctl + SC_NEXT, 1u);
    c = __shfl_sync(FULL, c, 0);
    while (c < C.n) {
        long long tr0 = clock64();
        uint32_t c_next = 0;
        if (lane == 0 && D.progress && (c & 1023u) == 0u) *D.progress = D.progress_base + c;
        // ---- the item's lists (prepared by the analysis) ----
        const int kk = (int)C.nbk[c];
        const uint32_t flat = C.pixel[c];
        const int y = (int)(flat / (uint32_t)W), x = (int)(flat - (uint32_t)y * (uint32_t)W);
        const uint32_t si = chunk_item_index(C, c);
        const uint32_t* rand_xy = C.rand_xy + (size_t)c * S.m;
        const uint8_t* rand_map = C.rand_map + (size_t)c * S.m;
        uint4 lowm = make_uint4(0u, 0u, 0u, 0u);
        if (REDO) lowm = C.low[c];
        // the lists are read exactly once: streaming loads (evict-first) keep the state and the example level in L2
        // (all k slots, so that these loads do not wait for the neighbour count: one round trip to memory instead of two)
        for (int j0 = 0; j0 < k; j0 += 64) {
            const int ja = j0 + lane, jb = ja + 32;
            uint32_t ova = 0u, ovb = 0u;
            float ga = 0.f, gb = 0.f;
            if (ja < k) { ova = __ldcs(reinterpret_cast<const uint32_t*>(C.nb + (size_t)c * k + ja)); ga = __ldcs(C.g + (size_t)c * k + ja); }
            if (jb < k) { ovb = __ldcs(reinterpret_cast<const uint32_t*>(C.nb + (size_t)c * k + jb)); gb = __ldcs(C.g + (size_t)c * k + jb); }
            if (ja < k) { ws.off[ja] = make_short2((short)(ova & 0xFFFFu), (short)(ova >> 16)); ws.g[ja] = ga; }
            if (jb < k) { ws.off[jb] = make_short2((short)(ovb & 0xFFFFu), (short)(ovb >> 16)); ws.g[jb] = gb; }
        }
        uint32_t rxy0 = 0, rxy1 = 0, rmp0 = 0, rmp1 = 0;
        if (lane < S.m) { rxy0 = __ldcs(rand_xy + lane); rmp0 = __ldcs(rand_map + lane); }
        if (lane + 32 < S.m) { rxy1 = __ldcs(rand_xy + lane + 32); rmp1 = __ldcs(rand_map + lane + 32); }
        __syncwarp();
        long long t1 = clock64();
        ItemOut o;
        o.c_knn = t1 - tr0; o.c_neigh = o.c_weight = o.c_score = 0; o.fetched = o.nominal = 0;
        o.kk = kk; o.ncand = 0; o.best = 0; o.bx = o.by = o.bmap = 0; o.bpatch = 0; o.score = 0.f; o.bcol = 0; o.bcol_valid = 0;
        long long waited = 0;
        bool aborted = false;
        if (kk > 0) {
            // ---- neighbour state (target pattern ms.rs:1151-1181, coherence candidates ms.rs:496-547), two per lane in flight ----
            int ncand = 0, reach = 0;
            for (int base = 0; base < kk; base += 64) {
                uint4 st[2];
                const uint4* src[2];
                short2 off[2];
                bool need[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int j = base + 32 * q + lane;
                    need[q] = false; src[q] = nullptr; off[q] = make_short2(0, 0); st[q] = make_uint4(0u, 0u, 0u, 0u);
                    if (j < kk) {
                        off[q] = ws.off[j];
                        int qx = x + off[q].x, qy = y + off[q].y;
                        if (S.tiling) { qx = imod(qx, W); qy = imod(qy, H); }
                        bool low = false;
                        if (REDO) {
                            const uint32_t wsel = (j >> 5) == 0 ? lowm.x : (j >> 5) == 1 ? lowm.y : (j >> 5) == 2 ? lowm.z : lowm.w;
                            low = (wsel >> (j & 31)) & 1u;
                        }
                        if (MG && (qy < D.y0 || qy >= D.y1)) {  // the owner's replica holds the authoritative copy
                            int r = qy / D.band_h;
                            r = r < D.world - 1 ? r : D.world - 1;
                            src[q] = ((REDO && !low) ? D.prev_r[r] : D.cur_r[r]) + ((size_t)qy * W + qx);
                        } else src[q] = ((REDO && !low) ? D.prev : D.cur) + ((size_t)qy * W + qx);
                        st[q] = MG ? ld_state_sys(src[q]) : ld_state(src[q]);
                        need[q] = REDO ? (low && st_tag(st[q].w) != D.tag) : (st_tag(st[q].w) == 0u);
                    }
                }
                if (__any_sync(FULL, need[0] || need[1])) {
                    // a neighbour has not been committed yet (its item is in flight on another warp): poll it
                    const long long tw0 = clock64();
                    unsigned long long t_begin = 0;
                    unsigned ns = 32, polls = 0;
                    do {
                        __nanosleep(ns);
                        if (ns < 512) ns <<= 1;
#pragma unroll
                        for (int q = 0; q < 2; ++q)
                            if (need[q]) {
                                st[q] = MG ? ld_state_sys(src[q]) : ld_state(src[q]);
                                need[q] = REDO ? (st_tag(st[q].w) != D.tag) : (st_tag(st[q].w) == 0u);
                            }
                        if ((++polls & 255u) == 0u) {  // watchdog in wall-clock time: a stalled run is an error, never a hang
                            int stop = 0;
                            if (lane == 0) {
                                const unsigned long long now = globaltimer_ns();
                                if (t_begin == 0) t_begin = now;
                                if (*((volatile uint32_t*)D.abort_flag)) stop = 1;
                                else if (now - t_begin > (unsigned long long)D.watchdog_ms * 1000000ull) { atomicExch(D.abort_flag, 1u); stop = 1; }
                            }
                            if (__shfl_sync(FULL, stop, 0)) { aborted = true; break; }
                        }
                    } while (__any_sync(FULL, need[0] || need[1]));
                    waited += clock64() - tw0;
                    if (aborted) break;
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int j = base + 32 * q + lane;
                    bool valid = false;
                    uint32_t cxy = 0, cpatch = 0;
                    uint16_t cmeta = 0;
                    if (j < kk) {
                        const short2 of = off[q];
                        reach = max(reach, max(abs((int)of.x), abs((int)of.y)));
                        ws.tcol[j] = st[q].x;
                        if (GUIDED) {
                            int gx = x + of.x, gy = y + of.y;
                            if (S.tiling) { gx = imod(gx, S.tgw); gy = imod(gy, S.tgh); }
                            ws.gcol[j] = ((unsigned)gx < (unsigned)S.tgw && (unsigned)gy < (unsigned)S.tgh)
                                             ? __ldg(S.tguide + (size_t)gy * S.tgw + gx) : OUTSIDE_RGBA;
                        }
                        const int sx = (int)(st[q].y & 0xFFFFu), sy = (int)(st[q].y >> 16);
                        const uint32_t map = st_idmap(st[q].w);  // id_map's MapId (ms.rs:510-511)
                        const int cx = sx - of.x, cy = sy - of.y;  // source of the neighbour + (p - n)
                        if (map < (uint32_t)S.n_ex) {
                            DevEx e = S.ex[map];
                            if ((unsigned)cx < (unsigned)e.w && (unsigned)cy < (unsigned)e.h)
                                valid = e.smask ? (__ldg(e.smask + (size_t)cy * e.w + cx) != 0) : true;
                        }
                        cxy = (uint32_t)cx | ((uint32_t)cy << 16);
                        cpatch = st[q].z;
                        cmeta = (uint16_t)map;
                    }
                    const unsigned b = __ballot_sync(FULL, valid);
                    if (valid) {
                        const int pos = ncand + __popc(b & lt);
                        ws.u.c.cxy[pos] = cxy; ws.u.c.cpatch[pos] = cpatch; ws.cmeta[pos] = cmeta;
                    }
                    ncand += __popc(b);
                }
            }
            if (aborted) break;
            __syncwarp();
            long long t2 = clock64();
            const float g0 = ws.g[0];
            const bool degenerate = !(g0 == g0);  // NaN weights: the pixel is its own only neighbour (see k_weights)
            resolve_tail<GUIDED, OPAQUE ? 1 : 0>(S, ws, sm.lut, sm.lutg, lane, kk, ncand, reach, degenerate, rand_xy, rand_map,
                                                 rxy0, rxy1, rmp0, rmp1, t1, t2, o);
            o.c_neigh -= waited;
        }
        long long tc0 = clock64();
        if (lane == 0) {
            if (kk > 0) {
                DevEx e = S.ex[o.bmap];
                const uint32_t col = o.bcol_valid ? o.bcol : __ldg(e.px + (size_t)o.by * e.w + o.bx);
                if (!REDO) S.score[flat] = o.score;  // first resolution only (ms.rs:365)
                const uint4 v = make_uint4(col, (uint32_t)o.bx | ((uint32_t)o.by << 16), o.bpatch, st_pack_w((uint32_t)o.bmap, (uint32_t)o.bmap, D.tag));
                if (MG) st_state_sys(D.cur + flat, v); else st_state(D.cur + flat, v);
                if (D.live_color) D.live_color[flat] = col;
            }
            if (D.tr_best) {
                const size_t ti = (size_t)(D.trace_base + si);
                D.tr_best[ti] = kk > 0 ? o.best : -1;
                D.tr_ncand[ti] = o.ncand; D.tr_nneigh[ti] = kk; D.tr_score[ti] = o.score;
            }
            ws.stat[ST_FETCHED] += o.fetched; ws.stat[ST_NOMINAL] += o.nominal; ws.stat[ST_CANDS] += (unsigned long long)o.ncand;
            ws.stat[ST_ITEMS] += 1ull;
            if (D.profile) {
                ws.stat[ST_CYC_READY] += (unsigned long long)waited;
                ws.stat[ST_CYC_KNN] += (unsigned long long)o.c_knn; ws.stat[ST_CYC_NEIGH] += (unsigned long long)o.c_neigh;
                ws.stat[ST_CYC_WEIGHT] += (unsigned long long)o.c_weight; ws.stat[ST_CYC_SCORE] += (unsigned long long)o.c_score;
                ws.stat[ST_CYC_COMMIT] += (unsigned long long)(clock64() - tc0);
            }
        }
        __syncwarp();
        if (lane == 0) c_next = atomicAdd(D.ctl + SC_NEXT, 1u);
        c = __shfl_sync(FULL, c_next, 0);
    }
    if (lane == 0 && S.counters) {
        for (int i = 0; i < ST_COUNT; ++i) if (ws.stat[i]) atomicAdd(&rs.stat[i], ws.stat[i]);
    }
    __syncthreads();
    if (S.counters && threadIdx.x < ST_COUNT && rs.stat[threadIdx.x]) atomicAdd(S.counters + threadIdx.x, rs.stat[threadIdx.x]);
}

// Device-side barrier between the ranks of a band-sharded run (one per phase): every rank stores its sequence number into
// every peer's flag block (peer-mapped memory, release at system scope) and then waits, polling its OWN block, until all
// peers have stored theirs.  Stream-ordered: no host synchronisation.
__global__ void k_mg_signal(uint32_t* const* flags_of_rank, int world, int rank, uint32_t seq) {
    const int r = threadIdx.x;
    if (r < world) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags_of_rank[r] + rank), "r"(seq) : "memory");
}
__global__ void k_mg_wait(const uint32_t* my_flags, int world, uint32_t seq, uint32_t watchdog_ms, uint32_t* abort_flag) {
    const int r = threadIdx.x;
    if (r >= world) return;
    unsigned long long t0 = 0;
    unsigned polls = 0;
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flags + r) : "memory");
        if ((int32_t)(v - seq) >= 0) break;
        __nanosleep(200);
        if ((++polls & 1023u) == 0u) {
            const unsigned long long now = globaltimer_ns();
            if (t0 == 0) t0 = now;
            if (*((volatile uint32_t*)abort_flag)) break;
            if (now - t0 > (unsigned long long)watchdog_ms * 1000000ull) { atomicExch(abort_flag, 1u); break; }
        }
    }
}
// ownership flags of the picks for the ordered compaction (band-sharded runs)
__global__ void k_own_flags(const uint32_t* picks, uint32_t n, int W, int band_h, int world, int rank, uint8_t* flag) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int r = (int)(picks[t] / (uint32_t)W) / band_h;
    r = r < world - 1 ? r : world - 1;
    flag[t] = r == rank ? 1 : 0;
}
__global__ void k_gather_u32(const uint32_t* src, const uint32_t* idx, uint32_t n, uint32_t* dst) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) dst[t] = src[idx[t]];
}

}  // namespace tsb
