// tsb_rng.cuh -- the reference's RNG stack, bit-exact, usable from host and device.
//
// rand_pcg 0.3.1 `Lcg64Xsh32` (= Pcg32), rand_core 0.6.3 `SeedableRng::seed_from_u64` (default impl),
// rand 0.8.5 `Rng::gen_range` -> `UniformInt::sample_single_inclusive`.
// Reference call sites: lib/src/ms.rs:386 (pick_random_unresolved), 454-458 (resolve_at_random),
// 549-564 (random candidates), 616-627 (debug colours), 803-804 (stage seed).
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define TSB_HD __host__ __device__ __forceinline__
#else
#define TSB_HD inline
#endif

namespace tsb {

struct Pcg32 {
    uint64_t state, inc;

    static constexpr uint64_t MUL = 6364136223846793005ULL;

    TSB_HD void step() { state = state * MUL + inc; }

    // Lcg64Xsh32::from_state_incr
    TSB_HD static Pcg32 from_state_incr(uint64_t s, uint64_t i) {
        Pcg32 r;
        r.state = s + i;
        r.inc = i;
        r.step();
        return r;
    }

    TSB_HD static uint32_t rotr(uint32_t x, uint32_t r) { return (x >> (r & 31)) | (x << ((32 - r) & 31)); }

    // SeedableRng::seed_from_u64: four PCG32 outputs (increment 11634580027462260723) fill the 16-byte seed,
    // then Lcg64Xsh32::from_seed reads two little-endian u64 and forces the increment odd.
    TSB_HD static Pcg32 seed_from_u64(uint64_t st) {
        uint32_t w[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            st = st * MUL + 11634580027462260723ULL;
            uint32_t xs = (uint32_t)(((st >> 18) ^ st) >> 27);
            w[c] = rotr(xs, (uint32_t)(st >> 59));
        }
        uint64_t a = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
        uint64_t b = (uint64_t)w[2] | ((uint64_t)w[3] << 32);
        return from_state_incr(a, b | 1);
    }

    TSB_HD uint32_t next_u32() {
        uint64_t s = state;
        step();
        return rotr((uint32_t)(((s >> 18) ^ s) >> 27), (uint32_t)(s >> 59));
    }

    // next_u64_via_u32: low word first
    TSB_HD uint64_t next_u64() {
        uint64_t lo = next_u32();
        uint64_t hi = next_u32();
        return (hi << 32) | lo;
    }

    TSB_HD static uint32_t clz32(uint32_t v) {
#ifdef __CUDA_ARCH__
        return (uint32_t)__clz((int)v);
#else
        return (uint32_t)__builtin_clz(v);
#endif
    }
    TSB_HD static uint32_t clz64(uint64_t v) {
#ifdef __CUDA_ARCH__
        return (uint32_t)__clzll((long long)v);
#else
        return (uint32_t)__builtin_clzll(v);
#endif
    }
    TSB_HD static uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
        return __umul64hi(a, b);
#else
        return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
    }

    // gen_range(0..n) for u32 (example width/height, ms.rs:456,458,563,564)
    TSB_HD uint32_t gen_range_u32(uint32_t n) {
        uint32_t zone = (n << clz32(n)) - 1u;
        for (;;) {
            uint64_t m = (uint64_t)next_u32() * (uint64_t)n;
            if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
        }
    }
    // gen_range(0..n) for usize on a 64-bit target (ms.rs:386,454,552): 64-bit draws
    TSB_HD uint64_t gen_range_usize(uint64_t n) {
        uint64_t zone = (n << clz64(n)) - 1ull;
        for (;;) {
            uint64_t v = next_u64();
            uint64_t lo = v * n;
            if (lo <= zone) return mulhi64(v, n);
        }
    }
    // gen_range(0..n) for u8 (ms.rs:616-627): u32 draws, exact modulus zone
    TSB_HD uint8_t gen_range_u8(uint8_t n) {
        uint32_t range = n;
        uint32_t zone = 0xFFFFFFFFu - ((0xFFFFFFFFu - range + 1u) % range);
        for (;;) {
            uint64_t m = (uint64_t)next_u32() * (uint64_t)range;
            if ((uint32_t)m <= zone) return (uint8_t)(m >> 32);
        }
    }
};

}  // namespace tsb
