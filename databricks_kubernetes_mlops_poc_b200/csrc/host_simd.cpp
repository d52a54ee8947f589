/*
 * host_simd.cpp -- host-side SIMD for the ranked-row encoder (compiled by g++, not nvcc; linked into libb200forest.so).
 *
 * rank_k(x) = #{ j : t'_j <= x } over the sorted distinct split values of numeric feature k (csrc/forest_rank.h) is the
 * one piece of arithmetic the ranked row format moves from the GPU kernel to the request encoder (reference
 * counterpart: the float32-vs-threshold compares inside sklearn's tree walk, databricks/src/02-register-model.ipynb:335-337).
 * A request batch needs n_rows x 14 of them, so it is done without branches and without per-element gathers:
 *
 *   the split values of a feature are stored as a wide search tree laid out level by level ("rank table"):
 *     fan[j]   = fan-out of level j: 16 for the upper levels, 16 / 32 / 48 / 64 for the last one
 *     full     = the sorted values padded with +inf to prod(fan) entries (> m, so every level's count stays below its fan-out)
 *     lvl[j][i] = full[(i + 1) * S_j - 1], S_j = prod(fan[j+1..])      the LARGEST value of the i-th chunk at level j
 *   one level = fan/16 64-byte loads of consecutive chunk maxima, vector compares against the broadcast x, popcounts:
 *   c_j = #{chunk maxima <= x} is the digit of the rank in the mixed radix, and selects the chunk to descend into.
 *   m <= 63: one level; m <= 1023: two (16 x up to 64: 4 KB per feature, so 14 features stay in L1); m <= 16383: three.
 *
 * AVX-512F path when the CPU has it (the B200 hosts do), else an AVX2 form of the same walk (2 x 8 lanes), else scalar.
 */
#include <immintrin.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

extern "C" {

struct b2f_ranktab {
    int32_t levels;      /* 1..4 */
    int32_t count;       /* m: real split values */
    int32_t fan[4];      /* fan-out per level: multiples of 16 */
    const float *lvl[4]; /* lvl[j]: prod(fan[0..j]) floats */
};

int b2f_simd_level(void) {
    static int level = -1;
    if (level < 0) {
        __builtin_cpu_init();
        level = __builtin_cpu_supports("avx512f") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
        const char *cap = getenv("B2F_SIMD"); /* test hook: "0" scalar, "1" at most AVX2 */
        if (cap && cap[0] >= '0' && cap[0] <= '2' && (cap[0] - '0') < level) level = cap[0] - '0';
    }
    return level;
}

static inline uint32_t rank_scalar(const b2f_ranktab *t, float x) {
    uint32_t idx = 0;
    for (int j = 0; j < t->levels; ++j) {
        const int fan = t->fan[j];
        const float *p = t->lvl[j] + (size_t)idx * fan;
        uint32_t c = 0;
        for (int i = 0; i < fan; ++i) c += p[i] <= x ? 1u : 0u;
        idx = idx * fan + c;
    }
    return idx;
}

__attribute__((target("avx2,popcnt"))) static inline uint32_t rank_avx2(const b2f_ranktab *t, float x) {
    const __m256 vx = _mm256_set1_ps(x);
    uint32_t idx = 0;
    for (int j = 0; j < t->levels; ++j) {
        const int fan = t->fan[j];
        const float *p = t->lvl[j] + (size_t)idx * fan;
        uint32_t c = 0;
        for (int i = 0; i < fan; i += 8) c += (uint32_t)__builtin_popcount((unsigned)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_loadu_ps(p + i), vx, _CMP_LE_OQ)));
        idx = idx * fan + c;
    }
    return idx;
}

__attribute__((target("avx512f,popcnt"))) static inline uint32_t rank_avx512(const b2f_ranktab *t, float x) {
    const __m512 vx = _mm512_set1_ps(x);
    uint32_t idx = 0;
    for (int j = 0; j < t->levels; ++j) {
        const int fan = t->fan[j];
        const float *p = t->lvl[j] + (size_t)idx * fan;
        uint32_t c = (uint32_t)__builtin_popcount((unsigned)_mm512_cmp_ps_mask(_mm512_loadu_ps(p), vx, _CMP_LE_OQ));
        for (int i = 16; i < fan; i += 16) c += (uint32_t)__builtin_popcount((unsigned)_mm512_cmp_ps_mask(_mm512_loadu_ps(p + i), vx, _CMP_LE_OQ));
        idx = idx * fan + c;
    }
    return idx;
}

#define RANK_COLUMN_BODY(RANKFN)                                                              \
    for (int64_t i = 0; i < n; ++i) {                                                         \
        const float v = x[i * x_stride];                                                      \
        out[i * out_stride] = (v != v) ? nan_rank : (uint16_t)RANKFN(t, v);                   \
    }

__attribute__((target("avx512f,popcnt"))) static void rank_column_avx512(const b2f_ranktab *t, const float *x, int64_t n, int64_t x_stride, uint16_t *out,
                                                                           int64_t out_stride, uint16_t nan_rank) {
    RANK_COLUMN_BODY(rank_avx512)
}
__attribute__((target("avx2,popcnt"))) static void rank_column_avx2(const b2f_ranktab *t, const float *x, int64_t n, int64_t x_stride, uint16_t *out,
                                                                      int64_t out_stride, uint16_t nan_rank) {
    RANK_COLUMN_BODY(rank_avx2)
}
static void rank_column_scalar(const b2f_ranktab *t, const float *x, int64_t n, int64_t x_stride, uint16_t *out, int64_t out_stride, uint16_t nan_rank) {
    RANK_COLUMN_BODY(rank_scalar)
}

/* ranks of n values of ONE feature: x[i * x_stride] -> out[i * out_stride] (strides in elements); NaN -> nan_rank */
void b2f_simd_rank_column(const b2f_ranktab *t, const float *x, int64_t n, int64_t x_stride, uint16_t *out, int64_t out_stride, uint16_t nan_rank) {
    switch (b2f_simd_level()) {
        case 2: rank_column_avx512(t, x, n, x_stride, out, out_stride, nan_rank); break;
        case 1: rank_column_avx2(t, x, n, x_stride, out, out_stride, nan_rank); break;
        default: rank_column_scalar(t, x, n, x_stride, out, out_stride, nan_rank); break;
    }
}

} /* extern "C" */
