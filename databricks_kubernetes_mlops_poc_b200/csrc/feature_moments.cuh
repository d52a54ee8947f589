/*
 * feature_moments.cuh -- K2: per-feature (count, mean, M2) over N encoded rows, sm_100a.
 *
 * BASELINE config 5 ("drift-monitor path: per-feature mean/var reduction").  The reference has no
 * mean/var computation; the nearest call on its path is `self.drift.predict(df[all].values)`
 * (reference databricks/src/02-register-model.ipynb:338).  This is a pure HBM-bound streaming
 * reduction: 96 B read per row, 24*3 float64 written per launch.
 *
 * Mapping: a row is six 16-byte vectors; thread t reads vector t%6 of row t/6 of its slab, so a warp
 * reads 512 contiguous bytes per load (LDG.128, fully coalesced).  Each thread keeps shifted sums
 * sum(x-K), sum((x-K)^2) and a count for its four words in float64 (K = the word's value in row 0,
 * read by every thread; the shift removes the catastrophic cancellation of the raw sum-of-squares
 * form).  Block partials are reduced through shared memory in a fixed order, written to global
 * memory, and the last block to finish (atomic ticket) reduces the partials in a fixed order, so the
 * result is deterministic.  NaN values are skipped (count is per word).
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/b2f.h"

#define B2F_MOM_ROWS_PER_BLOCK 64
#define B2F_MOM_THREADS (B2F_MOM_ROWS_PER_BLOCK * 6)
#define B2F_MOM_VALUES (B2F_ROW_WORDS * 3)

__device__ __forceinline__ double mom_word_value(uint32_t w, int word, int n_cat) {
    return word < n_cat ? (double)(int32_t)w : (double)__uint_as_float(w);
}

__global__ void __launch_bounds__(B2F_MOM_THREADS)
    k_feature_moments(const uint4 *__restrict__ rows, long long n, int n_cat, double *__restrict__ partials,
                      unsigned int *__restrict__ ticket, double *__restrict__ out) {
    __shared__ double red[B2F_MOM_THREADS][12 + 1];
    __shared__ double tot[B2F_MOM_VALUES];
    __shared__ bool is_last;

    const int q = threadIdx.x % 6;  /* which 16-byte vector of the row */
    const int rr = threadIdx.x / 6; /* row within the slab */

    /* pivot: row 0's values (NaN -> 0) */
    double K[4];
    {
        const uint4 v0 = n > 0 ? __ldg(rows + q) : make_uint4(0, 0, 0, 0);
        const uint32_t w0[4] = {v0.x, v0.y, v0.z, v0.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double x = mom_word_value(w0[k], q * 4 + k, n_cat);
            K[k] = (x == x) ? x : 0.0;
        }
    }

    double cnt[4] = {0, 0, 0, 0}, s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
    const long long stride = (long long)gridDim.x * B2F_MOM_ROWS_PER_BLOCK;
#pragma unroll 4
    for (long long row = (long long)blockIdx.x * B2F_MOM_ROWS_PER_BLOCK + rr; row < n; row += stride) {
        const uint4 v = __ldg(rows + row * 6 + q);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double x = mom_word_value(w[k], q * 4 + k, n_cat);
            if (x == x) {
                const double d = x - K[k];
                cnt[k] += 1.0;
                s[k] += d;
                ss[k] = fma(d, d, ss[k]);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        red[threadIdx.x][k * 3 + 0] = cnt[k];
        red[threadIdx.x][k * 3 + 1] = s[k];
        red[threadIdx.x][k * 3 + 2] = ss[k];
    }
    __syncthreads();

    /* 72 threads: (word, component) -> fixed-order sum over the 64 row lanes */
    if (threadIdx.x < B2F_MOM_VALUES) {
        const int word = threadIdx.x / 3, comp = threadIdx.x % 3;
        const int wq = word / 4, wk = word % 4;
        double a = 0.0;
        for (int r = 0; r < B2F_MOM_ROWS_PER_BLOCK; ++r) a += red[r * 6 + wq][wk * 3 + comp];
        partials[(size_t)blockIdx.x * B2F_MOM_VALUES + threadIdx.x] = a;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();

    /* last block: reduce the block partials in a fixed order -- 5 interleaved segments per value so
     * 360 threads work and every thread's loads are independent (the adds form 5 short chains) */
    {
        double *seg = &red[0][0]; /* reuse: [5][72] */
        __syncthreads();
        if (threadIdx.x < 5 * B2F_MOM_VALUES) {
            const int v = threadIdx.x % B2F_MOM_VALUES, sgm = threadIdx.x / B2F_MOM_VALUES;
            double a = 0.0;
#pragma unroll 8
            for (unsigned int b = sgm; b < gridDim.x; b += 5) a += __ldcg(partials + (size_t)b * B2F_MOM_VALUES + v);
            seg[sgm * B2F_MOM_VALUES + v] = a;
        }
        __syncthreads();
        if (threadIdx.x < B2F_MOM_VALUES) {
            double a = 0.0;
            for (int sgm = 0; sgm < 5; ++sgm) a += seg[sgm * B2F_MOM_VALUES + threadIdx.x];
            tot[threadIdx.x] = a;
        }
    }
    __syncthreads();
    if (threadIdx.x < B2F_ROW_WORDS) {
        const int word = threadIdx.x;
        const double c = tot[word * 3 + 0], S = tot[word * 3 + 1], SS = tot[word * 3 + 2];
        double Kw = 0.0; /* this word's pivot: its value in row 0 (NaN -> 0), as above */
        if (n > 0) {
            const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t *>(rows) + word);
            Kw = mom_word_value(w0, word, n_cat);
            if (!(Kw == Kw)) Kw = 0.0;
        }
        double mean = 0.0, m2 = 0.0;
        if (c > 0.0) {
            mean = Kw + S / c;
            m2 = SS - S * S / c;
            if (m2 < 0.0) m2 = 0.0;
        }
        out[word * 3 + 0] = c;
        out[word * 3 + 1] = mean;
        out[word * 3 + 2] = m2;
    }
    if (threadIdx.x == 0) *ticket = 0; /* re-arm for the next launch on this stream */
}
