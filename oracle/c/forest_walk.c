/*
 * oracle/c/forest_walk.c -- CPU restatement of the scoring hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates what the reference's `classifier.predict_proba(df[all_features])[:, 1]`
 * (reference databricks/src/02-register-model.ipynb:335-337; pipeline defined at
 * databricks/src/01-train-model.ipynb:195-231) computes inside scikit-learn
 * (un-vendored; pinned scikit-learn==1.1.1 at reference app/requirements.txt:14):
 *
 *   1. categorical: unknown / missing -> all-zero one-hot block (01-train-model.ipynb:200-206)
 *   2. numeric: NaN -> training median                            (01-train-model.ipynb:212)
 *   3. row -> 85 float32 values          (sklearn tree/_classes.py _validate_X_predict: dtype float32)
 *   4. per tree: X[i, feature] <= threshold ? left : right, X float32 vs threshold float64
 *                                         (sklearn tree/_tree.pyx _apply_dense)
 *   5. RF: float64 sum of leaf class-1 fractions, / n_trees; label = p1 > p0
 *                                         (sklearn ensemble/_forest.py predict_proba / predict)
 *      GBDT (BASELINE configs 2-4; no reference counterpart): raw = init + sum scale*value in
 *      tree order, proba = expit(raw), label = raw >= 0
 *                                         (sklearn ensemble/_gradient_boosting.pyx predict_stages, _gb.py)
 *
 * Used by tests as a second, independent checker and by bench.py's cpu_baseline leg as the
 * multi-threaded CPU "port".  Never linked into, or called from, the product library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_MAX_COLS 512

int oracle_forest_predict(int kind, int n_trees, double init_raw, double scale, int n_cat, int n_num,
                          const int32_t *cat_offsets, /* n_cat + 1 */
                          const double *medians,      /* n_num */
                          const int64_t *tree_off,    /* n_trees + 1 */
                          const int32_t *left, const int32_t *right, const int32_t *feature,
                          const double *threshold, const double *value,
                          const int32_t *codes, /* n x n_cat, -1 = unknown */
                          const double *nums,   /* n x n_num, NaN = missing */
                          int64_t n, double *proba1, int32_t *label, int threads) {
    const int n_ohe = cat_offsets[n_cat];
    const int n_cols = n_ohe + n_num;
    if (n_cols > ORACLE_MAX_COLS) return -1;
    int bad = 0;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int64_t i = 0; i < n; ++i) {
        float x[ORACLE_MAX_COLS];
        memset(x, 0, sizeof(float) * (size_t)n_cols);
        for (int j = 0; j < n_cat; ++j) {
            int32_t c = codes[i * n_cat + j];
            if (c >= 0 && c < cat_offsets[j + 1] - cat_offsets[j]) x[cat_offsets[j] + c] = 1.0f;
        }
        for (int j = 0; j < n_num; ++j) {
            double v = nums[i * n_num + j];
            if (isnan(v)) v = medians[j];
            float f = (float)v; /* round-to-nearest-even, as numpy astype(float32) */
            if (!isfinite(f)) bad |= 1;
            x[n_ohe + j] = f;
        }
        double acc = (kind == 1) ? init_raw : 0.0;
        for (int t = 0; t < n_trees; ++t) {
            const int64_t o = tree_off[t];
            int32_t node = 0;
            while (left[o + node] != -1) {
                if ((double)x[feature[o + node]] <= threshold[o + node])
                    node = left[o + node];
                else
                    node = right[o + node];
            }
            acc += (kind == 1) ? scale * value[o + node] : value[o + node];
        }
        if (kind == 0) {
            const double s1 = acc, s0 = (double)n_trees - acc;
            proba1[i] = s1 / (double)n_trees;
            label[i] = s1 > s0;
        } else {
            proba1[i] = 1.0 / (1.0 + exp(-acc));
            label[i] = acc >= 0.0;
        }
    }
    return bad ? -2 : 0;
}
