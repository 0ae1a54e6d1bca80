/*
 * CPU oracle (C restatement) for the GCN-layer hot path of LinChen-65/pygcn.
 *
 * TEST INFRASTRUCTURE ONLY: linked/loaded by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.  Never by pygcn_b200/.
 *
 * Restates, with plain loops:
 *   forward   pygcn/layers.py:33  support = mm(input, weight)
 *             pygcn/layers.py:34  output  = spmm(adj, support)   (adj in COO as built by
 *                                 pygcn/utils.py:407-414: the sparse product is a per-stored-
 *                                 entry axpy  out[row] += val * support[col]  in storage order)
 *             pygcn/layers.py:36  output + bias
 *   backward  autograd of the three lines above:
 *             db = sum_rows G ; dS = adj^T G ; dW = X^T dS ; dX = dS W^T
 *
 * The arithmetic of the reference lives in PyTorch (third party, not vendored under
 * /root/reference); this file follows the mathematical definition at the call sites and is
 * pinned by tests/golden/ (outputs of the reference run in the build container).
 *
 * Build: gcc -O2 -fopenmp -shared -fPIC -std=c11 -o oracle/_build/libgcn_oracle.so oracle/gcn_oracle.c -lm
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DEFINE_LAYER(T, SUF)                                                                      \
  static void mm_##SUF(int64_t m, int64_t k, int64_t n, const T* a, const T* b, T* c) {           \
    _Pragma("omp parallel for schedule(static)") for (int64_t i = 0; i < m; ++i) {                \
      T* ci = c + i * n;                                                                          \
      for (int64_t j = 0; j < n; ++j) ci[j] = (T)0;                                               \
      for (int64_t p = 0; p < k; ++p) {                                                           \
        const T aip = a[i * k + p];                                                               \
        const T* bp = b + p * n;                                                                  \
        for (int64_t j = 0; j < n; ++j) ci[j] += aip * bp[j];                                     \
      }                                                                                           \
    }                                                                                             \
  }                                                                                               \
  /* layers.py:33-36 */                                                                           \
  void oracle_layer_forward_##SUF(int64_t n_rows, int64_t n_cols, int64_t fin, int64_t fout,      \
                                  const T* x, int64_t nnz, const int64_t* row, const int64_t* col,\
                                  const T* val, const T* w, const T* bias, T* support, T* out) {  \
    mm_##SUF(n_cols, fin, fout, x, w, support);                                                   \
    memset(out, 0, sizeof(T) * (size_t)n_rows * (size_t)fout);                                    \
    for (int64_t e = 0; e < nnz; ++e) { /* serial COO axpy, storage order */                      \
      T* o = out + row[e] * fout;                                                                 \
      const T* s = support + col[e] * fout;                                                       \
      const T v = val[e];                                                                         \
      for (int64_t j = 0; j < fout; ++j) o[j] += v * s[j];                                        \
    }                                                                                             \
    if (bias)                                                                                     \
      for (int64_t i = 0; i < n_rows; ++i)                                                        \
        for (int64_t j = 0; j < fout; ++j) out[i * fout + j] += bias[j];                          \
  }                                                                                               \
  /* autograd of layers.py:33-36 */                                                               \
  void oracle_layer_backward_##SUF(int64_t n_rows, int64_t n_cols, int64_t fin, int64_t fout,     \
                                   const T* x, int64_t nnz, const int64_t* row,                   \
                                   const int64_t* col, const T* val, const T* w, const T* g,      \
                                   T* ds, T* dw, T* db, T* dx) {                                  \
    for (int64_t j = 0; j < fout; ++j) db[j] = (T)0;                                              \
    for (int64_t i = 0; i < n_rows; ++i)                                                          \
      for (int64_t j = 0; j < fout; ++j) db[j] += g[i * fout + j];                                \
    memset(ds, 0, sizeof(T) * (size_t)n_cols * (size_t)fout);                                     \
    for (int64_t e = 0; e < nnz; ++e) { /* adj^T G */                                             \
      T* o = ds + col[e] * fout;                                                                  \
      const T* s = g + row[e] * fout;                                                             \
      const T v = val[e];                                                                         \
      for (int64_t j = 0; j < fout; ++j) o[j] += v * s[j];                                        \
    }                                                                                             \
    /* dW[p][j] = sum_i x[i][p] ds[i][j] */                                                       \
    memset(dw, 0, sizeof(T) * (size_t)fin * (size_t)fout);                                        \
    for (int64_t i = 0; i < n_cols; ++i)                                                          \
      for (int64_t p = 0; p < fin; ++p) {                                                         \
        const T xip = x[i * fin + p];                                                             \
        for (int64_t j = 0; j < fout; ++j) dw[p * fout + j] += xip * ds[i * fout + j];            \
      }                                                                                           \
    /* dX[i][p] = sum_j ds[i][j] w[p][j] */                                                       \
    _Pragma("omp parallel for schedule(static)") for (int64_t i = 0; i < n_cols; ++i)             \
      for (int64_t p = 0; p < fin; ++p) {                                                         \
        T acc = (T)0;                                                                             \
        for (int64_t j = 0; j < fout; ++j) acc += ds[i * fout + j] * w[p * fout + j];             \
        dx[i * fin + p] = acc;                                                                    \
      }                                                                                           \
  }

DEFINE_LAYER(float, f32)
DEFINE_LAYER(double, f64)

/*
 * Multi-threaded CSR form of the same layer (fwd + bwd), used only as a *timed CPU baseline*
 * ("port", all host threads) next to the reference-as-written COO loop.  int32 indices.
 * t_* is the CSR of adj^T.  Returns 0.
 */
static void csr_spmm_f32(int64_t n, int64_t f, const int32_t* rowptr, const int32_t* col,
                         const float* val, const float* b, const float* bias, float* out) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n; ++i) {
    float* o = out + i * f;
    for (int64_t j = 0; j < f; ++j) o[j] = bias ? bias[j] : 0.0f;
    for (int32_t e = rowptr[i]; e < rowptr[i + 1]; ++e) {
      const float v = val[e];
      const float* s = b + (int64_t)col[e] * f;
      for (int64_t j = 0; j < f; ++j) o[j] += v * s[j];
    }
  }
}

int oracle_csr_layer_fwdbwd_f32(int64_t n, int64_t fin, int64_t fout, const int32_t* rowptr,
                                const int32_t* col, const float* val, const int32_t* t_rowptr,
                                const int32_t* t_col, const float* t_val, const float* x,
                                const float* w, const float* bias, const float* g, float* support,
                                float* out, float* ds, float* dw, float* db, float* dx) {
  mm_f32(n, fin, fout, x, w, support);
  csr_spmm_f32(n, fout, rowptr, col, val, support, bias, out);
  for (int64_t j = 0; j < fout; ++j) db[j] = 0.0f;
  for (int64_t i = 0; i < n; ++i)
    for (int64_t j = 0; j < fout; ++j) db[j] += g[i * fout + j];
  csr_spmm_f32(n, fout, t_rowptr, t_col, t_val, g, NULL, ds);
  memset(dw, 0, sizeof(float) * (size_t)fin * (size_t)fout);
#pragma omp parallel
  {
    float* loc = (float*)calloc((size_t)fin * (size_t)fout, sizeof(float));
#pragma omp for schedule(static)
    for (int64_t i = 0; i < n; ++i)
      for (int64_t p = 0; p < fin; ++p) {
        const float xip = x[i * fin + p];
        for (int64_t j = 0; j < fout; ++j) loc[p * fout + j] += xip * ds[i * fout + j];
      }
#pragma omp critical
    for (int64_t q = 0; q < fin * fout; ++q) dw[q] += loc[q];
    free(loc);
  }
  if (dx) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
      for (int64_t p = 0; p < fin; ++p) {
        float acc = 0.0f;
        for (int64_t j = 0; j < fout; ++j) acc += ds[i * fout + j] * w[p * fout + j];
        dx[i * fin + p] = acc;
      }
  }
  return 0;
}
