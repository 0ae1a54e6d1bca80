// (ReLU ->) fresh training-mode BatchNorm1d over the layer output, forward and backward.
//
// Reference: pygcn/models.py:41-45 `apply_bn` constructs a NEW nn.BatchNorm1d(F) on every call (affine weight 1, bias 0,
// training mode even under model.eval()) and models.py:49,53 apply it as  apply_bn(F.relu(gc(x, adj))) :
//     a = relu(y);  mean_c = mean_rows(a);  var_c = mean_rows((a - mean_c)^2)  (biased);
//     out = (a - mean_c) / sqrt(var_c + eps),  eps = 1e-5
// and its autograd backward (xh = out):
//     da = rstd_c * (g - mean_rows(g) - xh * mean_rows(g * xh));  dy = da * [y > 0].
// With torch that is ReLU (read + write), BatchNorm (statistics pass + normalise pass), and in backward a reduce pass,
// an apply pass and the ReLU mask pass: 13 panel reads / writes.  Here: one statistics pass + one apply pass each way,
// the ReLU and its mask folded into both: 8.  The layer output was just written by the SpMM, so on B200 the statistics pass reads it out of the
// 126 MB L2, not HBM, for every BASELINE shape up to 30 M elements.
//
// Deterministic: per-thread fp64 accumulation, per-CTA partials combined in CTA order by a warp per column, no atomics.
// Written at the end of round 1 after the GPU minutes were spent: compiled for sm_100a, NOT yet run on hardware;
// nothing on the measured layer path calls it (own translation unit, own entry points).
#include "common.cuh"

namespace gcnb {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 4 * kNumSMs;

int bn_blocks(int64_t n_rows) {
  int64_t b = ceil_div(n_rows, 128);
  if (b > kMaxBlocks) b = kMaxBlocks;
  return (int)(b < 1 ? 1 : b);
}

template <int V>
struct Vec;
template <>
struct Vec<1> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
  static __device__ __forceinline__ void store(float* p, const float (&v)[1]) { p[0] = v[0]; }
};
template <>
struct Vec<4> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[4]) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};

// Statistics pass.  MODE 0 (forward): (sum a, sum a^2) per column, a = relu ? max(y, 0) : y.
// MODE 1 (backward): (sum g, sum g * xh) per column, xh = (a - mean) * rstd.
// CTA b owns rows [b * rows_per_block, ...); thread (ty, tx) walks rows ty, ty + rl, ... of column group tx (V columns),
// four rows in flight; partial[b][0 | 1][f] in fp64.
template <int MODE, int V>
__global__ void __launch_bounds__(kThreads)
bn_stats_kernel(int64_t n_rows, int fv, int cw, int64_t rows_per_block, const float* __restrict__ y, int64_t ldy, int relu,
                const float* __restrict__ g, int64_t ldg, const float* __restrict__ mean, const float* __restrict__ rstd,
                double* __restrict__ partial) {
  __shared__ double red[2][V][kThreads];
  const int tx = threadIdx.x % cw;  // column-group lane
  const int ty = threadIdx.x / cw;  // row lane
  const int rl = kThreads / cw;
  const int f = fv * V;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = (r0 + rows_per_block < n_rows) ? (r0 + rows_per_block) : n_rows;
  for (int j0 = 0; j0 < fv; j0 += cw) {
    const int j = j0 + tx;
    double s1[V], s2[V];
#pragma unroll
    for (int t = 0; t < V; ++t) s1[t] = s2[t] = 0.0;
    if (j < fv) {
      float m[V], rs[V];
#pragma unroll
      for (int t = 0; t < V; ++t) {
        m[t] = (MODE == 1) ? __ldg(mean + j * V + t) : 0.f;
        rs[t] = (MODE == 1) ? __ldg(rstd + j * V + t) : 1.f;
      }
      auto add = [&](const float (&yv)[V], const float (&gv)[V]) {
#pragma unroll
        for (int t = 0; t < V; ++t) {
          const float a = relu ? fmaxf(yv[t], 0.f) : yv[t];
          if (MODE == 0) {
            s1[t] += (double)a;
            s2[t] += (double)a * (double)a;
          } else {
            const float xh = (a - m[t]) * rs[t];
            s1[t] += (double)gv[t];
            s2[t] += (double)gv[t] * (double)xh;
          }
        }
      };
      int64_t r = r0 + ty;
      for (; r + 3 * rl < r1; r += 4 * rl) {  // four rows in flight per thread
        float yv[4][V], gv[4][V];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          Vec<V>::load(y + (r + u * rl) * ldy + (int64_t)j * V, yv[u]);
          if (MODE == 1) Vec<V>::load(g + (r + u * rl) * ldg + (int64_t)j * V, gv[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) add(yv[u], MODE == 1 ? gv[u] : yv[u]);
      }
      for (; r < r1; r += rl) {
        float yv[V], gv[V];
        Vec<V>::load(y + r * ldy + (int64_t)j * V, yv);
        if (MODE == 1) Vec<V>::load(g + r * ldg + (int64_t)j * V, gv);
        add(yv, MODE == 1 ? gv : yv);
      }
    }
#pragma unroll
    for (int t = 0; t < V; ++t) {
      red[0][t][threadIdx.x] = s1[t];
      red[1][t][threadIdx.x] = s2[t];
    }
    __syncthreads();
    if (ty == 0 && j < fv) {
#pragma unroll
      for (int t = 0; t < V; ++t) {
        double a1 = 0.0, a2 = 0.0;
        for (int k = 0; k < rl; ++k) {  // row lanes in order: fixed summation order
          a1 += red[0][t][k * cw + tx];
          a2 += red[1][t][k * cw + tx];
        }
        partial[((int64_t)blockIdx.x * 2 + 0) * f + j * V + t] = a1;
        partial[((int64_t)blockIdx.x * 2 + 1) * f + j * V + t] = a2;
      }
    }
    __syncthreads();
  }
}

// One warp per column adds the CTA partials: lane p takes parts p, p + 32, ... in order, then a fixed xor tree.
// MODE 0: stat0 = mean, stat1 = rstd = 1 / sqrt(max(E[a^2] - mean^2, 0) + eps)   (fp64 until the final rounding)
// MODE 1: stat0 = mean_rows(g), stat1 = mean_rows(g * xh)
template <int MODE>
__global__ void __launch_bounds__(kThreads)
bn_finalize_kernel(int n_parts, int f, double inv_n, double eps, const double* __restrict__ partial, float* __restrict__ stat0,
                   float* __restrict__ stat1) {
  const int lane = threadIdx.x & 31;
  const int c = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
  if (c >= f) return;  // (whole warps leave together)
  double a1 = 0.0, a2 = 0.0;
  for (int p = lane; p < n_parts; p += 32) {
    a1 += partial[((int64_t)p * 2 + 0) * f + c];
    a2 += partial[((int64_t)p * 2 + 1) * f + c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if (lane == 0) {
    if (MODE == 0) {
      const double mu = a1 * inv_n;
      double var = a2 * inv_n - mu * mu;
      if (!(var > 0.0)) var = 0.0;
      stat0[c] = (float)mu;
      stat1[c] = (float)(1.0 / sqrt(var + eps));
    } else {
      stat0[c] = (float)(a1 * inv_n);
      stat1[c] = (float)(a2 * inv_n);
    }
  }
}

// Apply pass.  MODE 0: out = (a - mean) * rstd.  MODE 1: dy = [relu: y > 0] * rstd * (g - gbar - xh * gxbar).
// Same thread layout as the statistics pass (cw column-group lanes x rl row lanes, no index divisions); the per-column
// constants are loaded once per column tile.
template <int MODE, int V>
__global__ void __launch_bounds__(kThreads)
bn_apply_kernel(int64_t n_rows, int fv, int cw, const float* __restrict__ y, int64_t ldy, int relu, const float* __restrict__ g,
                int64_t ldg, const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gbar,
                const float* __restrict__ gxbar, float* __restrict__ out, int64_t ldo) {
  const int tx = threadIdx.x % cw;
  const int ty = threadIdx.x / cw;
  const int rl = kThreads / cw;
  for (int j0 = 0; j0 < fv; j0 += cw) {
    const int j = j0 + tx;
    if (j >= fv) continue;
    float m[V], rs[V], gb[V], gxb[V];
    Vec<V>::load(mean + j * V, m);
    Vec<V>::load(rstd + j * V, rs);
    if (MODE == 1) {
      Vec<V>::load(gbar + j * V, gb);
      Vec<V>::load(gxbar + j * V, gxb);
    }
    for (int64_t r = (int64_t)blockIdx.x * rl + ty; r < n_rows; r += (int64_t)gridDim.x * rl) {
      float yv[V], gv[V], o[V];
      Vec<V>::load(y + r * ldy + (int64_t)j * V, yv);
      if (MODE == 1) Vec<V>::load(g + r * ldg + (int64_t)j * V, gv);
#pragma unroll
      for (int t = 0; t < V; ++t) {
        const float a = relu ? fmaxf(yv[t], 0.f) : yv[t];
        const float xh = (a - m[t]) * rs[t];
        if (MODE == 0) {
          o[t] = xh;
        } else {
          const float da = rs[t] * (gv[t] - gb[t] - xh * gxb[t]);
          o[t] = (relu && !(yv[t] > 0.f)) ? 0.f : da;
        }
      }
      Vec<V>::store(out + r * ldo + (int64_t)j * V, o);
    }
  }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

size_t bn_ws_bytes(int64_t n_rows, int64_t f) { return (size_t)bn_blocks(n_rows) * 2 * (size_t)f * sizeof(double) + 16; }

int lanes_for(int fv) {
  int cw = 1;
  while (cw < fv && cw < kThreads) cw <<= 1;
  return cw;
}

int apply_grid(int64_t n, int cw) {  // CTAs of (kThreads / cw) rows, a few rows per thread
  int64_t b = ceil_div(n, (int64_t)(kThreads / cw) * 4);
  if (b > 16 * kNumSMs) b = 16 * kNumSMs;
  return (int)(b < 1 ? 1 : b);
}

// statistics + finalise of either direction
template <int MODE>
int bn_stats(int64_t n, int64_t f, const float* y, int64_t ldy, int relu, const float* g, int64_t ldg, const float* mean,
             const float* rstd, float eps, float* stat0, float* stat1, bool vec, double* partial, cudaStream_t st) {
  const int nb = bn_blocks(n);
  const int64_t rows_per_block = ceil_div(n, nb);
  const int fv = (int)(vec ? f / 4 : f);
  const int cw = lanes_for(fv);
  if (vec)
    bn_stats_kernel<MODE, 4><<<nb, kThreads, 0, st>>>(n, fv, cw, rows_per_block, y, ldy, relu, g, ldg, mean, rstd, partial);
  else
    bn_stats_kernel<MODE, 1><<<nb, kThreads, 0, st>>>(n, fv, cw, rows_per_block, y, ldy, relu, g, ldg, mean, rstd, partial);
  GCNB_LAUNCH_CHECK();
  bn_finalize_kernel<MODE><<<(unsigned)ceil_div(f, kThreads / 32), kThreads, 0, st>>>(nb, (int)f, 1.0 / (double)n, (double)eps,
                                                                                    partial, stat0, stat1);
  GCNB_LAUNCH_CHECK();
  return GCNB_OK;
}

}  // namespace
}  // namespace gcnb

using namespace gcnb;

extern "C" size_t gcnb_fresh_bn_workspace_bytes(int64_t n_rows, int64_t f) {
  if (n_rows <= 0 || f <= 0) return 16;
  return bn_ws_bytes(n_rows, f);
}

extern "C" int gcnb_fresh_bn_forward(int64_t n_rows, int64_t f, const float* d_y, int64_t ldy, int relu, float eps,
                                     float* d_out, int64_t ldo, float* d_mean, float* d_rstd, void* d_ws, size_t ws_bytes,
                                     void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GCNB_REQUIRE(n_rows >= 0 && f > 0 && f < (1 << 24), "fresh_bn_forward: shape out of range");
  GCNB_REQUIRE(eps >= 0.f, "fresh_bn_forward: eps must be >= 0");
  if (n_rows == 0) return GCNB_OK;
  GCNB_REQUIRE(d_y && d_out && d_mean && d_rstd, "fresh_bn_forward: null operand");
  GCNB_REQUIRE(ldy >= f && ldo >= f, "fresh_bn_forward: leading dimension smaller than width");
  const size_t need = bn_ws_bytes(n_rows, f);
  GCNB_REQUIRE(d_ws != nullptr && ws_bytes >= need && (reinterpret_cast<uintptr_t>(d_ws) & 7u) == 0,
               "fresh_bn_forward: workspace too small or misaligned (%zu < %zu)", ws_bytes, need);
  const bool vec = f % 4 == 0 && ldy % 4 == 0 && ldo % 4 == 0 && al16(d_y) && al16(d_out) && al16(d_mean) && al16(d_rstd);
  double* partial = reinterpret_cast<double*>(d_ws);
  GCNB_TRY((bn_stats<0>(n_rows, f, d_y, ldy, relu ? 1 : 0, nullptr, 0, nullptr, nullptr, eps, d_mean, d_rstd, vec, partial, st)));
  const int fv = (int)(vec ? f / 4 : f);
  const int cw = lanes_for(fv);
  if (vec)
    bn_apply_kernel<0, 4><<<apply_grid(n_rows, cw), kThreads, 0, st>>>(n_rows, fv, cw, d_y, ldy, relu ? 1 : 0, nullptr, 0, d_mean,
                                                                       d_rstd, nullptr, nullptr, d_out, ldo);
  else
    bn_apply_kernel<0, 1><<<apply_grid(n_rows, cw), kThreads, 0, st>>>(n_rows, fv, cw, d_y, ldy, relu ? 1 : 0, nullptr, 0, d_mean,
                                                                       d_rstd, nullptr, nullptr, d_out, ldo);
  GCNB_LAUNCH_CHECK();
  return GCNB_OK;
}

extern "C" int gcnb_fresh_bn_backward(int64_t n_rows, int64_t f, const float* d_y, int64_t ldy, int relu, const float* d_g,
                                      int64_t ldg, const float* d_mean, const float* d_rstd, float* d_dy, int64_t lddy,
                                      float* d_gstat, void* d_ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GCNB_REQUIRE(n_rows >= 0 && f > 0 && f < (1 << 24), "fresh_bn_backward: shape out of range");
  if (n_rows == 0) return GCNB_OK;
  GCNB_REQUIRE(d_y && d_g && d_mean && d_rstd && d_dy && d_gstat, "fresh_bn_backward: null operand");
  GCNB_REQUIRE(ldy >= f && ldg >= f && lddy >= f, "fresh_bn_backward: leading dimension smaller than width");
  const size_t need = bn_ws_bytes(n_rows, f);
  GCNB_REQUIRE(d_ws != nullptr && ws_bytes >= need && (reinterpret_cast<uintptr_t>(d_ws) & 7u) == 0,
               "fresh_bn_backward: workspace too small or misaligned (%zu < %zu)", ws_bytes, need);
  // d_gstat: [2][f] floats (mean_rows(g), mean_rows(g * xh)), rows 16-byte aligned on the vector path
  float* gbar = d_gstat;
  float* gxbar = d_gstat + f;
  const bool vec = f % 4 == 0 && ldy % 4 == 0 && ldg % 4 == 0 && lddy % 4 == 0 && al16(d_y) && al16(d_g) && al16(d_dy) &&
                   al16(d_mean) && al16(d_rstd) && al16(d_gstat);
  double* partial = reinterpret_cast<double*>(d_ws);
  GCNB_TRY((bn_stats<1>(n_rows, f, d_y, ldy, relu ? 1 : 0, d_g, ldg, d_mean, d_rstd, 0.f, gbar, gxbar, vec, partial, st)));
  const int fv = (int)(vec ? f / 4 : f);
  const int cw = lanes_for(fv);
  if (vec)
    bn_apply_kernel<1, 4><<<apply_grid(n_rows, cw), kThreads, 0, st>>>(n_rows, fv, cw, d_y, ldy, relu ? 1 : 0, d_g, ldg, d_mean,
                                                                       d_rstd, gbar, gxbar, d_dy, lddy);
  else
    bn_apply_kernel<1, 1><<<apply_grid(n_rows, cw), kThreads, 0, st>>>(n_rows, fv, cw, d_y, ldy, relu ? 1 : 0, d_g, ldg, d_mean,
                                                                       d_rstd, gbar, gxbar, d_dy, lddy);
  GCNB_LAUNCH_CHECK();
  return GCNB_OK;
}
