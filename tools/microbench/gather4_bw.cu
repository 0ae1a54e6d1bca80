// Microbenchmark: row gathers with the Blackwell TMA gather4 load (cp.async.bulk.tensor.2d ... tile::gather4: four rows
// of a 2-D tensor, chosen by four row coordinates, land in shared memory with one instruction) against 128-bit register
// loads, on the access pattern of the wide-row SpMM: 1 KB (or narrower) panel rows picked by a column stream.
// The question it answers: with the bytes in flight held in shared memory (a ring per warp, mbarrier per stage) instead
// of registers, how many TB/s of gathered rows does an SM pull from L2 / HBM, and is the TMA unit's row rate a limit?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o gather4_bw gather4_bw.cu
//   ./gather4_bw <rows> <width_floats> <ld_floats> <stages> <warps_per_cta> <hot_rows> <hot_pct> [box_rows]
//     rows       table rows (48000 = L2 resident at 1 KB rows, 2400000 = the products panel)
//     hot_rows / hot_pct: hot_pct % of the references fall on the first hot_rows rows (0 0 = uniform)
//     box_rows   second box dimension of the tensor map (1 = what CuTe encodes for gather4)
//     persist_mb > 0: cudaLimitPersistingL2CacheSize of that size + a persisting access-policy window over the hot rows
// Prints: correctness of the gather4 kernel against the register kernel (same partial sums), then time and TB/s of both.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__);      \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity))
    if (++spins > (1u << 26)) __trap();  // a protocol error traps instead of hanging the box
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const void* tmap, int x, int r0, int r1, int r2, int r3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
      "l"(tmap), "r"(x), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}

// One warp = one contiguous range of 4-row groups.  Ring of `stages` slots of 4 rows each, one mbarrier per slot; the
// warp issues group g + stages (lane 0) right after it has consumed group g.
template <int CH>
__global__ void __launch_bounds__(256) k_gather4(const __grid_constant__ CUtensorMap tm, const int* __restrict__ idx,
                                                 long groups_per_warp, int w, int stages, float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const long warp = (long)blockIdx.x * wpb + wib;
  const uint32_t row_b = (uint32_t)w * 4u, stage_b = 4u * row_b;
  const uint32_t ring = smem_u32(smem) + (uint32_t)wib * (uint32_t)stages * stage_b;
  const uint32_t bars = smem_u32(smem) + (uint32_t)wpb * (uint32_t)stages * stage_b + (uint32_t)wib * (uint32_t)stages * 8u;
  if (lane == 0)
    for (int s = 0; s < stages; ++s) mbar_init(bars + 8u * s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int* my = idx + warp * groups_per_warp * 4;
  const int nch = w / 4;
  float4 acc[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  int blk = 0, blk_next = __ldg(my + lane);
  auto issue = [&](long gi) {  // all lanes
    if ((gi & 7) == 0) {
      blk = blk_next;
      if ((gi + 8) * 4 + lane < groups_per_warp * 4) blk_next = __ldg(my + (gi + 8) * 4 + lane);
    }
    const int j = (int)(gi & 7) * 4;
    const int c0 = __shfl_sync(kFull, blk, j), c1 = __shfl_sync(kFull, blk, j + 1), c2 = __shfl_sync(kFull, blk, j + 2),
              c3 = __shfl_sync(kFull, blk, j + 3);
    if (lane == 0) {
      const uint32_t s = (uint32_t)(gi % stages);
      mbar_expect_tx(bars + 8u * s, stage_b);
      tma_gather4(ring + s * stage_b, &tm, 0, c0, c1, c2, c3, bars + 8u * s);
    }
  };
  const long pre = groups_per_warp < stages ? groups_per_warp : stages;
  for (long g = 0; g < pre; ++g) issue(g);
  for (long g = 0; g < groups_per_warp; ++g) {
    const uint32_t s = (uint32_t)(g % stages);
    mbar_wait(bars + 8u * s, (uint32_t)((g / stages) & 1));
    const uint32_t base = ring + s * stage_b;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        const int q = lane + 32 * k;
        if (q < nch) {
          const float4 x = lds128(base + r * row_b + 16u * q);
          acc[k].x += x.x; acc[k].y += x.y; acc[k].z += x.z; acc[k].w += x.w;
        }
      }
    }
    __syncwarp();
    if (g + stages < groups_per_warp) issue(g + stages);
  }
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int q = lane + 32 * k;
    if (q < nch) *reinterpret_cast<float4*>(out + warp * w + 4 * q) = acc[k];
  }
}

// The same sums with 128-bit register loads: U rows in flight per lane (what spmm_stream_kernel does today).
template <int CH, int U>
__global__ void __launch_bounds__(128) k_ldg(const float* __restrict__ t, long ld, const int* __restrict__ idx, long groups_per_warp,
                                             int w, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int* my = idx + warp * groups_per_warp * 4;
  const long n = groups_per_warp * 4;
  const int nch = w / 4;
  float4 acc[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long b = 0; b < n; b += 32) {
    const int mine = (b + lane < n) ? __ldg(my + b + lane) : 0;
    const int cnt = (int)((n - b) < 32 ? (n - b) : 32);
#pragma unroll 1
    for (int j = 0; j < cnt; j += U) {
      float4 x[U][CH];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = __shfl_sync(kFull, mine, (j + u) & 31);
#pragma unroll
        for (int k = 0; k < CH; ++k) {
          const int q = min(lane + 32 * k, nch - 1);
          x[u][k] = __ldg(reinterpret_cast<const float4*>(t + (long)c * ld) + q);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (j + u < cnt) {
#pragma unroll
          for (int k = 0; k < CH; ++k) { acc[k].x += x[u][k].x; acc[k].y += x[u][k].y; acc[k].z += x[u][k].z; acc[k].w += x[u][k].w; }
        }
    }
  }
#pragma unroll
  for (int k = 0; k < CH; ++k) {
    const int q = lane + 32 * k;
    if (q < nch) *reinterpret_cast<float4*>(out + warp * w + 4 * q) = acc[k];
  }
}

__global__ void k_fill(float* t, long n, uint32_t seed) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    t[i] = (float)(h & 0xffff) * (1.f / 65536.f) - 0.5f;
  }
}
__global__ void k_flush(float4* p, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  if (argc < 8) { printf("usage: rows width ld stages warps_per_cta hot_rows hot_pct [box_rows]\n"); return 2; }
  const long rows = atol(argv[1]);
  const int w = atoi(argv[2]), ld = atoi(argv[3]), stages = atoi(argv[4]), wpb = atoi(argv[5]);
  const long hot_rows = atol(argv[6]);
  const int hot_pct = atoi(argv[7]);
  const int box_rows = argc > 8 ? atoi(argv[8]) : 1;
  const int persist_mb = argc > 9 ? atoi(argv[9]) : 0;  // > 0: L2 persisting carve-out of that size + an access policy window over the hot rows
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  const size_t smem = (size_t)wpb * stages * 16 * w + (size_t)wpb * stages * 8;
  if (smem > 227 * 1024) { printf("ring too large: %zu bytes\n", smem); return 2; }
  const int ctas_per_sm = (int)((227 * 1024) / (smem + 1024)) > 0 ? (int)((227 * 1024) / (smem + 1024)) : 1;
  const int ctas = sms * (ctas_per_sm > 8 ? 8 : ctas_per_sm);
  const long warps = (long)ctas * wpb;
  const long groups_per_warp = 2048;  // 8192 rows per warp
  const long m = warps * groups_per_warp * 4;
  printf("rows %ld width %d ld %d stages %d warps/cta %d ctas %d (%d per SM) ring %zu B/cta  entries %ld  hot %ld rows %d%%  box {%d,%d}\n",
         rows, w, ld, stages, wpb, ctas, ctas / sms, smem, m, hot_rows, hot_pct, w, box_rows);

  float* t;
  CK(cudaMalloc(&t, (size_t)rows * ld * 4));
  k_fill<<<sms * 8, 256>>>(t, rows * ld, 12345u);
  std::vector<int> h(m);
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return s; };
  for (long i = 0; i < m; ++i) {
    const bool hot = hot_rows > 0 && (int)(rnd() % 100) < hot_pct;
    h[i] = (int)(rnd() % (uint64_t)(hot ? hot_rows : rows));
  }
  int* idx;
  CK(cudaMalloc(&idx, m * 4));
  CK(cudaMemcpy(idx, h.data(), m * 4, cudaMemcpyHostToDevice));
  float *o1, *o2;
  CK(cudaMalloc(&o1, warps * w * 4));
  CK(cudaMalloc(&o2, warps * w * 4 * 8));
  CK(cudaMemset(o1, 0, warps * w * 4));
  CK(cudaMemset(o2, 0xff, warps * w * 4));
  float4* fl;
  const long fl_n = (512l << 20) / 16;
  CK(cudaMalloc(&fl, fl_n * 16));

  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {(cuuint64_t)w, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {(cuuint32_t)w, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = reinterpret_cast<EncodeTiledFn>(fp)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, t, gdim, gstr, box, estr,
                                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); return 1; }

  if (persist_mb > 0 && hot_rows > 0) {
    // Are pinned hub rows worth anything?  The hot rows are the first hot_rows rows of the table: one contiguous window,
    // hitProp = persisting (they stay in the carve-out), everything else streams.
    printf("persistingL2CacheMaxSize %d MB, accessPolicyMaxWindowSize %d MB, L2 %d MB\n", prop.persistingL2CacheMaxSize >> 20,
           prop.accessPolicyMaxWindowSize >> 20, prop.l2CacheSize >> 20);
    CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)persist_mb << 20));
    cudaStreamAttrValue av = {};
    av.accessPolicyWindow.base_ptr = t;
    av.accessPolicyWindow.num_bytes = (size_t)hot_rows * ld * 4;
    av.accessPolicyWindow.hitRatio = 1.0f;
    av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    CK(cudaStreamSetAttribute(0, cudaStreamAttributeAccessPolicyWindow, &av));
  }
  const int ch = (w / 4 + 31) / 32;
  auto run_g4 = [&]() {
    if (ch == 1) {
      CK(cudaFuncSetAttribute(k_gather4<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_gather4<1><<<ctas, wpb * 32, smem>>>(tm, idx, groups_per_warp, w, stages, o1);
    } else {
      CK(cudaFuncSetAttribute(k_gather4<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      k_gather4<2><<<ctas, wpb * 32, smem>>>(tm, idx, groups_per_warp, w, stages, o1);
    }
  };
  // split = 1: the gather4 kernel's partition (for the check); 8: 1024 entries per warp, the SpMM's item size
  auto run_ldg = [&](int u, int split) {
    const int grid = (int)(warps * split / 4);
    const long gpw = groups_per_warp / split;
    if (ch == 1) {
      if (u == 4) k_ldg<1, 4><<<grid, 128>>>(t, ld, idx, gpw, w, o2);
      else k_ldg<1, 8><<<grid, 128>>>(t, ld, idx, gpw, w, o2);
    } else {
      if (u == 4) k_ldg<2, 4><<<grid, 128>>>(t, ld, idx, gpw, w, o2);
      else k_ldg<2, 8><<<grid, 128>>>(t, ld, idx, gpw, w, o2);
    }
  };
  run_ldg(8, 1);
  CK(cudaDeviceSynchronize());
  run_g4();
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> a(warps * w), b(warps * w);
  CK(cudaMemcpy(a.data(), o1, warps * w * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(b.data(), o2, warps * w * 4, cudaMemcpyDeviceToHost));
  double worst = 0, scale = 0;
  for (long i = 0; i < warps * w; ++i) {
    const double d = fabs((double)a[i] - (double)b[i]);
    if (d > worst) worst = d;
    if (fabs((double)b[i]) > scale) scale = fabs((double)b[i]);
  }
  printf("check gather4 vs ldg: max |diff| %.3e (scale %.3e) -> %s\n", worst, scale, worst <= 1e-3 * scale ? "OK" : "MISMATCH");

  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  auto time_it = [&](const char* name, auto fn) {
    float best = 1e9f, sum = 0.f;
    const int reps = 5;
    for (int i = 0; i < reps + 1; ++i) {
      k_flush<<<sms * 4, 256>>>(fl, fl_n);
      CK(cudaEventRecord(e0));
      fn();
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if (i > 0) { sum += ms; if (ms < best) best = ms; }
    }
    const double bytes = (double)m * w * 4;
    printf("%-28s %8.3f ms (min %.3f)  %6.2f TB/s gathered  %6.1f Grows/s\n", name, sum / reps, best, bytes / (sum / reps) / 1e9,
           m / (sum / reps) / 1e6);
  };
  time_it("gather4 ring", run_g4);
  time_it("ldg U=8 (1024-entry items)", [&]() { run_ldg(8, 8); });
  time_it("ldg U=4 (1024-entry items)", [&]() { run_ldg(4, 8); });
  CK(cudaDeviceSynchronize());
  return 0;
}
