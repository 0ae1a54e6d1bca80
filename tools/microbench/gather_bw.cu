// Microbenchmark: random gathers of 128-byte rows from an L2-resident table, the access pattern of
// the F=32 SpMM.  Which instruction shape gets the most bytes per second out of L2?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bw gather_bw.cu && ./gather_bw
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int kRowFloats = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// A: 8 lanes x LDG.128 per row, 4 rows per instruction, U in flight per lane
template <int U, bool NOALLOC>
__global__ void __launch_bounds__(256, 2) k_vec128(const int* __restrict__ idx, int64_t m, const float* __restrict__ t, float* out) {
  const int lane = threadIdx.x & 31, sub = lane & 7, slot = lane >> 3;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float4 acc = make_float4(0, 0, 0, 0);
  const char* base = reinterpret_cast<const char*>(t) + sub * 16;
  for (int64_t b0 = warp * (U * 4); b0 + U * 4 <= m; b0 += nwarps * (U * 4)) {
    int c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = __ldg(idx + b0 + u * 4 + slot);
    float4 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4* p = reinterpret_cast<const float4*>(base + (uint64_t)(uint32_t)c[u] * 128u);
      if (NOALLOC) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x[u].x), "=f"(x[u].y), "=f"(x[u].z), "=f"(x[u].w) : "l"(p));
      } else {
        x[u] = __ldg(p);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) { acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w; }
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

// B: 32 lanes x LDG.32 per row, 1 row per instruction
template <int U>
__global__ void __launch_bounds__(256, 2) k_scalar32(const int* __restrict__ idx, int64_t m, const float* __restrict__ t, float* out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  for (int64_t b0 = warp * 32; b0 + 32 <= m; b0 += nwarps * 32) {
    const int mine = __ldg(idx + b0 + lane);
#pragma unroll
    for (int j = 0; j < 32; j += U) {
      float x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = __shfl_sync(0xffffffffu, mine, j + u);
        x[u] = __ldg(t + (uint64_t)(uint32_t)c * kRowFloats + lane);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) acc += x[u];
    }
  }
  if (acc == 12345.678f) out[0] = acc;
}

// C: 16 lanes x LDG.64 per row, 2 rows per instruction
template <int U>
__global__ void __launch_bounds__(256, 2) k_vec64(const int* __restrict__ idx, int64_t m, const float* __restrict__ t, float* out) {
  const int lane = threadIdx.x & 31, sub = lane & 15, slot = lane >> 4;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float2 acc = make_float2(0, 0);
  const char* base = reinterpret_cast<const char*>(t) + sub * 8;
  for (int64_t b0 = warp * (U * 2); b0 + U * 2 <= m; b0 += nwarps * (U * 2)) {
    int c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) c[u] = __ldg(idx + b0 + u * 2 + slot);
    float2 x[U];
#pragma unroll
    for (int u = 0; u < U; ++u) x[u] = __ldg(reinterpret_cast<const float2*>(base + (uint64_t)(uint32_t)c[u] * 128u));
#pragma unroll
    for (int u = 0; u < U; ++u) { acc.x += x[u].x; acc.y += x[u].y; }
  }
  if (acc.x + acc.y == 12345.678f) out[0] = acc.x;
}

// F: 1-D TMA bulk copies (cp.async.bulk, 128 B per row) into a shared-memory ring, consumed with LDS.
// One warp = one pipeline: lane l issues the copy of row l of a 32-row stage.
template <int STAGES>
__global__ void __launch_bounds__(256, 2) k_bulk(const int* __restrict__ idx, int64_t m, const float* __restrict__ t, float* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  // per warp: STAGES x (32 rows x 128 B) + STAGES mbarriers
  uint8_t* my = smem + (size_t)w * (STAGES * 4096 + 64);
  const uint32_t bar0 = smem_u32(my + STAGES * 4096);
  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8 * s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](int64_t b0, int s) {
    const int c = __ldg(idx + b0 + lane);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * s), "r"(4096) : "memory");
    __syncwarp();
    const float* src = t + (uint64_t)(uint32_t)c * kRowFloats;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(my + s * 4096 + lane * 128)), "l"(src), "r"(128), "r"(bar0 + 8 * s) : "memory");
  };
  float4 acc = make_float4(0, 0, 0, 0);
  int64_t b_issue = warp * 32;
  const int64_t stride = nwarps * 32;
  int n_issued = 0;
  for (; n_issued < STAGES - 1 && b_issue + 32 <= m; ++n_issued, b_issue += stride) issue(b_issue, n_issued);
  int cons = 0;
  uint32_t phase = 0;
  for (int64_t b0 = warp * 32; b0 + 32 <= m; b0 += stride) {
    if (b_issue + 32 <= m) { issue(b_issue, n_issued % STAGES); ++n_issued; b_issue += stride; }
    const int s = cons % STAGES;
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar0 + 8 * s), "r"(phase) : "memory");
    }
    // consume: 32 rows x 32 floats; lane reads float4 #(lane&7) of rows (lane>>3) + 4k
    const float4* st = reinterpret_cast<const float4*>(my + s * 4096);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 x = st[((lane >> 3) + 4 * k) * 8 + (lane & 7)];
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    __syncwarp();
    ++cons;
    if (cons % STAGES == 0) phase ^= 1;
  }
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

// G: the gathers themselves as 16-byte cp.async (LDGSTS) into a per-warp shared-memory ring, read back with LDS.128.
// No destination registers are held while a gather is in flight, so the number of lines in flight is bounded by shared
// memory (STAGES - 1 stages of 32 rows x 128 B per warp), not by the register file -- the question for the SpMM, where
// 64 % of the stall samples wait on the join of four register-held gathers (profiles/r01_spmm_group_v3_source_stalls.txt).
// Price: every gathered line crosses the L1/TEX pipe twice (fill + LDS).  Written at the end of round 1, not yet run.
template <int STAGES>
__global__ void __launch_bounds__(256) k_cpasync(const int* __restrict__ idx, int64_t m, const float* __restrict__ t, float* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, sub = lane & 7, slot = lane >> 3;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  uint8_t* my = smem + (size_t)w * (STAGES * 4096);
  const char* base = reinterpret_cast<const char*>(t) + sub * 16;
  auto issue = [&](int64_t b0, int s) {  // rows slot, slot + 4, ... of the stage; lane moves chunk `sub` of each
    const int mine = __ldg(idx + b0 + lane);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = __shfl_sync(0xffffffffu, mine, slot + 4 * k);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(my + s * 4096 + (slot + 4 * k) * 128 + sub * 16)),
                   "l"(base + (uint64_t)(uint32_t)c * 128u) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float4 acc = make_float4(0, 0, 0, 0);
  const int64_t stride = nwarps * 32;
  int64_t b_issue = warp * 32;
  int n_issued = 0;
  for (; n_issued < STAGES - 1; ++n_issued, b_issue += stride) {
    if (b_issue + 32 <= m) issue(b_issue, n_issued);
    else asm volatile("cp.async.commit_group;" ::: "memory");  // keep the group count uniform
  }
  int cons = 0;
  for (int64_t b0 = warp * 32; b0 + 32 <= m; b0 += stride) {
    if (b_issue + 32 <= m) issue(b_issue, n_issued % STAGES);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    ++n_issued;
    b_issue += stride;
    asm volatile("cp.async.wait_group %0;" ::"n"(STAGES - 1) : "memory");
    __syncwarp();
    const float4* st = reinterpret_cast<const float4*>(my + (cons % STAGES) * 4096);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 x = st[(slot + 4 * k) * 8 + sub];
      acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    __syncwarp();
    ++cons;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (acc.x + acc.y + acc.z + acc.w == 12345.678f) out[0] = acc.x;
}

template <typename F>
float time_ms(F f, void* flush, size_t flush_bytes, int reps = 5) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  float best = 1e9f, sum = 0;
  for (int i = 0; i < reps + 1; ++i) {
    CK(cudaMemsetAsync(flush, i, flush_bytes));
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (i > 0) { sum += ms; if (ms < best) best = ms; }
  }
  return sum / reps;
}

int main(int argc, char** argv) {
  const int64_t n = argc > 1 ? atoll(argv[1]) : 100000;
  const int64_t m = argc > 2 ? atoll(argv[2]) : 10094750 / 256 * 256;
  const int sorted_window = argc > 3 ? atoi(argv[3]) : 0;
  std::vector<int> h(m);
  uint64_t s = 88172645463325252ull;
  for (int64_t i = 0; i < m; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = (int)(s % (uint64_t)n); }
  (void)sorted_window;
  int* idx;
  float *t, *out;
  void* flush;
  const size_t flush_bytes = 512ull << 20;
  CK(cudaMalloc(&idx, m * 4));
  CK(cudaMalloc(&t, n * 128));
  CK(cudaMalloc(&out, 1024));
  CK(cudaMalloc(&flush, flush_bytes));
  CK(cudaMemcpy(idx, h.data(), m * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(t, 0, n * 128));
  const double gb = (double)m * 128 / 1e9;
  auto report = [&](const char* name, float ms) { printf("%-44s %8.3f ms  %7.2f TB/s  %7.1f Grows/s\n", name, ms, gb / ms, m / ms / 1e6); };
  const int sms = 148;
  for (int ctas : {2, 4, 8}) {
    char nm[128];
    const int grid = sms * ctas;
    snprintf(nm, 128, "A vec128 U=8 grid=%dx148", ctas);  report(nm, time_ms([&] { k_vec128<8, false><<<grid, 256>>>(idx, m, t, out); }, flush, flush_bytes));
    snprintf(nm, 128, "A vec128 U=16 grid=%dx148", ctas); report(nm, time_ms([&] { k_vec128<16, false><<<grid, 256>>>(idx, m, t, out); }, flush, flush_bytes));
    snprintf(nm, 128, "A vec128 U=4 grid=%dx148", ctas);  report(nm, time_ms([&] { k_vec128<4, false><<<grid, 256>>>(idx, m, t, out); }, flush, flush_bytes));
    snprintf(nm, 128, "D vec128 no_allocate U=8 grid=%dx148", ctas); report(nm, time_ms([&] { k_vec128<8, true><<<grid, 256>>>(idx, m, t, out); }, flush, flush_bytes));
    snprintf(nm, 128, "B scalar32 U=16 grid=%dx148", ctas); report(nm, time_ms([&] { k_scalar32<16><<<grid, 256>>>(idx, m, t, out); }, flush, flush_bytes));
    snprintf(nm, 128, "B scalar32 U=32 grid=%dx148", ctas); report(nm, time_ms([&] { k_scalar32<32><<<grid, 256>>>(idx, m, t, out); }, flush, flush_bytes));
    snprintf(nm, 128, "C vec64 U=16 grid=%dx148", ctas); report(nm, time_ms([&] { k_vec64<16><<<grid, 256>>>(idx, m, t, out); }, flush, flush_bytes));
  }
  CK(cudaFuncSetAttribute(k_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (4 * 4096 + 64)));
  CK(cudaFuncSetAttribute(k_bulk<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (2 * 4096 + 64)));
  for (int ctas : {1, 2}) {
    char nm[128];
    snprintf(nm, 128, "F TMA bulk 128B x32/stage, 4 stages, grid=%dx148", ctas);
    report(nm, time_ms([&] { k_bulk<4><<<sms * ctas, 256, 8 * (4 * 4096 + 64)>>>(idx, m, t, out); }, flush, flush_bytes));
    snprintf(nm, 128, "F TMA bulk 128B x32/stage, 2 stages, grid=%dx148", ctas);
    report(nm, time_ms([&] { k_bulk<2><<<sms * ctas, 256, 8 * (2 * 4096 + 64)>>>(idx, m, t, out); }, flush, flush_bytes));
  }
  // G: cp.async gathers into shared memory (lines in flight bounded by shared memory, not registers)
  {
    auto run_g = [&](auto kern, int stages, int ctas) {
      const int smem_bytes = 8 * stages * 4096;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
      char nm[128];
      snprintf(nm, 128, "G cp.async 16B -> smem, %d stages, grid=%dx148 (%d lines in flight / SM)", stages, ctas,
               ctas * 8 * (stages - 1) * 32);
      report(nm, time_ms([&] { kern<<<sms * ctas, 256, smem_bytes>>>(idx, m, t, out); }, flush, flush_bytes));
    };
    run_g(k_cpasync<2>, 2, 3);
    run_g(k_cpasync<2>, 2, 2);
    run_g(k_cpasync<3>, 3, 2);
    run_g(k_cpasync<4>, 4, 1);
    run_g(k_cpasync<6>, 6, 1);
  }
  // table warm in L2 (no flush between), pattern A
  {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    k_vec128<8, false><<<sms * 4, 256>>>(idx, m, t, out);
    CK(cudaEventRecord(a));
    k_vec128<8, false><<<sms * 4, 256>>>(idx, m, t, out);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    report("A vec128 U=8 grid=4x148, L2 warm (idx too)", ms);
  }
  return 0;
}
