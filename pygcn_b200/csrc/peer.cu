// Peer-memory exchange of the dense panels for the row-partitioned multi-GPU layer.
//
// The reference is single-GPU; its `torch.spmm(adj, support)` needs every row of `support`.  With A
// cut into row blocks (one per GPU), rank p needs the panel rows owned by every rank q.  Instead of
// one NCCL all-gather followed by one SpMM, the exchange is done by our own kernel over NVLink peer
// memory, one source slot at a time, so the SpMM over the column block of source q starts as soon as
// slot q has landed:
//
//   * every rank owns a "gathered" buffer [world x pad_rows x F] (cudaMalloc'd here, exported with
//     CUDA IPC, mapped by the peers with peer access enabled) plus two flag words per source rank;
//   * push kernel (communication stream): for k = 1..world-1 copy this rank's slot into the gathered
//     buffer of rank p-k with 16-byte stores to the mapped peer pointer, then publish
//     flag[p] = epoch there (st.release.sys after a system-scope fence; the last CTA to finish the
//     peer publishes).  Before touching a peer's slot it waits for that peer's ack of the previous
//     exchange (the peer is done reading the old contents);
//   * consumer side (compute stream): wait kernel spins (bounded) on flag[q] >= epoch, the SpMM over
//     column block q runs, an ack kernel writes ack[p] = epoch into rank q's ack words.
// All state (epochs, counters) lives in device memory, so the whole step can be CUDA-graph replayed.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace gcnb {
namespace {

constexpr int kPushThreads = 256;
// Polls before a waiting kernel gives up with __trap (a protocol bug must not hang the GPU): 2^26 polls of >= 64 ns are
// a minute or more.  The exchange needs the ranks in lock step -- a rank that stalls longer than this on the host (a
// checkpoint, a data-loader hiccup) kills its peers' contexts -- so DistGraphConvolution(exchange="auto") never picks it
// (dist.py); GCNB_PEER_SPIN_LIMIT=<polls> changes the limit, 0 waits for ever.
__device__ unsigned long long d_peer_spin_limit = 1ull << 26;

int peer_spin_limit_init() {
  static bool done[64] = {};
  int dev = 0;
  GCNB_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || done[dev]) return GCNB_OK;
  const char* e = getenv("GCNB_PEER_SPIN_LIMIT");
  if (e) {
    unsigned long long v = strtoull(e, nullptr, 10);
    GCNB_CUDA(cudaMemcpyToSymbol(d_peer_spin_limit, &v, sizeof(v)));
  }
  done[dev] = true;
  return GCNB_OK;
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void spin_until_ge(const uint32_t* p, uint32_t want) {
  unsigned long long spins = 0;
  const unsigned long long limit = d_peer_spin_limit;
  while ((int32_t)(ld_acquire_sys(p) - want) < 0) {
    __nanosleep(64);
    if (limit != 0 && ++spins > limit) __trap();
  }
}

struct PushArgs {
  float4* dst[GCNB_MAX_PEERS];         // peer k: mapped address of OUR slot in that peer's gathered buffer
  uint32_t* flag[GCNB_MAX_PEERS];      // peer k: mapped address of flag[our rank] in that peer's flag words
  const uint32_t* ack[GCNB_MAX_PEERS]; // peer k: LOCAL address of ack[peer k's rank] (written by that peer)
};

// epoch[0] += 1 (one thread); also clears the per-peer arrival counters of the push kernel
__global__ void epoch_bump_kernel(uint32_t* epoch, uint32_t* counters, int n) {
  if (threadIdx.x == 0) epoch[0] += 1;
  if (threadIdx.x < n) counters[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(kPushThreads)
peer_push_kernel(const float4* __restrict__ src, size_t n16, int n_peers, PushArgs a, const uint32_t* __restrict__ epoch,
                 uint32_t* __restrict__ counters) {
  const uint32_t ep = epoch[0];
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int k = 0; k < n_peers; ++k) {
    // the peer has finished reading what the previous exchange put into this slot
    if (threadIdx.x == 0) spin_until_ge(a.ack[k], ep - 1);
    __syncthreads();
    float4* __restrict__ dst = a.dst[k];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += 4 * stride) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * stride < n16) v[u] = src[i + u * stride];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i + u * stride < n16) dst[i + u * stride] = v[u];
    }
    __threadfence_system();  // this thread's peer stores are ordered before what follows
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t done = atomicAdd(counters + k, 1u) + 1u;
      if (done == gridDim.x) {  // last CTA for this peer: every CTA's stores are fenced -> publish
        __threadfence_system();
        st_release_sys(a.flag[k], ep);
      }
    }
  }
}

__global__ void wait_flag_kernel(const uint32_t* flag, const uint32_t* epoch, uint32_t lag) {
  if (threadIdx.x == 0) spin_until_ge(flag, epoch[0] - lag);
}

__global__ void ack_kernel(uint32_t* peer_ack, const uint32_t* epoch) {
  if (threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys(peer_ack, epoch[0]);
  }
}

}  // namespace
}  // namespace gcnb

using namespace gcnb;

extern "C" int gcnb_symm_alloc(size_t bytes, void** d_ptr, void* handle_out) {
  GCNB_REQUIRE(d_ptr != nullptr && handle_out != nullptr && bytes > 0, "symm_alloc: bad argument");
  void* p = nullptr;
  GCNB_CUDA(cudaMalloc(&p, bytes));
  GCNB_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("symm_alloc: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return GCNB_E_CUDA;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == GCNB_IPC_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  *d_ptr = p;
  return GCNB_OK;
}

extern "C" int gcnb_symm_open(const void* handle, void** d_ptr) {
  GCNB_REQUIRE(handle != nullptr && d_ptr != nullptr, "symm_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  GCNB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *d_ptr = p;
  return GCNB_OK;
}

extern "C" int gcnb_symm_close(void* d_ptr) {
  if (d_ptr) GCNB_CUDA(cudaIpcCloseMemHandle(d_ptr));
  return GCNB_OK;
}

extern "C" int gcnb_symm_free(void* d_ptr) {
  if (d_ptr) GCNB_CUDA(cudaFree(d_ptr));
  return GCNB_OK;
}

extern "C" int gcnb_peer_epoch_bump(uint32_t* d_epoch, uint32_t* d_counters, int n_peers, void* stream) {
  GCNB_REQUIRE(d_epoch != nullptr && d_counters != nullptr && n_peers >= 0 && n_peers <= GCNB_MAX_PEERS,
               "peer_epoch_bump: bad argument");
  epoch_bump_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_epoch, d_counters, n_peers);
  GCNB_LAUNCH_CHECK();
  return GCNB_OK;
}

extern "C" int gcnb_peer_push(const void* d_src, size_t bytes, int n_peers, void* const* peer_dst,
                              uint32_t* const* peer_flag, const uint32_t* const* local_ack, const uint32_t* d_epoch,
                              uint32_t* d_counters, int n_ctas, void* stream) {
  GCNB_TRY(peer_spin_limit_init());
  GCNB_REQUIRE(n_peers >= 0 && n_peers <= GCNB_MAX_PEERS, "peer_push: at most %d peers", GCNB_MAX_PEERS);
  if (n_peers == 0 || bytes == 0) return GCNB_OK;
  GCNB_REQUIRE(d_src && peer_dst && peer_flag && local_ack && d_epoch && d_counters, "peer_push: null argument");
  GCNB_REQUIRE(bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(d_src) & 15u) == 0, "peer_push: 16-byte granularity");
  PushArgs a;
  for (int k = 0; k < n_peers; ++k) {
    GCNB_REQUIRE(peer_dst[k] && peer_flag[k] && local_ack[k] && (reinterpret_cast<uintptr_t>(peer_dst[k]) & 15u) == 0,
                 "peer_push: bad peer pointer %d", k);
    a.dst[k] = reinterpret_cast<float4*>(peer_dst[k]);
    a.flag[k] = peer_flag[k];
    a.ack[k] = local_ack[k];
  }
  if (n_ctas < 1) n_ctas = 32;
  if (n_ctas > 4 * kNumSMs) n_ctas = 4 * kNumSMs;
  peer_push_kernel<<<n_ctas, kPushThreads, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(d_src), bytes / 16,
                                                                     n_peers, a, d_epoch, d_counters);
  GCNB_LAUNCH_CHECK();
  return GCNB_OK;
}

extern "C" int gcnb_peer_wait(const uint32_t* d_flag, const uint32_t* d_epoch, void* stream) {
  GCNB_TRY(peer_spin_limit_init());
  GCNB_REQUIRE(d_flag && d_epoch, "peer_wait: null argument");
  wait_flag_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_flag, d_epoch, 0u);
  GCNB_LAUNCH_CHECK();
  return GCNB_OK;
}

extern "C" int gcnb_peer_wait_lag(const uint32_t* d_word, const uint32_t* d_epoch, uint32_t lag, void* stream) {
  GCNB_TRY(peer_spin_limit_init());
  GCNB_REQUIRE(d_word && d_epoch, "peer_wait_lag: null argument");
  wait_flag_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_word, d_epoch, lag);
  GCNB_LAUNCH_CHECK();
  return GCNB_OK;
}

extern "C" int gcnb_peer_copy(void* dst, const void* src, size_t bytes, void* stream) {
  GCNB_REQUIRE(dst && src, "peer_copy: null argument");
  if (bytes) GCNB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return GCNB_OK;
}

extern "C" int gcnb_peer_ack(uint32_t* peer_ack, const uint32_t* d_epoch, void* stream) {
  GCNB_REQUIRE(peer_ack && d_epoch, "peer_ack: null argument");
  ack_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(peer_ack, d_epoch);
  GCNB_LAUNCH_CHECK();
  return GCNB_OK;
}
