// C ABI of the row-partitioned layer's exchange step (nothing in the reference, which is single-GPU):
// the needed-rows-only ("halo") exchange of a dense panel over an NCCL communicator the CALLER owns.
//
//   out_p = A[p, :] . P   needs, on rank p, the rows of the panel P that appear as a column in A's row block p.
//   On a power-law graph that is a third of the remote rows (tools/halo_fraction.py), so instead of an all-gather
//   every rank sends each peer exactly the rows that peer reads:
//     pack     one kernel gathers, for all destinations at once, the rows `send_rows` of this rank's panel into a
//              contiguous send buffer (one warp per row, 16-byte copies);
//     exchange one grouped round of ncclSend / ncclRecv on the caller's stream: the rows from source q land at
//              row recv_offset[q] of the compact panel the SpMM over the renumbered block gathers from.
//   The communicator is passed in as an opaque ncclComm_t (from torch: ProcessGroupNCCL._comm_ptr(), or the
//   application's own); the library never creates, splits or destroys one.  NCCL is bound at run time (dlopen of the
//   libnccl.so.2 already in the process), so libgcnb200.so has no link-time dependency on it.
#include <dlfcn.h>
#include <stdlib.h>

#include <new>
#include <vector>

#include "common.cuh"

namespace gcnb {
namespace {

// the slice of nccl.h this file uses (ABI-stable since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
typedef int ncclResult_t;
constexpr int kNcclFloat32 = 7;  // ncclFloat32 in ncclDataType_t
struct NcclApi {
  ncclResult_t (*group_start)();
  ncclResult_t (*group_end)();
  ncclResult_t (*send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
  const char* (*error_string)(ncclResult_t);
  bool ok = false;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.ok ? &api : nullptr;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy torch already loaded, if any
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return nullptr;
  api.group_start = reinterpret_cast<decltype(api.group_start)>(dlsym(h, "ncclGroupStart"));
  api.group_end = reinterpret_cast<decltype(api.group_end)>(dlsym(h, "ncclGroupEnd"));
  api.send = reinterpret_cast<decltype(api.send)>(dlsym(h, "ncclSend"));
  api.recv = reinterpret_cast<decltype(api.recv)>(dlsym(h, "ncclRecv"));
  api.error_string = reinterpret_cast<decltype(api.error_string)>(dlsym(h, "ncclGetErrorString"));
  api.ok = api.group_start && api.group_end && api.send && api.recv && api.error_string;
  return api.ok ? &api : nullptr;
}

#define GCNB_NCCL(api, expr)                                                                          \
  do {                                                                                                \
    ncclResult_t _r = (expr);                                                                         \
    if (_r != 0) {                                                                                    \
      ::gcnb::set_error("%s failed: %s (%s:%d)", #expr, (api)->error_string(_r), __FILE__, __LINE__); \
      return GCNB_E_CUDA;                                                                             \
    }                                                                                                 \
  } while (0)

// dst[i, 0:f] = src[rows[i], 0:f]: one warp per row, 16-byte copies when both sides allow
__global__ void __launch_bounds__(256)
halo_pack_kernel(int64_t n, const int32_t* __restrict__ rows, const float* __restrict__ src, int64_t lds, int f,
                 float* __restrict__ dst, int vec) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n) return;
  const float* s = src + (int64_t)__ldg(rows + i) * lds;
  float* d = dst + i * (int64_t)f;
  if (vec) {
    for (int q = lane; q < f / 4; q += 32)
      __stcs(reinterpret_cast<float4*>(d) + q, __ldg(reinterpret_cast<const float4*>(s) + q));
  } else {
    for (int j = lane; j < f; j += 32) d[j] = __ldg(s + j);
  }
}

}  // namespace
}  // namespace gcnb

struct gcnb_halo {
  int rank = 0, world = 0;
  std::vector<int64_t> send_count, send_offset, recv_count, recv_offset;  // rows, per peer rank
  int64_t total_send = 0, total_recv = 0;
  int32_t* d_send_rows = nullptr;  // [total_send] local row ids, grouped by destination rank (owned copy)
};

using namespace gcnb;

extern "C" int gcnb_halo_create(int rank, int world, const int64_t* h_send_counts, const int32_t* d_send_rows,
                                const int64_t* h_recv_counts, const int64_t* h_recv_offsets, void* stream,
                                gcnb_halo** out) {
  GCNB_REQUIRE(out != nullptr, "halo_create: out is null");
  *out = nullptr;
  GCNB_REQUIRE(world >= 1 && rank >= 0 && rank < world && world <= 1024, "halo_create: bad rank / world (%d / %d)", rank, world);
  GCNB_REQUIRE(h_send_counts && h_recv_counts && h_recv_offsets, "halo_create: null count arrays");
  gcnb_halo* h = new (std::nothrow) gcnb_halo();
  GCNB_REQUIRE(h != nullptr, "halo_create: host allocation failed");
  h->rank = rank;
  h->world = world;
  h->send_count.assign(h_send_counts, h_send_counts + world);
  h->recv_count.assign(h_recv_counts, h_recv_counts + world);
  h->recv_offset.assign(h_recv_offsets, h_recv_offsets + world);
  h->send_offset.resize(world);
  for (int r = 0; r < world; ++r) {
    if (h->send_count[r] < 0 || h->recv_count[r] < 0 || (r == rank && (h->send_count[r] != 0 || h->recv_count[r] != 0))) {
      delete h;
      GCNB_REQUIRE(false, "halo_create: counts must be >= 0 and 0 for the rank itself");
    }
    h->send_offset[r] = h->total_send;
    h->total_send += h->send_count[r];
    h->total_recv += h->recv_count[r];
  }
  if (h->total_send > 0) {
    if (d_send_rows == nullptr) {
      delete h;
      GCNB_REQUIRE(false, "halo_create: send rows are null");
    }
    cudaError_t e = cudaMalloc(&h->d_send_rows, (size_t)h->total_send * sizeof(int32_t));
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(h->d_send_rows, d_send_rows, (size_t)h->total_send * sizeof(int32_t), cudaMemcpyDeviceToDevice,
                          (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) {
      if (h->d_send_rows) cudaFree(h->d_send_rows);
      delete h;
      set_error("halo_create: %s", cudaGetErrorString(e));
      return GCNB_E_CUDA;
    }
  }
  *out = h;
  return GCNB_OK;
}

extern "C" void gcnb_halo_free(gcnb_halo* h) {
  if (!h) return;
  if (h->d_send_rows) cudaFree(h->d_send_rows);
  delete h;
}

extern "C" int64_t gcnb_halo_send_rows(const gcnb_halo* h) { return h ? h->total_send : 0; }
extern "C" int64_t gcnb_halo_recv_rows(const gcnb_halo* h) { return h ? h->total_recv : 0; }

extern "C" int gcnb_halo_nccl_available(void) { return nccl_api() != nullptr ? 1 : 0; }

extern "C" int gcnb_halo_pack(const gcnb_halo* h, const float* d_panel, int64_t ldp, int64_t f, float* d_sendbuf,
                              void* stream) {
  GCNB_REQUIRE(h != nullptr, "halo_pack: null plan");
  GCNB_REQUIRE(f > 0 && ldp >= f && f < (1ll << 24), "halo_pack: bad width / leading dimension");
  if (h->total_send == 0) return GCNB_OK;
  GCNB_REQUIRE(d_panel != nullptr && d_sendbuf != nullptr, "halo_pack: null buffer");
  const int vec = (f % 4 == 0) && (ldp % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_panel) & 15u) == 0) &&
                  ((reinterpret_cast<uintptr_t>(d_sendbuf) & 15u) == 0);
  halo_pack_kernel<<<(unsigned)ceil_div(h->total_send, 8), 256, 0, (cudaStream_t)stream>>>(
      h->total_send, h->d_send_rows, d_panel, ldp, (int)f, d_sendbuf, vec);
  GCNB_LAUNCH_CHECK();
  return GCNB_OK;
}

extern "C" int gcnb_halo_exchange(const gcnb_halo* h, void* nccl_comm, const float* d_sendbuf, int64_t f,
                                  float* d_compact, void* stream) {
  GCNB_REQUIRE(h != nullptr && nccl_comm != nullptr, "halo_exchange: null plan / communicator");
  GCNB_REQUIRE(f > 0, "halo_exchange: bad width");
  NcclApi* api = nccl_api();
  GCNB_REQUIRE(api != nullptr, "halo_exchange: libnccl.so.2 not found in the process (dlopen)");
  if (h->total_send == 0 && h->total_recv == 0) return GCNB_OK;
  GCNB_REQUIRE((h->total_send == 0 || d_sendbuf != nullptr) && (h->total_recv == 0 || d_compact != nullptr),
               "halo_exchange: null buffer");
  ncclComm_t comm = reinterpret_cast<ncclComm_t>(nccl_comm);
  cudaStream_t st = (cudaStream_t)stream;
  GCNB_NCCL(api, api->group_start());
  for (int k = 1; k < h->world; ++k) {
    const int to = (h->rank + k) % h->world, from = (h->rank - k + h->world) % h->world;
    if (h->send_count[to] > 0)
      GCNB_NCCL(api, api->send(d_sendbuf + h->send_offset[to] * f, (size_t)(h->send_count[to] * f), kNcclFloat32, to, comm, st));
    if (h->recv_count[from] > 0)
      GCNB_NCCL(api, api->recv(d_compact + h->recv_offset[from] * f, (size_t)(h->recv_count[from] * f), kNcclFloat32, from,
                               comm, st));
  }
  GCNB_NCCL(api, api->group_end());
  return GCNB_OK;
}
