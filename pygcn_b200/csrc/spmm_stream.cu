// Streaming CSR SpMM for wide panel rows and power-law graphs:  out = A * B (+ bias) (ReLU)   -- pygcn/layers.py:34-36
// (and dS = A^T * G on the CSR of A^T).
//
// Why a second kernel.  On the products-shaped R-MAT graph (BASELINE configs[3]; median row 3 entries, 43 % of the
// entries in rows of more than 1024) the warp-per-row kernel is latency-bound: 48 registers of gathered data per
// thread cap it at ~110 KB in flight per SM, every short row pays a rowptr -> (col, val) -> gather chain, and the hub
// rows need two more launches (profiles/r02_rmat_probe_before.txt: 7.1 ms at F = 256, DRAM 60 %, L2 41 % busy,
// 43 % of the warp slots filled).  Here
//   * the work is cut by STORED ENTRIES, not rows ("merge-style"): item i = entries [1024 i, 1024 (i + 1)) of the
//     pair stream, whatever rows they belong to; warps take items round-robin, so the load is balanced by
//     construction and the (col, val) stream is read sequentially, one coalesced 8-byte load per lane and batch;
//   * every gathered panel row is ONE 1-D TMA copy (cp.async.bulk global -> shared, completion on an mbarrier):
//     bytes in flight are bounded by shared memory (~190 KB per SM), not by registers, and the copy carries an L2
//     eviction hint: rows of the most referenced columns (heat class tag of the pair, common.cuh) are kept
//     (evict_last), the rest streams through (evict_first);
//   * a batch of B rows is consumed out of shared memory with conflict-free LDS.128; the end-of-row tag of the pair
//     says when the accumulator is complete: rows inside one item are stored directly (bias / ReLU / dropout
//     epilogue), the two rows an item may share with its neighbours go to scratch as partial sums and a small
//     fix-up kernel adds them in item order (no atomics, run-to-run deterministic).
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "ptx.cuh"

namespace gcnb {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kStreamWarps = 4;  // warps per CTA; CTAs per SM follow from the shared-memory ring

__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst_smem, uint64_t src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_plain(uint32_t dst_smem, uint64_t src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint2 ldg_pair_stream(const uint2* p, uint64_t pol) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 q;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "r"(addr));
  return q;
}

__device__ __forceinline__ void fma4(float4& acc, float v, float a, float b, float c, float d) {
  acc.x = fmaf(v, a, acc.x);
  acc.y = fmaf(v, b, acc.y);
  acc.z = fmaf(v, c, acc.z);
  acc.w = fmaf(v, d, acc.w);
}
// acc[0..A) += v * (one 16-byte chunk of a panel row): 4 fp32, or 8 bf16 (-> fp32 by a shift / a mask)
template <bool BF16>
__device__ __forceinline__ void fma_chunk(float4* acc, float v, const uint4& x) {
  if constexpr (!BF16) {
    fma4(acc[0], v, __uint_as_float(x.x), __uint_as_float(x.y), __uint_as_float(x.z), __uint_as_float(x.w));
  } else {
    fma4(acc[0], v, __uint_as_float(x.x << 16), __uint_as_float(x.x & 0xffff0000u), __uint_as_float(x.y << 16),
         __uint_as_float(x.y & 0xffff0000u));
    fma4(acc[1], v, __uint_as_float(x.z << 16), __uint_as_float(x.z & 0xffff0000u), __uint_as_float(x.w << 16),
         __uint_as_float(x.w & 0xffff0000u));
  }
}

// out[row, 4q .. 4q+3] = epilogue(a)   (the same epilogue as spmm.cu: accumulate, bias, ReLU, dropout mask)
__device__ __forceinline__ void store_out4(float* out_row, int64_t row, int q, int f, bool vec_out, float4 a, const Epilogue& ep) {
  float r[4] = {a.x, a.y, a.z, a.w};
  if (ep.accumulate) {
    if (vec_out) {
      const float4 o = *reinterpret_cast<const float4*>(out_row + 4 * q);
      r[0] += o.x; r[1] += o.y; r[2] += o.z; r[3] += o.w;
    } else {
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (4 * q + t < f) r[t] += out_row[4 * q + t];
    }
  }
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int c = 4 * q + t;
    if (c < f) {
      if (ep.bias) r[t] += __ldg(ep.bias + c);
      if (ep.relu) r[t] = fmaxf(r[t], 0.f);
      if (ep.mask) r[t] = __ldg(ep.mask + row * ep.ld_mask + c) ? r[t] * ep.mask_scale : 0.f;
    }
  }
  if (vec_out) {
    __stcs(reinterpret_cast<float4*>(out_row + 4 * q), make_float4(r[0], r[1], r[2], r[3]));  // written once, read by the next kernel
  } else {
#pragma unroll
    for (int t = 0; t < 4; ++t)
      if (4 * q + t < f) out_row[4 * q + t] = r[t];
  }
}

// CH  = 16-byte chunks of a panel row per lane (chunk lane + 32 k), BF16 = panel element type,
// B   = rows per batch (one TMA copy each, issued by lanes 0..B-1), NB = batches in the per-warp ring.
template <int CH, bool BF16, int B, int NB>
__global__ void __launch_bounds__(kStreamWarps * 32)
spmm_stream_kernel(int nnz, int n_items, const uint32_t* __restrict__ items, const uint2* __restrict__ pair,
                   const void* __restrict__ b, uint32_t ldb_bytes, uint32_t row_bytes, int f, Epilogue ep,
                   float* __restrict__ out, int64_t ldo, int vec_out, float* __restrict__ partial, int ldp,
                   int hot_class_max, int hint_mode) {
  constexpr int A = BF16 ? 2 : 1;            // float4 accumulators per chunk
  constexpr int kElems = BF16 ? 8 : 4;       // panel elements per chunk
  constexpr int BPI = kStreamItem / B;       // batches per item
  static_assert(kStreamItem % B == 0 && B <= 32, "a batch is at most one entry per lane");
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t stage_bytes = (uint32_t)B * row_bytes;
  const uint32_t ring = smem_u32(smem) + (uint32_t)warp * (uint32_t)NB * stage_bytes;
  const uint32_t bars = smem_u32(smem) + (uint32_t)kStreamWarps * (uint32_t)NB * stage_bytes + (uint32_t)warp * NB * 8u;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < NB; ++s) mbar_init(bars + 8u * s, 1);
    fence_barrier_init();
  }
  __syncthreads();

  const uint64_t pol_hot = policy_evict_last();
  const uint64_t pol_cold = policy_evict_first();
  const uint64_t bbase = reinterpret_cast<uint64_t>(b);
  const int nch = (f + kElems - 1) / kElems;  // 16-byte chunks per panel row
  uint32_t qoff[CH];                          // byte offset of this lane's chunks inside a staged row (clamped: valid smem)
#pragma unroll
  for (int k = 0; k < CH; ++k) qoff[k] = 16u * (uint32_t)min(lane + 32 * k, nch - 1);

  const int gw = blockIdx.x * kStreamWarps + warp;
  const int GW = gridDim.x * kStreamWarps;
  const int my_items = gw < n_items ? (n_items - gw + GW - 1) / GW : 0;
  const int nb = my_items * BPI;  // batches of this warp (the last item's tail batches may be empty)

  // batch k of this warp -> first stored entry and number of entries
  auto batch_e0 = [&](int k) { return (gw + (k / BPI) * GW) * kStreamItem + (k % BPI) * B; };
  auto batch_cnt = [&](int k) { return k < nb ? max(0, min(B, nnz - batch_e0(k))) : 0; };
  auto load_pair = [&](int k) {
    uint2 p = make_uint2(0u, 0u);
    if (k < nb) {
      const int e = batch_e0(k) + lane;
      if (lane < B && e < nnz) p = ldg_pair_stream(pair + e, pol_cold);
    }
    return p;
  };
  auto issue = [&](int s, const uint2& p, int cnt) {
    if (cnt <= 0) return;
    const uint32_t bar = bars + 8u * s;
    if (lane == 0) mbar_expect_tx(bar, (uint32_t)cnt * row_bytes);
    __syncwarp();
    if (lane < cnt) {
      const uint32_t c = p.x & kPairColMask;
      const uint64_t src = bbase + (uint64_t)c * ldb_bytes;
      const uint32_t dst = ring + (uint32_t)s * stage_bytes + (uint32_t)lane * row_bytes;
      if (hint_mode == 0) {
        bulk_g2s_plain(dst, src, row_bytes, bar);
      } else {
        const bool hot = (int)((p.x >> kPairClassShift) & 15u) <= hot_class_max;
        if (hot) bulk_g2s_hint(dst, src, row_bytes, bar, pol_hot);
        else if (hint_mode == 2) bulk_g2s_hint(dst, src, row_bytes, bar, pol_cold);
        else bulk_g2s_plain(dst, src, row_bytes, bar);
      }
    }
  };

  float4 acc[CH * A];
#pragma unroll
  for (int k = 0; k < CH * A; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto zero_acc = [&]() {
#pragma unroll
    for (int k = 0; k < CH * A; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  // complete row -> out (epilogue); partial sums of a row shared with a neighbouring item -> scratch, raw
  auto store_row = [&](int row) {
    float* out_row = out + (int64_t)row * ldo;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int q = lane + 32 * k;
      if (q < nch) {
#pragma unroll
        for (int t = 0; t < A; ++t)
          if (4 * (q * A + t) < f) store_out4(out_row, row, q * A + t, f, vec_out != 0, acc[k * A + t], ep);
      }
    }
  };
  auto store_partial = [&](int slot) {
    float* prow = partial + (int64_t)slot * ldp;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      const int q = lane + 32 * k;
      if (q < nch) {
#pragma unroll
        for (int t = 0; t < A; ++t)
          if (4 * (q * A + t) < f) *reinterpret_cast<float4*>(prow + 4 * (q * A + t)) = acc[k * A + t];
      }
    }
  };

  // ---- prologue: fill the ring
  uint2 pr[NB];  // pairs of the batch occupying stage s (the values and row-end tags are needed when it is consumed)
#pragma unroll
  for (int s = 0; s < NB; ++s) pr[s] = load_pair(s);
#pragma unroll
  for (int s = 0; s < NB; ++s) issue(s, pr[s], batch_cnt(s));
  uint2 pnext = load_pair(NB);
  uint32_t phase = 0u;  // bit s = parity the next wait on stage s expects
  int row = 0;
  bool head = false;    // the current row began in an earlier item: its sum goes to partial[2 * item]
  bool open = false;    // acc holds entries of a row whose end has not been seen
  uint32_t info_next = my_items > 0 ? __ldg(items + gw) : 0u;

  for (int k0 = 0; k0 < nb; k0 += NB) {
#pragma unroll
    for (int s = 0; s < NB; ++s) {
      const int k = k0 + s;
      const int cnt = batch_cnt(k);
      const int item = gw + (k / BPI) * GW;
      if (k < nb && k % BPI == 0) {  // first batch of an item
        row = (int)(info_next & 0x7fffffffu);
        head = (info_next >> 31) != 0u;
        open = false;
        zero_acc();
        const int nxt = item + GW;
        if (nxt < n_items) info_next = __ldg(items + nxt);
      }
      if (cnt > 0) {
        mbar_wait(bars + 8u * s, (phase >> s) & 1u);
        phase ^= 1u << s;
        const uint32_t sbase = ring + (uint32_t)s * stage_bytes;
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
          const float v = __uint_as_float(__shfl_sync(kFull, pr[s].y, j));
          const uint32_t tag = __shfl_sync(kFull, pr[s].x, j);
          uint4 x[CH];
#pragma unroll
          for (int c = 0; c < CH; ++c) x[c] = lds128(sbase + (uint32_t)j * row_bytes + qoff[c]);
#pragma unroll
          for (int c = 0; c < CH; ++c) fma_chunk<BF16>(&acc[c * A], v, x[c]);
          open = true;
          if (tag & kPairRowEnd) {  // (warp-uniform)
            if (head) store_partial(2 * item); else store_row(row);
            head = false;
            open = false;
            ++row;
            zero_acc();
          }
        }
        // last batch of the item: a row that continues in the next item leaves its partial sum
        if ((k % BPI == BPI - 1 || batch_e0(k) + cnt >= nnz) && open) {
          store_partial(2 * item + (head ? 0 : 1));
          open = false;
        }
      }
      __syncwarp();  // every lane is done reading stage s before the copy engine overwrites it
      pr[s] = pnext;
      issue(s, pr[s], batch_cnt(k + NB));
      pnext = load_pair(k + NB + 1);
    }
  }
}

// Rows shared by several items: out[r] = epilogue(tail partial of the item the row starts in + head partials of the
// items it continues through, in item order).  One warp per item boundary; the warp at the row's FIRST boundary does it.
__global__ void __launch_bounds__(256)
spmm_stream_fixup_kernel(int n_items, const uint32_t* __restrict__ items, const int32_t* __restrict__ rowptr,
                         const float* __restrict__ partial, int ldp, int f, Epilogue ep, float* __restrict__ out, int64_t ldo) {
  const int lane = threadIdx.x & 31;
  const int bnd = blockIdx.x * 8 + (threadIdx.x >> 5) + 1;  // boundary between items bnd - 1 and bnd
  if (bnd >= n_items) return;
  const uint32_t info = __ldg(items + bnd);
  if (!(info >> 31)) return;  // item bnd starts with a fresh row
  const int row = (int)(info & 0x7fffffffu);
  const int start = __ldg(rowptr + row);
  if (start < (bnd - 1) * kStreamItem) return;  // the row already crossed an earlier boundary
  const int last = (__ldg(rowptr + row + 1) - 1) / kStreamItem;
  for (int j = lane; j < f; j += 32) {
    float acc = partial[(int64_t)(2 * (bnd - 1) + 1) * ldp + j];
    for (int i = bnd; i <= last; ++i) acc += partial[(int64_t)(2 * i) * ldp + j];
    if (ep.accumulate) acc += out[(int64_t)row * ldo + j];
    if (ep.bias) acc += __ldg(ep.bias + j);
    if (ep.relu) acc = fmaxf(acc, 0.f);
    if (ep.mask) acc = __ldg(ep.mask + (int64_t)row * ep.ld_mask + j) ? acc * ep.mask_scale : 0.f;
    out[(int64_t)row * ldo + j] = acc;
  }
}

// tuning (environment at first use, gcnb_set_tuning afterwards)
int g_stream_mode = -1;      // 0 off, 1 auto, 2 forced wherever eligible
int g_stream_hot_mb = -1;    // L2 budget for the rows of the hot classes (MB)
int g_stream_hint = -1;      // 0 no hints, 1 hot rows evict_last, 2 + cold rows evict_first
int g_stream_min_row = -1;   // auto: smallest panel row (bytes) that takes this kernel
int g_stream_batch = -1;     // 0 auto, else rows per batch (8 / 16 / 32)

void stream_init() {
  if (g_stream_mode >= 0) return;
  const char* e = getenv("GCNB_SPMM_STREAM");
  g_stream_mode = e ? atoi(e) : 1;
  e = getenv("GCNB_L2_HOT_MB");
  g_stream_hot_mb = e ? atoi(e) : 48;
  e = getenv("GCNB_STREAM_HINT");
  g_stream_hint = e ? atoi(e) : 2;
  e = getenv("GCNB_STREAM_MIN_ROW_BYTES");
  g_stream_min_row = e ? atoi(e) : 256;
  e = getenv("GCNB_STREAM_BATCH");
  g_stream_batch = e ? atoi(e) : 0;
}

template <int CH, bool BF16, int B, int NB>
int launch_stream_t(const CsrView& a, const void* b, int64_t ldb_bytes, uint32_t row_bytes, int f, const Epilogue& ep,
                    float* out, int64_t ldo, bool vec_out, float* partial, int ldp, int hot_class_max, cudaStream_t st) {
  auto kernel = spmm_stream_kernel<CH, BF16, B, NB>;
  const size_t smem = (size_t)kStreamWarps * NB * B * row_bytes + (size_t)kStreamWarps * NB * 8 + 128;
  int dev = 0;
  GCNB_CUDA(cudaGetDevice(&dev));
  static bool configured[64] = {};  // per device: the opt-in is a per-device attribute of the function
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    GCNB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  GCNB_REQUIRE(smem <= 227 * 1024, "spmm(stream): ring of %zu bytes does not fit shared memory", smem);
  int per_sm = (int)((size_t)(200 * 1024) / smem);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int grid = kNumSMs * per_sm;
  const int want = (int)ceil_div(a.n_stream_items, kStreamWarps);
  if (grid > want) grid = want;
  kernel<<<grid, kStreamWarps * 32, smem, st>>>((int)a.nnz, (int)a.n_stream_items, a.stream_items, a.pair, b,
                                                (uint32_t)ldb_bytes, row_bytes, f, ep, out, ldo, vec_out ? 1 : 0, partial,
                                                ldp, hot_class_max, g_stream_hint);
  GCNB_LAUNCH_CHECK();
  if (a.n_stream_items > 1) {
    spmm_stream_fixup_kernel<<<(unsigned)ceil_div(a.n_stream_items - 1, 8), 256, 0, st>>>(
        (int)a.n_stream_items, a.stream_items, a.rowptr, partial, ldp, f, ep, out, ldo);
    GCNB_LAUNCH_CHECK();
  }
  return GCNB_OK;
}

inline int stream_ldp(int64_t f) { return (int)(ceil_div(f, 4) * 4); }

}  // namespace

int spmm_stream_mode() {
  stream_init();
  return g_stream_mode;
}

void spmm_stream_set(int key, int value) {
  stream_init();
  if (key == GCNB_TUNE_SPMM_STREAM) g_stream_mode = value;
  else if (key == GCNB_TUNE_STREAM_HOT_MB) g_stream_hot_mb = value;
  else if (key == GCNB_TUNE_STREAM_HINT) g_stream_hint = value;
  else if (key == GCNB_TUNE_STREAM_MIN_ROW_BYTES) g_stream_min_row = value;
  else if (key == GCNB_TUNE_STREAM_BATCH) g_stream_batch = value;
}

// nch = 16-byte chunks per panel row.  The view needs tagged pairs and items, and no empty row (an empty row has no
// entry to carry the end-of-row tag; graphs with self loops -- every GCN adjacency -- have none).
bool spmm_stream_eligible(const CsrView& a, int64_t row_bytes, int nch) {
  stream_init();
  if (g_stream_mode == 0) return false;
  if (!a.pair_tagged || a.stream_items == nullptr || a.n_stream_items == 0 || a.bin_rows[0] != 0) return false;
  if (nch > 64 || row_bytes % 16 != 0 || row_bytes > 1024) return false;
  if (g_stream_mode == 2) return true;
  // auto: rows wide enough that one copy per row pays, and enough items to fill the machine
  return row_bytes >= g_stream_min_row && a.n_stream_items >= 2 * kNumSMs * kStreamWarps;
}

size_t spmm_stream_workspace_bytes(const CsrView& a, int64_t f) {
  if (!a.pair_tagged || a.n_stream_items == 0) return 0;
  return (size_t)a.n_stream_items * 2 * (size_t)stream_ldp(f) * sizeof(float);
}

int spmm_stream_launch(const CsrView& a, const void* b, int64_t ldb_bytes, int f, bool bf16, const Epilogue& ep,
                       float* out, int64_t ldo, bool vec_out, void* ws, size_t ws_bytes, cudaStream_t st) {
  stream_init();
  const int kE = bf16 ? 8 : 4;
  const int nch = (int)ceil_div(f, kE);
  const uint32_t row_bytes = 16u * (uint32_t)nch;
  GCNB_REQUIRE(ws != nullptr && ws_bytes >= spmm_stream_workspace_bytes(a, f), "spmm(stream): workspace too small");
  GCNB_REQUIRE(ldb_bytes < (1ll << 32) && ldb_bytes % 16 == 0, "spmm(stream): panel row stride must be a multiple of 16 bytes");
  float* partial = reinterpret_cast<float*>(ws);
  const int ldp = stream_ldp(f);
  // hottest classes whose rows fit the L2 budget: class c holds at most 1024 * 2^c columns
  int hot = -1;
  for (int c = 0; c < 15; ++c)
    if ((1024ll << c) * (int64_t)row_bytes <= (int64_t)g_stream_hot_mb * (1ll << 20)) hot = c;
  // rows per batch: ~24-32 KB of ring per warp (3 batches), so that 6-8 warps share an SM's shared memory
  int batch = g_stream_batch;
  if (batch != 8 && batch != 16 && batch != 32) batch = row_bytes > 512 ? 8 : (row_bytes > 256 ? 16 : 32);
#define GCNB_STREAM_CASE(CH_, BF_, B_) \
  return launch_stream_t<CH_, BF_, B_, 3>(a, b, ldb_bytes, row_bytes, f, ep, out, ldo, vec_out, partial, ldp, hot, st)
#define GCNB_STREAM_BATCHES(CH_, BF_)        \
  do {                                       \
    if (batch == 8) GCNB_STREAM_CASE(CH_, BF_, 8);   \
    if (batch == 16) GCNB_STREAM_CASE(CH_, BF_, 16); \
    GCNB_STREAM_CASE(CH_, BF_, 32);          \
  } while (0)
  if (bf16) {
    if (nch <= 32) GCNB_STREAM_BATCHES(1, true);
    GCNB_STREAM_BATCHES(2, true);
  } else {
    if (nch <= 32) GCNB_STREAM_BATCHES(1, false);
    GCNB_STREAM_BATCHES(2, false);
  }
#undef GCNB_STREAM_BATCHES
#undef GCNB_STREAM_CASE
}

}  // namespace gcnb
