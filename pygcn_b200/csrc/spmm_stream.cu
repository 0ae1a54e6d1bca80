// Streaming CSR SpMM for wide panel rows and power-law graphs:  out = A * B (+ bias) (ReLU)   -- pygcn/layers.py:34-36
// (and dS = A^T * G on the CSR of A^T).
//
// Why a second kernel.  On the products-shaped R-MAT graph (BASELINE configs[3]; median row 3 entries, 43 % of the
// entries in rows of more than 1024) the warp-per-row kernel leaves the machine half empty: every short row pays a
// rowptr -> (col, val) -> gather chain, 8-warp CTAs wait for their longest row (43 % of the warp slots filled), and
// the hub rows need two more launches (profiles/r02_rmat_probe_before.txt: 7.1 ms at F = 256, DRAM 60 %, L2 41 % busy).
// Here
//   * the work is cut by STORED ENTRIES, not rows ("merge-style"): item i = entries [1024 i, 1024 (i + 1)) of the
//     pair stream, whatever rows they belong to, one warp per item: the load is balanced by construction, the
//     (col, val) stream is read sequentially (one coalesced 8-byte load per lane and 32 entries, the next block
//     prefetched), and there is no per-row start-up;
//   * a gathered panel row is read by the whole warp with 128-bit loads (U rows in flight per lane), optionally with
//     an L2 eviction hint: rows of the most referenced columns (heat class tag of the pair, common.cuh) are kept
//     (evict_last), the rest streams through (evict_first);
//   * the end-of-row tag of the pair says when the accumulator is complete: rows inside one item are stored directly
//     (bias / ReLU / dropout epilogue), the two rows an item may share with its neighbours go to scratch as partial
//     sums and a small fix-up kernel adds them in item order (no atomics, run-to-run deterministic).
// Two staged variants were measured and dropped (same schedule, rows copied global -> shared first):
//   one 1-D TMA bulk copy per row -- UBLKCP takes uniform registers, per-lane addresses compile to an ELECT / R2UR
//   waterfall of ~10 dependent instructions per row: 17.4 ms at F = 256 (profiles/r02_rmat_probe_stream_tma_bulk.txt);
//   16-byte cp.async into a per-warp ring -- the per-copy L2 policy is again a uniform-register descriptor (4 R2UR per
//   row) and the shared-memory round trip adds instructions: 88 issued per entry, 22.4 ms
//   (profiles/r02_rmat_probe_stream_cpasync_ring.txt).
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
__device__ __forceinline__ uint4 ldg128(uint64_t p) {
  uint4 v;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ldg128_hint(uint64_t p, uint64_t pol) {
  uint4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ uint2 ldg_pair_stream(const uint2* p, uint64_t pol) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u32 {%0, %1}, [%2], %3;" : "=r"(v.x), "=r"(v.y) : "l"(p), "l"(pol));
  return v;
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

// One warp per item.  CH = 16-byte chunks of a panel row per lane (chunk lane + 32 k), BF16 = panel element type,
// U = gathered rows in flight per lane, HINT = 0 plain loads, 1 hot rows evict_last, 2 + the other rows evict_first.
template <int CH, bool BF16, int U, int HINT>
__global__ void __launch_bounds__(kStreamWarps * 32, (CH * U >= 16) ? 3 : 5)
spmm_stream_kernel(int nnz, int n_items, const uint32_t* __restrict__ items, const uint2* __restrict__ pair,
                   const void* __restrict__ b, uint32_t ldb_bytes, int f, Epilogue ep, float* __restrict__ out,
                   int64_t ldo, int vec_out, float* __restrict__ partial, int ldp, int hot_class_max) {
  constexpr int A = BF16 ? 2 : 1;       // float4 accumulators per chunk
  constexpr int kElems = BF16 ? 8 : 4;  // panel elements per chunk
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * kStreamWarps + (threadIdx.x >> 5);
  if (item >= n_items) return;
  const uint64_t pol_stream = policy_evict_first();
  uint64_t pol_hot = 0, pol_cold = 0;
  if constexpr (HINT >= 1) pol_hot = policy_evict_last();
  if constexpr (HINT >= 2) pol_cold = pol_stream;
  const int nch = (f + kElems - 1) / kElems;  // 16-byte chunks per panel row
  // this lane's chunks of row 0 (lanes past the width read a clamped, valid chunk whose sum is never stored)
  uint64_t bl[CH];
#pragma unroll
  for (int k = 0; k < CH; ++k) bl[k] = reinterpret_cast<uint64_t>(b) + 16u * (uint32_t)min(lane + 32 * k, nch - 1);

  float4 acc[CH * A];
  auto zero_acc = [&]() {
#pragma unroll
    for (int k = 0; k < CH * A; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  };
  zero_acc();
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

  const uint32_t info = __ldg(items + item);
  int row = (int)(info & 0x7fffffffu);
  bool head = (info >> 31) != 0u;  // the current row began in an earlier item: its sum goes to partial[2 * item]
  bool open = false;               // acc holds entries of a row whose end has not been seen
  const int e0 = item * kStreamItem;
  const int e1 = min(e0 + kStreamItem, nnz);
  uint2 p = make_uint2(0u, 0u);
  if (e0 + lane < e1) p = ldg_pair_stream(pair + e0 + lane, pol_stream);
  for (int base = e0; base < e1; base += 32) {
    uint2 pn = make_uint2(0u, 0u);  // the next 32 pairs fly while these are consumed
    if (base + 32 + lane < e1) pn = ldg_pair_stream(pair + base + 32 + lane, pol_stream);
    const int cnt = min(32, e1 - base);
    const uint32_t ends = __ballot_sync(kFull, (p.x & kPairRowEnd) != 0u);  // (lanes past cnt hold zeros)
#pragma unroll 1
    for (int j = 0; j < cnt; j += U) {
      uint4 x[U][CH];
#pragma unroll
      for (int u = 0; u < U; ++u) {  // entries past cnt: column 0, value 0 -- a harmless gather, skipped below
        const uint32_t w = __shfl_sync(kFull, p.x, (j + u) & 31);
        const uint64_t off = (uint64_t)(w & kPairColMask) * ldb_bytes;
        if constexpr (HINT == 0) {
#pragma unroll
          for (int k = 0; k < CH; ++k) x[u][k] = ldg128(bl[k] + off);
        } else {
          const bool hot = (int)((w >> kPairClassShift) & 15u) <= hot_class_max;  // (warp-uniform)
          if (hot) {
#pragma unroll
            for (int k = 0; k < CH; ++k) x[u][k] = ldg128_hint(bl[k] + off, pol_hot);
          } else if (HINT >= 2) {
#pragma unroll
            for (int k = 0; k < CH; ++k) x[u][k] = ldg128_hint(bl[k] + off, pol_cold);
          } else {
#pragma unroll
            for (int k = 0; k < CH; ++k) x[u][k] = ldg128(bl[k] + off);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (j + u < cnt) {  // (warp-uniform)
          const float v = __uint_as_float(__shfl_sync(kFull, p.y, (j + u) & 31));
#pragma unroll
          for (int k = 0; k < CH; ++k) fma_chunk<BF16>(&acc[k * A], v, x[u][k]);
          open = true;
          if ((ends >> (j + u)) & 1u) {
            if (head) store_partial(2 * item); else store_row(row);
            head = false;
            open = false;
            ++row;
            zero_acc();
          }
        }
      }
    }
    p = pn;
  }
  // a row that continues in the next item leaves its partial sum
  if (open) store_partial(2 * item + (head ? 0 : 1));
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
int g_stream_batch = -1;     // 0 auto, else gathered rows in flight per lane (2 / 4 / 8)

void stream_init() {
  if (g_stream_mode >= 0) return;
  const char* e = getenv("GCNB_SPMM_STREAM");
  g_stream_mode = e ? atoi(e) : 1;
  e = getenv("GCNB_L2_HOT_MB");
  g_stream_hot_mb = e ? atoi(e) : 48;
  e = getenv("GCNB_STREAM_HINT");
  g_stream_hint = e ? atoi(e) : 0;
  e = getenv("GCNB_STREAM_MIN_ROW_BYTES");
  g_stream_min_row = e ? atoi(e) : 256;
  e = getenv("GCNB_STREAM_BATCH");
  g_stream_batch = e ? atoi(e) : 0;
}

template <int CH, bool BF16, int U, int HINT>
int launch_stream_t(const CsrView& a, const void* b, int64_t ldb_bytes, int f, const Epilogue& ep, float* out, int64_t ldo,
                    bool vec_out, float* partial, int ldp, int hot_class_max, cudaStream_t st) {
  const int grid = (int)ceil_div(a.n_stream_items, kStreamWarps);
  spmm_stream_kernel<CH, BF16, U, HINT><<<grid, kStreamWarps * 32, 0, st>>>(
      (int)a.nnz, (int)a.n_stream_items, a.stream_items, a.pair, b, (uint32_t)ldb_bytes, f, ep, out, ldo, vec_out ? 1 : 0,
      partial, ldp, hot_class_max);
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
bool spmm_stream_eligible(const CsrView& a, int64_t row_bytes, int nch, bool bf16) {
  stream_init();
  if (g_stream_mode == 0) return false;
  if (!a.pair_tagged || a.stream_items == nullptr || a.n_stream_items == 0 || a.bin_rows[0] != 0) return false;
  if (nch > 64 || row_bytes % 16 != 0 || row_bytes > 1024) return false;
  if (g_stream_mode == 2) return true;
  // auto (same profile): fp32 panel rows of >= 256 bytes on graphs with enough items to fill the machine.  Narrower rows
  // (F = 48: 2.9 ms against 1.9 ms) and bf16 panels (F = 100: 4.7 against 2.6 ms) stay on the row / group kernels: there
  // the gathers, not the row bookkeeping, are the smaller part of the work
  // ... and only on graphs with skewed row lengths.  On a regular graph (Reddit shape: every row ~490 entries, panel
  // half L2-resident) the row kernels' 40 resident warps gather at 19 TB/s, this kernel's 16 at 9.4
  // (profiles/r02_bench_products_v1.json: 12.5 against 6.1-7.3 ms); its gain is the balance and the missing per-row
  // start-up where rows of 1 and of 90 000 entries mix (products shape: 6.5 against 7.1 ms)
  const bool skewed = a.n_rows > 0 && a.max_degree >= 16 * (a.nnz / a.n_rows + 1);
  return !bf16 && skewed && row_bytes >= g_stream_min_row && a.n_stream_items >= 2 * kNumSMs * kStreamWarps;
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
  const int64_t row_bytes = 16 * (int64_t)nch;
  GCNB_REQUIRE(ws != nullptr && ws_bytes >= spmm_stream_workspace_bytes(a, f), "spmm(stream): workspace too small");
  GCNB_REQUIRE(ldb_bytes < (1ll << 32) && ldb_bytes % 16 == 0, "spmm(stream): panel row stride must be a multiple of 16 bytes");
  float* partial = reinterpret_cast<float*>(ws);
  const int ldp = stream_ldp(f);
  // hottest classes whose rows fit the L2 budget: class c holds at most 1024 * 2^c columns
  int hot = -1;
  for (int c = 0; c < 15; ++c)
    if ((1024ll << c) * (int64_t)row_bytes <= (int64_t)g_stream_hot_mb * (1ll << 20)) hot = c;
  // gathered rows in flight per lane (profiles/r02_rmat_probe_stream_register_gathers.txt, products-shaped R-MAT):
  // rows wider than 512 bytes, two 16-byte loads per lane and row: 8 (6.54 ms at F = 256; 6.78 with 4);
  // narrower rows: 4 (3.37 ms at F = 100; 3.75 with 8 -- the 86 registers of the wider batch cost a resident CTA)
  int u = g_stream_batch;
  if (u != 2 && u != 4 && u != 8) u = nch > 32 ? 8 : 4;
  const int hint = g_stream_hint;
#define GCNB_STREAM_CASE(CH_, BF_, U_, H_) \
  return launch_stream_t<CH_, BF_, U_, H_>(a, b, ldb_bytes, f, ep, out, ldo, vec_out, partial, ldp, hot, st)
#define GCNB_STREAM_HINTS(CH_, BF_, U_)                  \
  do {                                                   \
    if (hint == 0) GCNB_STREAM_CASE(CH_, BF_, U_, 0);    \
    if (hint == 1) GCNB_STREAM_CASE(CH_, BF_, U_, 1);    \
    GCNB_STREAM_CASE(CH_, BF_, U_, 2);                   \
  } while (0)
#define GCNB_STREAM_US(CH_, BF_)                 \
  do {                                           \
    if (u == 2) GCNB_STREAM_HINTS(CH_, BF_, 2);  \
    if (u == 4) GCNB_STREAM_HINTS(CH_, BF_, 4);  \
    GCNB_STREAM_HINTS(CH_, BF_, 8);              \
  } while (0)
  if (bf16) {
    if (nch <= 32) GCNB_STREAM_US(1, true);
    GCNB_STREAM_US(2, true);
  } else {
    if (nch <= 32) GCNB_STREAM_US(1, false);
    GCNB_STREAM_US(2, false);
  }
#undef GCNB_STREAM_US
#undef GCNB_STREAM_HINTS
#undef GCNB_STREAM_CASE
}

}  // namespace gcnb
