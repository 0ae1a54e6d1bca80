"""Multi-GPU GCN layer: 1-D row partition of A-hat over the GPUs of one node.

The reference is single-process / single-GPU (no torch.distributed anywhere); this is the
B200-native scaling of its hot path (pygcn/layers.py:32-38 + autograd):

  * rank p owns the contiguous row block p of A-hat (and the matching rows of X, out, G), block
    boundaries balance stored entries (nnz), W and b are replicated;
  * the row block is cut into the diagonal block A[p,p] (columns this rank owns) and the remote
    block A[p,~p] whose column ids address the all-gathered panel directly:
        out_p = A[p,p] . S_p + A[p,~p] . allgather(S) ,      S_q = X_q W  computed on rank q
    the diagonal block runs first while the NCCL all-gather of the panels is in flight over NVLink
    (communicator stream); the remote block accumulates once it has landed and applies bias/ReLU.
    (A finer per-source-rank column blocking was measured first and rejected: every extra pass over
    the rows costs a read-modify-write of the output rows and shorter gather loops; profiles/.)
  * backward uses the same scheme on the row block of A-hat^T with G as the exchanged panel
    (atomic-free, deterministic), then dW/db are summed with one all-reduce.

  * exchange "peer" (default on GPUs with peer access): the all-gather is replaced by our own push
    kernel over NVLink peer memory (csrc/peer.cu) and the row block is cut into one column block per
    source rank, consumed in the order p, p+1, ... as the slots land (flag per source), so the
    transfer of slot q+1 overlaps the SpMM over block q.  `PeerExchange` below; the NCCL path stays as
    exchange "nccl".

One process per GPU (`torchrun`), `torch.distributed` for the plumbing.  The arithmetic is behind
an `ops` object: `CudaOps` (libgcnb200.so) in production; the CPU tests of the host logic
(tests/test_dist_gloo.py, gloo, world size 2) inject a numpy implementation.
"""
from __future__ import annotations

import ctypes
import json
import os
import sys
import time

import torch
import torch.distributed as dist
from torch.autograd.function import once_differentiable


# ---------------------------------------------------------------------------- host logic
def partition_rows_by_nnz(rowptr, world, row_weight=0.0):
    """Row-block boundaries [b_0=0, ..., b_world=N] with ~equal cost per block, cost = stored entries +
    row_weight * rows.

    rowptr: 1-D integer tensor/array of length N+1 (host).  Boundaries are non-decreasing; a block
    may be empty when there are fewer rows than ranks.
    row_weight = 0 (default) balances stored entries alone: right for the SpMM, and on graphs with even
    degrees the blocks then hold equal row counts too.  On power-law graphs they do not (products-shaped R-MAT,
    8 ranks: 21 K ... 909 K rows, tools/halo_fraction.py), and the all-gather ships world * max(rows) padded
    panel rows; row_weight = the average degree weighs both equally and bounds that padding.
    """
    rp = torch.as_tensor(rowptr, dtype=torch.int64).cpu()
    n = rp.numel() - 1
    if row_weight:
        rp = rp.to(torch.float64) + float(row_weight) * torch.arange(n + 1, dtype=torch.float64)
    total = rp[-1].item()
    bounds = [0]
    for k in range(1, world):
        target = (total * k) // world if not row_weight else total * k / world
        r = int(torch.searchsorted(rp, torch.tensor([target], dtype=rp.dtype), right=False)[0])
        r = min(max(r, bounds[-1]), n)
        bounds.append(r)
    bounds.append(n)
    return bounds


def exchange_phases(rank, world):
    """Consumption order of the source ranks for the pipelined exchange, grouped into phases:
    [[p], [p+1, p+2], [p+3, p+4], ..., [last]] (mod world).  One SpMM per phase: this rank's own slot
    first (nothing to wait for), pairs in the middle (a row restricted to ONE of 8 sources is ~12
    stored entries -- too short to gather efficiently), and a single source at the end so that little
    work is left once the last slot has landed."""
    order = [(rank + k) % world for k in range(world)]
    phases = [[order[0]]]
    rest = order[1:]
    while len(rest) > 1:
        phases.append(rest[:2])
        rest = rest[2:]
    if rest:
        phases.append(rest)
    return phases


class DistGraph:
    """Row block `rank` of A-hat and of A-hat^T, each split into the diagonal block (columns owned
    by this rank, local ids) and the remote block (all other columns, ids remapped to the layout of
    the all-gathered panel: source q occupies rows [q*pad_rows, q*pad_rows + n_q))."""

    # Splitting the row block costs a second pass over the rows (read-modify-write of the output,
    # shorter gather loops: measured +35 % SpMM time on the uniform CBG graph), and can hide at most
    # the diagonal block's share of the work behind the all-gather.  It pays only when the partition
    # has locality, i.e. most stored entries sit in the diagonal block.
    SPLIT_MIN_DIAG_FRACTION = 0.6

    def __init__(self, rank, world, bounds, pad_rows, fwd_diag, fwd_remote, bwd_diag, bwd_remote, nnz_local,
                 nnz_global, split=True):
        self.rank, self.world = rank, world
        self.bounds = list(bounds)
        self.pad_rows = pad_rows
        # split=True : (diag, remote) pairs.  split=False: *_diag is None and *_remote holds the whole
        # row block in the all-gather column layout (own slot included).
        self.split = split
        self.fwd_diag, self.fwd_remote = fwd_diag, fwd_remote
        self.bwd_diag, self.bwd_remote = bwd_diag, bwd_remote
        self.nnz_local, self.nnz_global = nnz_local, nnz_global
        # exchange "peer": one column block per phase (group of source ranks, exchange_phases), columns in
        # the gathered layout, for A and A^T
        self.phases = None
        self.fwd_blocks = None
        self.bwd_blocks = None
        # exchange "nccl": ship only the rows each block really holds (allgather_slots); off = padded all-gather
        self.exact_slots = os.environ.get("GCNB_DIST_EXACT_SLOTS", "0") == "1"

    def n_rows(self, q=None):
        q = self.rank if q is None else q
        return self.bounds[q + 1] - self.bounds[q]

    @staticmethod
    def padded_rows(bounds):
        """Rows per source slot of the all-gathered panel: the largest block, rounded up to 8."""
        m = max(bounds[q + 1] - bounds[q] for q in range(len(bounds) - 1))
        return max(8, (m + 7) // 8 * 8)

    @classmethod
    def from_graph(cls, graph, rank, world, bounds=None, split=None, per_source=False, row_weight=0.0):
        """Cut the row block of `rank` out of a full device `Graph` (CUDA).  split=None decides from
        the share of stored entries in the diagonal block (SPLIT_MIN_DIAG_FRACTION).
        per_source=True additionally cuts one column block per source rank (exchange "peer")."""
        from . import _lib
        from .graph import Graph, _stream_ptr

        lib = _lib.load()
        if bounds is None:
            bounds = partition_rows_by_nnz(graph.csr()[0].cpu(), world, row_weight)
        r0, r1 = bounds[rank], bounds[rank + 1]
        pad = cls.padded_rows(bounds)
        hb = (ctypes.c_int64 * (world + 1))(*bounds)

        def cut_diag(transpose):
            with torch.cuda.device(graph.device):
                out = ctypes.c_void_p()
                st = lib.gcnb_graph_block(graph._h, 1 if transpose else 0, r0, r1, r0, r1, 0, max(pad, 1),
                                          _stream_ptr(graph.device), ctypes.byref(out))
                _lib.check(st, "gcnb_graph_block")
                return Graph(out.value, graph.device, "diag[%d]%s" % (rank, "^T" if transpose else ""))

        def cut_gathered(transpose, exclude):
            with torch.cuda.device(graph.device):
                out = ctypes.c_void_p()
                st = lib.gcnb_graph_block_gathered(graph._h, 1 if transpose else 0, r0, r1, world, hb, pad, exclude,
                                                   _stream_ptr(graph.device), ctypes.byref(out))
                _lib.check(st, "gcnb_graph_block_gathered")
                return Graph(out.value, graph.device, "%s[%d]%s" % ("remote" if exclude >= 0 else "rows", rank,
                                                                    "^T" if transpose else ""))

        def cut_sources(transpose, qs):
            mask = 0
            for q in qs:
                mask |= 1 << q
            with torch.cuda.device(graph.device):
                out = ctypes.c_void_p()
                st = lib.gcnb_graph_block_sources(graph._h, 1 if transpose else 0, r0, r1, world, hb, pad, mask,
                                                  _stream_ptr(graph.device), ctypes.byref(out))
                _lib.check(st, "gcnb_graph_block_sources")
                return Graph(out.value, graph.device, "A[%d,%s]%s" % (rank, qs, "^T" if transpose else ""))

        fd = cut_diag(False)
        if world == 1:
            return cls(rank, world, bounds, pad, fd, None, cut_diag(True), None, fd.nnz, graph.nnz, True)
        full = cut_gathered(False, -1)
        if split is None:
            split = fd.nnz >= cls.SPLIT_MIN_DIAG_FRACTION * max(full.nnz, 1)
        if split:
            dg = cls(rank, world, bounds, pad, fd, cut_gathered(False, rank), cut_diag(True),
                     cut_gathered(True, rank), full.nnz, graph.nnz, True)
        else:
            dg = cls(rank, world, bounds, pad, None, full, None, cut_gathered(True, -1), full.nnz, graph.nnz, False)
        if per_source:
            dg.phases = exchange_phases(rank, world)
            dg.fwd_blocks = [cut_sources(False, qs) for qs in dg.phases]
            dg.bwd_blocks = [cut_sources(True, qs) for qs in dg.phases]
        return dg


# ---------------------------------------------------------------------------- partitioned graph build
def _partition_counts(src, dst, n, r0, r1):
    """Stage 1 of the partitioned build on one rank: the rows [r0, r1) of the UN-normalised adjacency
    max(C, C^T) + I (C = edge counts; pygcn/utils.py:360-368) from the edges that touch those rows.
    A row of the symmetrised matrix depends only on the edges incident to its node, so the block is exact
    although the rank never sees the rest of the graph.  Returns (local rowptr int64, global col int64,
    counts fp64, row sums fp64)."""
    from .graph import Graph

    touch = ((src >= r0) & (src < r1)) | ((dst >= r0) & (dst < r1))
    u = Graph.from_edges(src[touch], dst[touch], n, symmetrize=True, self_loops=True, row_normalize=False)
    rowptr, col, val = u.csr()
    e0, e1 = int(rowptr[r0]), int(rowptr[r1])
    lrp = (rowptr[r0:r1 + 1] - e0).to(torch.int64)
    lcol = col[e0:e1].to(torch.int64)
    a = val[e0:e1].to(torch.float64)
    rows = torch.repeat_interleave(torch.arange(r1 - r0, device=src.device), lrp[1:] - lrp[:-1])
    rowsum = torch.zeros(r1 - r0, dtype=torch.float64, device=src.device).index_add_(0, rows, a)  # small integers: exact
    del u
    return lrp, lcol, a, rows, rowsum


def _partition_blocks(rank, world, bounds, pad, lrp, lcol, a, rows, rowsum_global, n_global_nnz=0):
    """Stage 2: with every node's row sum known (all-gathered), normalise like utils.normalize
    (pygcn/utils.py:390-397: r_inv = rowsum^-1 in fp64, inf -> 0, fp32 cast of the product) and build the
    rank's row block of A (values r_inv[row] * a) and of A^T (values r_inv[col] * a: the counts are
    symmetric, so row j of A^T has the pattern of row j of A), columns remapped to the gathered layout."""
    from . import _lib
    from .graph import Graph, _stream_ptr

    dev = lcol.device
    r0, r1 = bounds[rank], bounds[rank + 1]
    r_inv = 1.0 / rowsum_global
    r_inv[torch.isinf(r_inv)] = 0.0
    val_f = (r_inv[r0:r1][rows] * a).to(torch.float32)
    val_b = (r_inv[lcol] * a).to(torch.float32)
    bt = torch.tensor(bounds, dtype=torch.int64, device=dev)
    part = torch.searchsorted(bt, lcol, right=True) - 1
    gcol = (part * pad + (lcol - bt[part])).contiguous()
    lib = _lib.load()

    def make(vals, name):
        out = ctypes.c_void_p()
        v = vals.contiguous()
        crow = lrp.contiguous()
        with torch.cuda.device(dev):
            st = lib.gcnb_graph_from_csr(r1 - r0, world * pad, v.numel(), crow.data_ptr(), gcol.data_ptr(), v.data_ptr(),
                                         _stream_ptr(dev), ctypes.byref(out))
        _lib.check(st, "gcnb_graph_from_csr")
        return Graph(out.value, dev, name)

    fwd = make(val_f, "rows[%d] (partitioned build)" % rank)
    bwd = make(val_b, "rows[%d]^T (partitioned build)" % rank)
    return DistGraph(rank, world, bounds, pad, None, fwd, None, bwd, fwd.nnz, n_global_nnz, False)


def build_partitioned(src, dst, n, rank, world, bounds=None, group=None, row_weight=0.0):
    """This rank's DistGraph from a (replicated) edge list WITHOUT building the whole adjacency on any GPU:
    what a papers100M-sized graph needs (3.2 G stored entries exceed one int32 handle).
    Bit-identical to cutting the block out of the single-GPU `Graph.from_edges` result.  Collective:
    one all-gather of the fp64 row sums (8 B per node) and one all-reduce of the entry counts."""
    dev = src.device
    if bounds is None:  # balance stored entries with the incident-edge count as the estimate
        deg = torch.bincount(src.long(), minlength=n) + torch.bincount(dst.long(), minlength=n) + 1
        rp = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), deg.cumsum(0)])
        bounds = partition_rows_by_nnz(rp, world, row_weight)
    pad = DistGraph.padded_rows(bounds)
    lrp, lcol, a, rows, rowsum = _partition_counts(src, dst, n, bounds[rank], bounds[rank + 1])
    slot = torch.ones(pad, dtype=torch.float64, device=dev)
    slot[: rowsum.numel()] = rowsum
    gathered = torch.empty(world * pad, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_gather_into_tensor(gathered, slot, group=group)
    else:
        gathered.copy_(slot)
    rowsum_global = torch.cat([gathered[q * pad: q * pad + bounds[q + 1] - bounds[q]] for q in range(world)])
    nnz = torch.tensor([a.numel()], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(nnz, group=group)
    return _partition_blocks(rank, world, bounds, pad, lrp, lcol, a, rows, rowsum_global, int(nnz.item()))


# ---------------------------------------------------------------------------- arithmetic backends
class CudaOps:
    """The product backend: every call is a kernel of libgcnb200.so on the current stream."""

    def __init__(self, precision="auto"):
        from . import _lib
        from . import functional as F_

        self._lib = _lib
        self.lib = _lib.load()
        self.F = F_
        self.precision = F_._PRECISIONS[precision]

    def _sp(self, dev):
        return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def gemm(self, a, b, out=None):
        """a [m,k] @ b [k,n] (any strides); written to the first m rows of `out` when given."""
        m, k = a.shape
        n = b.shape[1]
        if out is None:
            out = torch.empty((m, n), dtype=torch.float32, device=a.device)
        with torch.cuda.device(a.device):
            ws = self.F._ws(self.lib.gcnb_gemm_workspace_bytes(m, n, k, self.precision), a.device)
            st = self.lib.gcnb_gemm(m, n, k, a.data_ptr(), a.stride(0), a.stride(1), b.data_ptr(), b.stride(0),
                                    b.stride(1), out.data_ptr(), max(n, 1), self.precision, ws.data_ptr(), ws.numel(),
                                    self._sp(a.device))
        self._lib.check(st, "gcnb_gemm")
        return out

    def spmm_block(self, block, dense, out, accumulate, bias=None, relu=False):
        """out (+)= block @ dense (+ bias) (relu)."""
        f = dense.shape[1]
        flags = (self._lib.SPMM_ACCUMULATE if accumulate else 0) | (self._lib.SPMM_RELU if relu else 0)
        with torch.cuda.device(dense.device):
            ws = self.F._ws(self.lib.gcnb_spmm_workspace_bytes(block._h, 0, f), dense.device)
            st = self.lib.gcnb_spmm(block._h, flags, dense.data_ptr(), dense.stride(0) if dense.shape[0] > 1 else f, f,
                                    bias.data_ptr() if bias is not None else None, out.data_ptr(), out.stride(0)
                                    if out.shape[0] > 1 else f, ws.data_ptr(), ws.numel(), self._sp(dense.device))
        self._lib.check(st, "gcnb_spmm")
        return out

    def colsum(self, g, y=None, gm=None):
        """(column sums of g [masked by y > 0], the masked g).  When `gm` (>= n rows) is given the
        masked gradient -- or a plain copy of g -- is staged into its first n rows."""
        n, f = g.shape
        out = torch.empty((f,), dtype=torch.float32, device=g.device)
        if gm is None and y is not None:
            gm = torch.empty_like(g)
        with torch.cuda.device(g.device):
            ws = self.F._ws(self.lib.gcnb_colsum_workspace_bytes(n, f), g.device)
            st = self.lib.gcnb_colsum(n, f, g.data_ptr(), g.stride(0) if n > 1 else f,
                                      y.data_ptr() if y is not None else None, f,
                                      gm.data_ptr() if gm is not None else None, f, out.data_ptr(), ws.data_ptr(),
                                      ws.numel(), self._sp(g.device))
        self._lib.check(st, "gcnb_colsum")
        return out, (gm if gm is not None else g)

    def empty(self, shape, like):
        return torch.empty(shape, dtype=torch.float32, device=like.device)

    # -- the bf16 panel tier (precision="bf16", <= 2e-2): the exchanged panel is rounded once and gathered at half the bytes
    def to_bf16(self, panel):
        """[rows, f] fp32 -> [rows, 8 * ceil(f / 8)] bf16 (round to nearest even, zero padded): gcnb_to_bf16."""
        rows, f = panel.shape
        ld8 = (f + 7) // 8 * 8
        out = torch.empty((rows, ld8), dtype=torch.bfloat16, device=panel.device)
        with torch.cuda.device(panel.device):
            st = self.lib.gcnb_to_bf16(rows, f, panel.data_ptr(), panel.stride(0) if rows > 1 else f, out.data_ptr(), ld8,
                                       self._sp(panel.device))
        self._lib.check(st, "gcnb_to_bf16")
        return out

    def empty_bf16(self, shape, like):
        return torch.empty(shape, dtype=torch.bfloat16, device=like.device)

    def spmm_block_bf16(self, block, dense, f, out, accumulate, bias=None, relu=False):
        """out (+)= block @ dense[:, :f] (+ bias) (relu) with a bf16 panel `dense` [n_cols, ld8]: gcnb_spmm_bf16
        (fp32 accumulation and output)."""
        flags = (self._lib.SPMM_ACCUMULATE if accumulate else 0) | (self._lib.SPMM_RELU if relu else 0)
        with torch.cuda.device(dense.device):
            ws = self.F._ws(self.lib.gcnb_spmm_workspace_bytes(block._h, 0, f), dense.device)
            st = self.lib.gcnb_spmm_bf16(block._h, flags, dense.data_ptr(), dense.stride(0) if dense.shape[0] > 1 else
                                         dense.shape[1], f, bias.data_ptr() if bias is not None else None, out.data_ptr(),
                                         out.stride(0) if out.shape[0] > 1 else f, ws.data_ptr(), ws.numel(),
                                         self._sp(dense.device))
        self._lib.check(st, "gcnb_spmm_bf16")
        return out

    # -- build-time helpers of the halo exchange (HaloPlan): blocks as CSR tensors and back
    def block_csr(self, block):
        rowptr, col, val = block.csr()
        return rowptr.long(), col.long(), val

    def block_from_csr(self, rowptr, col, val, n_rows, n_cols):
        from .graph import Graph

        adj = torch.sparse_csr_tensor(rowptr, col, val, size=(n_rows, n_cols), check_invariants=False)
        return Graph.from_torch(adj)

    def selection_block(self, ids, n_cols):
        """[len(ids), n_cols] with a single 1.0 per row: block @ dense packs the rows `ids` of dense (exact copies)."""
        k = ids.numel()
        rowptr = torch.arange(k + 1, dtype=torch.int64, device=ids.device)
        return self.block_from_csr(rowptr, ids.long(), torch.ones(k, dtype=torch.float32, device=ids.device), k, n_cols)


# ---------------------------------------------------------------------------- peer-memory exchange
class _DevMem:
    """A raw device allocation seen as a tensor through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"data": (ptr, False), "shape": tuple(shape), "typestr": typestr, "version": 3,
                                         "strides": None}


class PeerExchangeUnavailable(RuntimeError):
    """CUDA IPC / peer access is not available between the ranks (raised on EVERY rank, after a collective
    agreement, so that all of them fall back to the NCCL exchange together)."""


class PeerExchange:
    """One panel exchange buffer: gathered [world * pad_rows, f] fp32 on every rank, written by the
    peers' push kernels over NVLink peer memory, plus flag / ack words per source rank (csrc/peer.cu).

    Protocol per exchange (all on device, CUDA-graph replayable):
      start(): epoch += 1 on the compute stream, then on the communication stream the push kernel
               copies this rank's slot to ranks p-1, p-2, ... and publishes flag[p] = epoch there;
      wait(q): the compute stream spins until slot q has landed;  done(q): acks it to rank q so q's next
               push may overwrite it;  finish(): joins the communication stream."""

    CTRL_BYTES = 1024
    PUSH_CTAS = 32
    # "ce": the slot is moved by the copy engines (cudaMemcpyAsync to the mapped peer pointer, 635 GB/s
    # per peer pair measured, no SMs taken from the SpMM) between our wait-for-ack and publish-flag
    # kernels; "sm": our push kernel stores it over NVLink itself (P2P stores, ~380-460 GB/s with 32-296
    # CTAs, tools/peer_bw.py).  GCNB_PEER_PUSH overrides.
    PUSH_MODE = os.environ.get("GCNB_PEER_PUSH", "ce")

    def __init__(self, rank, world, pad_rows, f, device, group=None):
        from . import _lib

        self._lib = _lib
        self.lib = _lib.load()
        self.rank, self.world, self.pad_rows, self.f = rank, world, pad_rows, f
        self.device = device
        if world - 1 > 15:
            raise RuntimeError("peer exchange supports at most 16 ranks")
        self.slot_bytes = pad_rows * f * 4
        data = (world * self.slot_bytes + 255) // 256 * 256
        self.off_flags, self.off_acks = data, data + 256
        self.off_epoch, self.off_counters = data + 512, data + 768
        self.bytes = data + self.CTRL_BYTES
        ok, why = 1, ""
        self.ptr, self.peer_ptr = None, {}
        with torch.cuda.device(device):
            ptr = ctypes.c_void_p()
            handle = ctypes.create_string_buffer(64)
            try:
                _lib.check(self.lib.gcnb_symm_alloc(self.bytes, ctypes.byref(ptr), handle), "gcnb_symm_alloc")
                self.ptr = ptr.value
            except Exception as e:  # no IPC in this environment: every rank must learn it (collectives below)
                ok, why = 0, repr(e)
            handles = [None] * world
            dist.all_gather_object(handles, handle.raw if ok else None, group=group)
            if ok and all(h is not None for h in handles):
                try:
                    for q in range(world):
                        if q == rank:
                            continue
                        pp = ctypes.c_void_p()
                        _lib.check(self.lib.gcnb_symm_open(ctypes.create_string_buffer(handles[q], 64), ctypes.byref(pp)),
                                   "gcnb_symm_open")
                        self.peer_ptr[q] = pp.value
                except Exception as e:
                    ok, why = 0, repr(e)
            else:
                ok = 0
            flag = torch.tensor([float(ok)], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)  # also: every rank has mapped every buffer
            if flag.item() < 1:
                self.close()
                raise PeerExchangeUnavailable("peer-memory exchange unavailable on at least one rank (%s)" % (why or "peer"))
        # push order: the rank that consumes our slot first (p-1) is served first
        order = [(rank - k) % world for k in range(1, world)]
        n = len(order)
        self._n_peers = n
        self._dst = (ctypes.c_void_p * max(n, 1))(*[self.peer_ptr[r] + rank * self.slot_bytes for r in order])
        self._flag = (ctypes.c_void_p * max(n, 1))(*[self.peer_ptr[r] + self.off_flags + 4 * rank for r in order])
        self._ack = (ctypes.c_void_p * max(n, 1))(*[self.ptr + self.off_acks + 4 * r for r in order])
        self.gathered = torch.as_tensor(_DevMem(self.ptr, (world * pad_rows, f)), device=device)
        self.my_slot = self.gathered[rank * pad_rows:(rank + 1) * pad_rows]
        self.comm = torch.cuda.Stream(device=device, priority=-1)
        self._ev = torch.cuda.Event()

    def _sp(self, stream=None):
        return ctypes.c_void_p((stream or torch.cuda.current_stream(self.device)).cuda_stream)

    def start(self):
        lib, ck = self.lib, self._lib.check
        with torch.cuda.device(self.device):
            ck(lib.gcnb_peer_epoch_bump(self.ptr + self.off_epoch, self.ptr + self.off_counters, self._n_peers, self._sp()),
               "gcnb_peer_epoch_bump")
            self._ev.record(torch.cuda.current_stream(self.device))
            self.comm.wait_event(self._ev)
            cs = self._sp(self.comm)
            if self.PUSH_MODE == "sm":
                ck(lib.gcnb_peer_push(self.ptr + self.rank * self.slot_bytes, self.slot_bytes, self._n_peers, self._dst,
                                      self._flag, self._ack, self.ptr + self.off_epoch, self.ptr + self.off_counters,
                                      self.PUSH_CTAS, cs), "gcnb_peer_push")
            else:
                ep = self.ptr + self.off_epoch
                for k in range(self._n_peers):
                    ck(lib.gcnb_peer_wait_lag(self._ack[k], ep, 1, cs), "gcnb_peer_wait_lag")  # peer done with the old slot
                    ck(lib.gcnb_peer_copy(self._dst[k], self.ptr + self.rank * self.slot_bytes, self.slot_bytes, cs),
                       "gcnb_peer_copy")
                    ck(lib.gcnb_peer_ack(self._flag[k], ep, cs), "gcnb_peer_ack")               # publish flag = epoch

    def wait(self, q):
        if q != self.rank:
            with torch.cuda.device(self.device):
                self._lib.check(self.lib.gcnb_peer_wait(self.ptr + self.off_flags + 4 * q, self.ptr + self.off_epoch,
                                                        self._sp()), "gcnb_peer_wait")

    def done(self, q):
        if q != self.rank:
            with torch.cuda.device(self.device):
                self._lib.check(self.lib.gcnb_peer_ack(self.peer_ptr[q] + self.off_acks + 4 * self.rank,
                                                       self.ptr + self.off_epoch, self._sp()), "gcnb_peer_ack")

    def finish(self):
        torch.cuda.current_stream(self.device).wait_stream(self.comm)

    def close(self):
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in getattr(self, "peer_ptr", {}).values():
                self.lib.gcnb_symm_close(p)
            self.peer_ptr = {}
            if getattr(self, "ptr", None):
                self.lib.gcnb_symm_free(self.ptr)
            self.ptr = None


class CollectiveExchange:
    """The same interface on top of one all-gather (gloo in the CPU tests of the host logic, or NCCL):
    start() gathers every slot at once, wait / done are no-ops."""

    def __init__(self, rank, world, pad_rows, f, like, group=None):
        self.rank, self.world, self.pad_rows, self.f, self.group = rank, world, pad_rows, f, group
        self.gathered = torch.zeros((world * pad_rows, f), dtype=torch.float32, device=like.device)
        self.my_slot = self.gathered[rank * pad_rows:(rank + 1) * pad_rows]

    def start(self):
        dist.all_gather_into_tensor(self.gathered, self.my_slot.clone(), group=self.group)

    def wait(self, q):
        pass

    def done(self, q):
        pass

    def finish(self):
        pass

    def close(self):
        pass


class MulticastExchange:
    """Exchange "nvls" (opt-in; written after round 1's multi-GPU minutes were spent -- NOT yet run on hardware):
    gathered [world * pad_rows, f] lives in torch's symmetric memory (torch.distributed._symmetric_memory: a cuMem
    allocation every peer maps, bound to an NVLS multicast object -- plumbing), and the all-gather is OUR kernel:
    every rank stores its slot once to the multicast address (gcnb_multimem_push, multimem.st) and the NVSwitch
    replicates it into all ranks' buffers, so a GPU sends one slot instead of world - 1 and the exchange is bounded
    by its ingress alone.  allgather() = barrier (every rank is done reading the previous contents) -> push ->
    barrier (every slot has landed); both barriers are the symmetric-memory handle's device barriers on the current
    stream, the whole sequence is stream-ordered and capturable."""

    PUSH_CTAS = int(os.environ.get("GCNB_NVLS_PUSH_CTAS", "64"))

    def __init__(self, rank, world, pad_rows, f, device, group=None):
        from . import _lib

        self._lib = _lib
        self.lib = _lib.load()
        self.rank, self.world, self.pad_rows, self.f, self.device = rank, world, pad_rows, f, device
        self.slot_bytes = pad_rows * f * 4
        ok, why = 1, ""
        self.hdl = None
        try:
            import torch.distributed._symmetric_memory as symm

            pg = group if group is not None else dist.group.WORLD
            self._group_name = pg.group_name
            with torch.cuda.device(device):
                self.gathered = symm.empty((world * pad_rows, f), dtype=torch.float32, device=device)
                self.hdl = symm.rendezvous(self.gathered, pg)
            if not int(self.hdl.multicast_ptr):  # 0 when the group has no NVLS multicast object
                ok, why = 0, "no NVLS multicast support for this group"
        except Exception as e:  # every rank must learn it (collective below)
            ok, why = 0, repr(e)
        flag = torch.tensor([float(ok)], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if flag.item() < 1:
            raise PeerExchangeUnavailable("multicast exchange unavailable on at least one rank (%s)" % (why or "peer"))
        self.mc_slot = int(self.hdl.multicast_ptr) + rank * self.slot_bytes
        self.my_slot = self.gathered[rank * pad_rows:(rank + 1) * pad_rows]

    def allgather(self):
        """Every rank's my_slot -> every rank's gathered, in stream order on the current stream."""
        if os.environ.get("GCNB_NVLS_LIBRARY_PUSH") == "1":
            # debugging aid for the first hardware run: torch's own multimem all-gather over the same symmetric
            # buffer (barrier, multimem.st, barrier in one library kernel) -- tells a bug in our push from an
            # environment without NVLS.  Never the product path.
            torch.ops.symm_mem.multimem_all_gather_out(self.my_slot.clone(), self._group_name, self.gathered)
            return
        with torch.cuda.device(self.device):
            self.hdl.barrier(channel=0)  # nobody still reads the slots of the previous exchange
            self._lib.check(self.lib.gcnb_multimem_push(self.mc_slot, self.my_slot.data_ptr(), self.slot_bytes, self.PUSH_CTAS,
                                                        ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)),
                            "gcnb_multimem_push")
            self.hdl.barrier(channel=1)  # every slot has landed everywhere

    def close(self):
        self.hdl = None
        self.gathered = self.my_slot = None


def dist_spmm_gathered(ops, dgraph, diag, remote, exch, bias=None, relu=False):
    """out_p over an exchange object that gathers in place (MulticastExchange): this rank's slot is already in
    exch.my_slot; the diagonal block (if the row block is split) runs before the exchange."""
    out = ops.empty((dgraph.n_rows(), exch.f), exch.gathered)
    if dgraph.split:
        ops.spmm_block(diag, exch.my_slot, out, False)
    exch.allgather()
    return ops.spmm_block(remote, exch.gathered, out, dgraph.split, bias, relu)


def dist_spmm_pipelined(ops, dgraph, blocks, exch, bias=None, relu=False):
    """out_p = sum_i blocks[i] @ gathered (+ bias) (relu) over the phases of dgraph.phases (groups of
    source ranks p, p+1, ...): this rank's slot is already in exch.my_slot; every other slot is consumed
    as soon as it has landed (exch.wait) and released to its owner right after (exch.done).  The bias /
    ReLU epilogue rides on the last block."""
    out = ops.empty((dgraph.n_rows(), exch.f), exch.gathered)
    exch.start()
    n = len(dgraph.phases)
    for i, qs in enumerate(dgraph.phases):
        last = i == n - 1
        for q in qs:
            exch.wait(q)
        ops.spmm_block(blocks[i], exch.gathered, out, i > 0, bias if last else None, relu and last)
        for q in qs:
            exch.done(q)
    exch.finish()
    return out


# ---------------------------------------------------------------------------- the exchange + layer
class HaloPlan:
    """Needed-rows-only ("halo") exchange for one direction of one DistGraph (opt-in, exchange "halo"; host logic
    covered by the gloo tests, composed of kernels that are measured -- the packing is an SpMM with a selection
    block -- but NOT yet run on GPUs as a whole).

    The all-gather ships every row of every slot; a rank only reads the rows that appear as a column in its block:
    95.6 % of them on a uniform graph, 35 % on a products-shaped R-MAT graph (tools/halo_fraction.py).  Built once
    (collective: the lists of needed rows are exchanged): need[q] = sorted local ids of source q's rows this rank
    reads; give[r] = ids of this rank's rows that rank r reads; pack[r] = selection block so that pack[r] @ panel is
    the send buffer for r; block = the rank's (remote or whole) row block with its columns renumbered to the compact
    panel [own slot (unsplit blocks only) | halo of source 0 | halo of source 1 | ...]."""

    def __init__(self, ops, dgraph, block, group=None):
        p, world, pad = dgraph.rank, dgraph.world, dgraph.pad_rows
        rowptr, col, val = ops.block_csr(block)
        src = torch.div(col, pad, rounding_mode="floor")
        loc = col - src * pad
        self.own = 0 if dgraph.split else pad  # a split row block keeps its own columns in the diagonal block
        self.need, self.offset = {}, {}
        new_col = torch.empty_like(col)
        mine = src == p
        new_col[mine] = loc[mine]
        off = self.own
        for q in range(world):
            if q == p:
                continue
            m = src == q
            ids = torch.unique(loc[m])  # sorted
            self.need[q], self.offset[q] = ids, off
            if ids.numel():
                new_col[m] = off + torch.searchsorted(ids, loc[m])
            off += ids.numel()
        self.n_compact = max(off, 1)
        lists = [None] * world
        dist.all_gather_object(lists, {q: v.cpu() for q, v in self.need.items()}, group=group)
        self.give = {r: lists[r][p].to(col.device) for r in range(world) if r != p}
        n_p = dgraph.n_rows()
        self.pack = {r: ops.selection_block(ids, max(pad, 1)) for r, ids in self.give.items() if ids.numel()}
        self.block = ops.block_from_csr(rowptr, new_col, val, n_p, self.n_compact)
        self.rows_received = off - self.own
        self.rows_all_gather = (world - 1) * pad


def dist_spmm_halo(ops, dgraph, plan, diag, panel, bias=None, relu=False, group=None):
    """out_p with the needed-rows-only exchange: every peer r gets pack[r] @ panel (the rows of this rank it reads),
    this rank receives need[q] rows from every source q into the compact panel, then one SpMM over the renumbered
    block (after the diagonal block when the row block is split)."""
    f = panel.shape[1]
    p, world = dgraph.rank, dgraph.world
    out = ops.empty((dgraph.n_rows(), f), panel)
    compact = ops.empty((plan.n_compact, f), panel)
    sends, p2p = [], []
    for k in range(1, world):
        r = (p + k) % world
        if r in plan.pack:
            buf = ops.empty((plan.give[r].numel(), f), panel)
            ops.spmm_block(plan.pack[r], panel, buf, False)
            sends.append(buf)
            p2p.append(dist.P2POp(dist.isend, buf, r if group is None else dist.get_global_rank(group, r), group))
        q = (p - k) % world
        n_q = plan.need[q].numel()
        if n_q:
            p2p.append(dist.P2POp(dist.irecv, compact[plan.offset[q]: plan.offset[q] + n_q],
                                  q if group is None else dist.get_global_rank(group, q), group))
    reqs = dist.batch_isend_irecv(p2p) if p2p else []
    if dgraph.split:
        ops.spmm_block(diag, panel, out, False)           # runs while the halo rows are on the wire
    else:
        compact[: panel.shape[0]].copy_(panel)            # own slot at the front of the compact panel
    for rq in reqs:
        rq.wait()
    return ops.spmm_block(plan.block, compact, out, dgraph.split, bias, relu)


class _Works:
    """wait() on a list of request objects (batch_isend_irecv) like on one collective's work handle."""

    def __init__(self, reqs):
        self.reqs = reqs

    def wait(self):
        for r in self.reqs:
            r.wait()


def allgather_slots(gathered, panel, dgraph, group=None, async_op=False):
    """Every rank's slot `panel` [pad_rows, w] -> gathered [world * pad_rows, w] on every rank.

    Default: one all_gather_into_tensor of the padded slots.  With dgraph.exact_slots (opt-in: DistGraph.exact_slots =
    True or GCNB_DIST_EXACT_SLOTS=1) only the rows a block really holds travel: one grouped batch of point-to-point
    sends / receives (the same slot to every peer, every peer's n_q rows into its place), which is what skewed
    partitions need -- on a products-shaped R-MAT graph the padded slots are 3x the real rows (tools/halo_fraction.py).
    The layout of `gathered` (source q at row q * pad_rows) is the same either way, so the blocks do not change."""
    if not getattr(dgraph, "exact_slots", False):
        return dist.all_gather_into_tensor(gathered, panel, group=group, async_op=async_op)
    world, p, pad = dgraph.world, dgraph.rank, dgraph.pad_rows
    n_p = dgraph.n_rows()
    if n_p:
        gathered[p * pad: p * pad + n_p].copy_(panel[:n_p])
    ops_ = []
    for k in range(1, world):
        dst, src = (p + k) % world, (p - k) % world
        if n_p:
            ops_.append(dist.P2POp(dist.isend, panel[:n_p], dst if group is None else dist.get_global_rank(group, dst), group))
        n_s = dgraph.n_rows(src)
        if n_s:
            ops_.append(dist.P2POp(dist.irecv, gathered[src * pad: src * pad + n_s],
                                   src if group is None else dist.get_global_rank(group, src), group))
    works = _Works(dist.batch_isend_irecv(ops_) if ops_ else [])
    if async_op:
        return works
    works.wait()
    return None


def chunk_columns(f, chunks):
    """Column ranges [(c0, c1), ...] that cut a width-f panel into at most `chunks` pieces whose
    boundaries are multiples of 4 floats (gathered rows stay 16-byte aligned); fewer pieces when f
    is too narrow to give every piece at least 8 columns."""
    chunks = max(1, min(int(chunks), f // 8 if f >= 8 else 1))
    quads = (f + 3) // 4
    cuts = [min(f, 4 * ((quads * k + chunks - 1) // chunks)) for k in range(chunks + 1)]
    cuts[-1] = f
    return [(cuts[k], cuts[k + 1]) for k in range(chunks) if cuts[k + 1] > cuts[k]]


def dist_spmm_chunked(ops, dgraph, block, panel, chunks, bias=None, relu=False, group=None):
    """The unsplit row block against the all-gathered panel, pipelined over COLUMN chunks of the panel:
    every chunk's all-gather is issued at once (they queue on the communicator's stream), and the SpMM
    over chunk k (width f/chunks, its own slice of the output, bias slice and ReLU fused) starts when
    that chunk has landed -- while chunk k+1 is still on the wire.  No extra kernels, no protocol state:
    the exchange stays NCCL's, only its granularity changes.  The price is the adjacency's index stream,
    read once per chunk instead of once, and narrower gathered rows."""
    world = dgraph.world
    f = panel.shape[1]
    out = ops.empty((dgraph.n_rows(), f), panel)
    pending = []
    for c0, c1 in chunk_columns(f, chunks):
        send = panel[:, c0:c1].contiguous()
        gathered = ops.empty((world * dgraph.pad_rows, c1 - c0), panel)
        work = allgather_slots(gathered, send, dgraph, group, async_op=True)
        pending.append((c0, c1, gathered, work, send))
    for c0, c1, gathered, work, _send in pending:
        work.wait()
        ops.spmm_block(block, gathered, out[:, c0:c1], False, bias[c0:c1] if bias is not None else None, relu)
    return out


def dist_spmm(ops, dgraph, diag, remote, panel, bias=None, relu=False, group=None, chunks=1):
    """out_p = diag @ panel[:n_p] + remote @ allgather(panel) (+ bias) (relu).

    `panel` is this rank's [pad_rows, F] slot (rows past n_p are padding nobody references).  The
    all-gather of the slots (NCCL over NVLink, on the communicator's stream) runs while the
    diagonal block is multiplied; the remote block is accumulated once the panel has landed.
    chunks > 1 (unsplit row blocks only): dist_spmm_chunked."""
    world = dgraph.world
    if world > 1 and not dgraph.split and chunks > 1 and len(chunk_columns(panel.shape[1], chunks)) > 1:
        return dist_spmm_chunked(ops, dgraph, remote, panel, chunks, bias, relu, group)
    out = ops.empty((dgraph.n_rows(), panel.shape[1]), panel)
    if world == 1:
        return ops.spmm_block(diag, panel, out, False, bias, relu)
    gathered = ops.empty((world * dgraph.pad_rows, panel.shape[1]), panel)
    if not dgraph.split:  # no locality to exploit: one pass over the whole row block
        allgather_slots(gathered, panel, dgraph, group)
        return ops.spmm_block(remote, gathered, out, False, bias, relu)
    work = allgather_slots(gathered, panel, dgraph, group, async_op=True)
    ops.spmm_block(diag, panel, out, False)
    work.wait()
    return ops.spmm_block(remote, gathered, out, True, bias, relu)


def dist_spmm_bf16(ops, dgraph, diag, remote, panel, bias=None, relu=False, group=None):
    """dist_spmm with the exchanged panel in bf16 (DistGraphConvolution(precision="bf16"), the <= 2e-2 tier; opt-in, a
    composition of measured kernels -- gcnb_to_bf16, gcnb_spmm_bf16 -- not yet run over NCCL): this rank's slot is rounded
    to bf16 once, the all-gather moves HALF the bytes (the exchange is what bounds the multi-GPU step),
    and the SpMM gathers half the bytes per row; accumulation, bias, ReLU and the output stay fp32.  Every rank rounds
    the same values the single-GPU bf16 tier rounds, so the result equals that tier's to summation order."""
    world = dgraph.world
    f = panel.shape[1]
    out = ops.empty((dgraph.n_rows(), f), panel)
    send = ops.to_bf16(panel)
    if world == 1:
        return ops.spmm_block_bf16(diag, send, f, out, False, bias, relu)
    gathered = ops.empty_bf16((world * dgraph.pad_rows, send.shape[1]), panel)
    if not dgraph.split:
        allgather_slots(gathered, send, dgraph, group)
        return ops.spmm_block_bf16(remote, gathered, f, out, False, bias, relu)
    work = allgather_slots(gathered, send, dgraph, group, async_op=True)
    ops.spmm_block_bf16(diag, send, f, out, False)
    work.wait()
    return ops.spmm_block_bf16(remote, gathered, f, out, True, bias, relu)


def halo_plans(ops, dgraph, group=None):
    """(forward, backward) HaloPlan of a DistGraph, built on first use (collective) and kept on the graph."""
    plans = getattr(dgraph, "_halo_plans", None)
    if plans is None:
        plans = (HaloPlan(ops, dgraph, dgraph.fwd_remote, group), HaloPlan(ops, dgraph, dgraph.bwd_remote, group))
        dgraph._halo_plans = plans
    return plans


def dist_layer_forward(ops, dgraph, x, w, b, relu=False, group=None, exch=None, chunks=1, bf16=False):
    """Row block of  A (X W) + b  (pygcn/layers.py:33-36) for this rank.  With `exch` (an exchange
    object for [pad_rows, Fout] panels) the per-source-block pipelined scheme is used; `chunks` > 1
    pipelines the NCCL exchange over column chunks of the panel instead (dist_spmm_chunked)."""
    if exch is not None and exch != "halo" and dgraph.world > 1:
        ops.gemm(x, w, out=exch.my_slot)             # X_p W straight into this rank's slot
        if hasattr(exch, "allgather"):  # gathers in place (MulticastExchange)
            return dist_spmm_gathered(ops, dgraph, dgraph.fwd_diag, dgraph.fwd_remote, exch, b, relu)
        return dist_spmm_pipelined(ops, dgraph, dgraph.fwd_blocks, exch, b, relu)
    support = ops.empty((dgraph.pad_rows, w.shape[1]), x)
    ops.gemm(x, w, out=support)
    if exch == "halo" and dgraph.world > 1:
        return dist_spmm_halo(ops, dgraph, halo_plans(ops, dgraph, group)[0], dgraph.fwd_diag, support, b, relu, group)
    if bf16:
        return dist_spmm_bf16(ops, dgraph, dgraph.fwd_diag, dgraph.fwd_remote, support, b, relu, group)
    return dist_spmm(ops, dgraph, dgraph.fwd_diag, dgraph.fwd_remote, support, b, relu, group, chunks)


def dist_layer_backward(ops, dgraph, x, w, g, y=None, need_dx=True, has_bias=True, group=None, exch=None, chunks=1,
                        bf16=False):
    """(dX rows of this rank or None, dW, db): dW/db are already summed over ranks."""
    fin, fout = w.shape
    if exch is not None and exch != "halo" and dgraph.world > 1:
        db, _ = ops.colsum(g, y, exch.my_slot)       # local part of db; G (masked) staged into this rank's slot
        if hasattr(exch, "allgather"):  # gathers in place (MulticastExchange)
            ds = dist_spmm_gathered(ops, dgraph, dgraph.bwd_diag, dgraph.bwd_remote, exch)
        else:
            ds = dist_spmm_pipelined(ops, dgraph, dgraph.bwd_blocks, exch)  # rows p of A^T G
    else:
        gm = ops.empty((dgraph.pad_rows, fout), g)
        db, _ = ops.colsum(g, y, gm)                 # local part of db; G (masked) staged into its slot
        if exch == "halo" and dgraph.world > 1:
            ds = dist_spmm_halo(ops, dgraph, halo_plans(ops, dgraph, group)[1], dgraph.bwd_diag, gm, None, False, group)
        elif bf16:
            ds = dist_spmm_bf16(ops, dgraph, dgraph.bwd_diag, dgraph.bwd_remote, gm, None, False, group)
        else:
            ds = dist_spmm(ops, dgraph, dgraph.bwd_diag, dgraph.bwd_remote, gm, None, False, group, chunks)  # rows p of A^T G
    dw = ops.gemm(x.t(), ds)                         # local part of X^T dS
    if dgraph.world > 1:
        flat = torch.cat([dw.reshape(-1), db.reshape(-1)])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        dw = flat[: fin * fout].reshape(fin, fout)
        db = flat[fin * fout:]
    dx = ops.gemm(ds, w.t()) if need_dx else None
    return dx, dw, (db if has_bias else None)


class _DistGCNLayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, dgraph, relu, ops, group, exch_f=None, exch_b=None, chunks=1, bf16=False):
        out = dist_layer_forward(ops, dgraph, x, weight, bias, relu, group, exch_f, chunks, bf16)
        ctx.dgraph, ctx.relu, ctx.ops, ctx.group, ctx.exch_b, ctx.chunks = dgraph, relu, ops, group, exch_b, chunks
        ctx.bf16 = bf16
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight, out if relu else None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        x, w, y = ctx.saved_tensors
        dx, dw, db = dist_layer_backward(ctx.ops, ctx.dgraph, x, w, g.contiguous(), y, ctx.needs_input_grad[0],
                                         ctx.has_bias, ctx.group, ctx.exch_b, ctx.chunks, ctx.bf16)
        return dx, dw, db, None, None, None, None, None, None, None, None


class DistGraphConvolution(torch.nn.Module):
    """`GraphConvolution` over a row-partitioned graph: forward(x_local, dist_graph) -> out_local.
    Parameters are replicated (same seed on every rank = same init as the reference layer);
    `.grad` of weight/bias comes out already all-reduced."""

    def __init__(self, in_features, out_features, bias=True, *, fuse_relu=False, precision="auto", group=None,
                 exchange="auto", nccl_chunks=None):
        super().__init__()
        from .layers import GraphConvolution

        if exchange not in ("auto", "peer", "nccl", "nvls", "halo"):
            raise ValueError("exchange must be 'auto', 'peer', 'nccl', 'nvls' or 'halo'")
        self.inner = GraphConvolution(in_features, out_features, bias, fuse_relu=fuse_relu, precision=precision)
        self.group = group
        self.exchange = exchange
        # exchange "nccl" on an unsplit row block: number of column chunks the panel all-gather is pipelined
        # over (dist_spmm_chunked); 1 = one all-gather, then one SpMM (the measured default)
        self.nccl_chunks = int(os.environ.get("GCNB_DIST_NCCL_CHUNKS", "1")) if nccl_chunks is None else int(nccl_chunks)
        if self.nccl_chunks < 1:
            raise ValueError("nccl_chunks must be >= 1")
        self._ops = None
        self._exch = None  # (dgraph id, forward exchange, backward exchange): per layer, so that a slot is
        #                    never overwritten by another layer's panel while a peer still reads it

    @staticmethod
    def resolve_exchange(exchange, world):
        """'auto': measured on 8 x B200 (profiles/r01_dist_probe8.txt, r01_bench_n*_peer/nccl.json): the
        pipelined peer-memory exchange wins at 2 GPUs (0.325 vs 0.380 ms/step); from 4 GPUs on its 2
        small kernels per peer wait for SM slots behind the SpMM's CTAs and the phase-split SpMM costs
        +38 %, so one NCCL all-gather + one SpMM is faster (0.704 vs 0.853 ms at 8)."""
        if exchange == "auto":
            return "peer" if world == 2 else "nccl"
        return exchange

    def _exchanges(self, dgraph, dev):
        kind = self.resolve_exchange(self.exchange, dgraph.world)
        if self.inner.precision == "bf16":  # bf16 panels travel through the all-gather exchange only (dist_spmm_bf16)
            if self.exchange not in ("auto", "nccl") or self.nccl_chunks != 1:
                raise ValueError("precision='bf16' of the row-partitioned layer needs exchange 'auto' / 'nccl' and nccl_chunks=1")
            return None, None
        if kind == "halo" and dgraph.world > 1:
            return "halo", "halo"  # needed-rows-only exchange: the plans live on the DistGraph (halo_plans)
        if dgraph.world == 1 or kind == "nccl" or (kind == "peer" and dgraph.fwd_blocks is None):
            return None, None
        cls = MulticastExchange if kind == "nvls" else PeerExchange
        if self._exch is None or self._exch[0] is not dgraph:
            if self._exch is not None:
                self._exch[1].close()
                self._exch[2].close()
                self._exch = None
            f = self.inner.out_features
            try:
                ef = cls(dgraph.rank, dgraph.world, dgraph.pad_rows, f, dev, self.group)
                try:
                    eb = cls(dgraph.rank, dgraph.world, dgraph.pad_rows, f, dev, self.group)
                except PeerExchangeUnavailable:
                    ef.close()
                    raise
            except PeerExchangeUnavailable as e:  # agreed on by all ranks: use the NCCL exchange from now on
                sys.stderr.write("pygcn_b200.dist: %s; falling back to exchange='nccl'\n" % e)
                self.exchange = "nccl"
                return None, None
            self._exch = (dgraph, ef, eb)
        return self._exch[1], self._exch[2]

    @property
    def weight(self):
        return self.inner.weight

    @property
    def bias(self):
        return self.inner.bias

    def forward(self, input, dgraph):
        if self._ops is None:
            self._ops = CudaOps(self.inner.precision)
        if not input.is_cuda:
            raise RuntimeError("DistGraphConvolution runs on CUDA devices only (no CPU fallback)")
        ef, eb = self._exchanges(dgraph, input.device)
        return _DistGCNLayerFn.apply(input.contiguous(), self.inner.weight, self.inner.bias, dgraph,
                                     self.inner.fuse_relu, self._ops, self.group, ef, eb, self.nccl_chunks,
                                     self.inner.precision == "bf16")


# ---------------------------------------------------------------------------- bench entry (N > 1)
def bench_main(args, wl):
    """`bench.py --gpus N` under torchrun: weak scaling, N x the single-GPU CBG graph, row-partitioned."""
    import bench as B
    import pygcn_b200 as P
    from . import _lib

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    _lib.check(lib.gcnb_check_device(), "gcnb_check_device")
    sampler = B.ClockSampler(local_rank) if rank == 0 else None

    partitioned = bool(wl.get("partitioned"))
    t0 = time.perf_counter()
    if partitioned:
        # fixed total size (BASELINE configs[4]): the whole adjacency never exists on one GPU, every rank builds its
        # row block of A and A^T from the replicated edge list; the exchange is the NCCL all-gather
        n_global = wl["n"]
        exchange = "nccl"
        gen = torch.Generator(device=dev).manual_seed(0)
        src = torch.randint(0, n_global, (wl["n_raw"],), generator=gen, device=dev, dtype=torch.int32)
        dst = torch.randint(0, n_global, (wl["n_raw"],), generator=gen, device=dev, dtype=torch.int32)
        dgraph = build_partitioned(src, dst, n_global, rank, world)
        nnz_global = dgraph.nnz_global
        del src, dst
        torch.cuda.empty_cache()
        args.no_cuda_graph = True  # a captured step would pin a second 58 GB gathered panel in the graph's pool
    else:
        n_global = wl["n"] * world
        wlg = dict(wl, n=n_global)
        exchange = DistGraphConvolution.resolve_exchange(os.environ.get("GCNB_DIST_EXCHANGE", "auto"), world)
        full = B.make_graph(P, torch, wlg, dev)          # same seed on every rank: identical global graph
        dgraph = DistGraph.from_graph(full, rank, world, per_source=(exchange == "peer"))
        nnz_global = full.nnz
        del full
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    n_local = dgraph.n_rows()
    fin, fout = wl["fin"], wl["fout"]

    torch.manual_seed(42)
    layer = DistGraphConvolution(fin, fout, exchange=exchange).to(dev)
    gen = torch.Generator(device="cpu")
    x_host = torch.randn(n_local, fin, generator=gen.manual_seed(1 + rank)).pin_memory()
    g_host = torch.randn(n_local, fout, generator=gen.manual_seed(100 + rank)).pin_memory()
    x, g = x_host.to(dev), g_host.to(dev)
    flush_buf = torch.empty(B.L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def flush():
        _lib.check(lib.gcnb_l2_flush(ctypes.c_void_p(flush_buf.data_ptr()), flush_buf.numel(),
                                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "l2_flush")

    def step():
        layer.inner.weight.grad = None
        layer.inner.bias.grad = None
        out = layer(x, dgraph)
        out.backward(g)

    if sampler:
        sampler.start()
    warm = max(args.warmup, 3)
    for _ in range(warm):
        flush()
        step()
    torch.cuda.synchronize()
    dist.barrier()
    if layer._exch is None and exchange in ("peer", "nvls"):
        exchange = "nccl"  # requested peer / multicast exchange was not available (agreed on by all ranks)

    # capture the step (kernels + NCCL all-gathers / all-reduce) in a CUDA graph: removes the Python
    # and launch gaps from the device time, like the single-GPU arm.  All ranks must agree.
    cg = None
    if not args.no_cuda_graph:
        ok = torch.ones(1, device=dev)
        try:
            s_ = torch.cuda.Stream()
            s_.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_):
                for _ in range(2):
                    step()
            torch.cuda.current_stream().wait_stream(s_)
            torch.cuda.synchronize()
            layer.inner.weight.grad = None
            layer.inner.bias.grad = None
            cg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cg):
                out_static = layer(x, dgraph)
                out_static.backward(g)
            torch.cuda.synchronize()
        except Exception as e:  # pragma: no cover
            sys.stderr.write("rank %d: CUDA graph capture failed (%r); timing eager launches\n" % (rank, e))
            cg = None
            ok.zero_()
            torch.cuda.synchronize()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() == 0:
            cg = None
    eager_step = step
    if cg is not None:
        step = cg.replay
        for _ in range(warm):
            flush()
            step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if sampler:
        sampler.in_region = True
    for i in range(args.steps):
        flush()
        ev[i][0].record()
        step()
        ev[i][1].record()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.in_region = False
    total_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev)], device=dev, dtype=torch.float64)
    dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    ms_per_step = total_ms.item() / args.steps
    value = nnz_global / (ms_per_step * 1e-3)

    # end to end from pinned host buffers, per rank; max over ranks.  As in the single-GPU arm the inputs are
    # double-buffered: step i+1's H2D copies run on a copy stream while step i computes; every step still copies
    # its own X and G in and reads dW / db back inside the timed region.  No L2 flush here: one step touches the
    # row block's index stream and the gathered panel, more than the L2 holds.
    def e2e_sequential():
        total = 0.0
        for i in range(2 + args.steps):
            flush()
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            xd = x_host.to(dev, non_blocking=True)
            gd = g_host.to(dev, non_blocking=True)
            layer.inner.weight.grad = None
            layer.inner.bias.grad = None
            o = layer(xd, dgraph)
            o.backward(gd)
            layer.inner.weight.grad.cpu()
            layer.inner.bias.grad.cpu()
            torch.cuda.synchronize()
            if i >= 2:
                total += time.perf_counter() - t0
        return total

    def e2e_pipelined():
        copy_stream = torch.cuda.Stream()
        bufs = [(torch.empty_like(x), torch.empty_like(g)) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            b = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(done[b])  # the step that last read this buffer pair has finished
                bufs[b][0].copy_(x_host, non_blocking=True)
                bufs[b][1].copy_(g_host, non_blocking=True)
                ready[b].record(copy_stream)

        def loop(k):
            for b in range(2):
                done[b].record()
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            prefetch(0)
            for i in range(k):
                if i + 1 < k:
                    prefetch(i + 1)
                b = i % 2
                torch.cuda.current_stream().wait_event(ready[b])
                layer.inner.weight.grad = None
                layer.inner.bias.grad = None
                o = layer(bufs[b][0], dgraph)
                o.backward(bufs[b][1])
                done[b].record()
                layer.inner.weight.grad.cpu()  # the step's result, read back every step (synchronises)
                layer.inner.bias.grad.cpu()
            torch.cuda.synchronize()
            return time.perf_counter() - t0

        loop(3)
        return loop(args.steps)

    e2e_how = "wall clock over K eager steps of layer(x_dev, dist_graph); backward; grads.cpu(); max over ranks; "
    # (the pipelined loop was written after the round's multi-GPU minutes were spent: opt-in until it has run once)
    if partitioned or os.environ.get("GCNB_BENCH_E2E", "sequential") != "pipelined":
        e2e_s = e2e_sequential()  # (a second pair of 7 GB input buffers is not worth it at the papers size)
        e2e_how += "inputs copied from pinned host memory at the start of every step"
    else:
        try:
            e2e_s = e2e_pipelined()
            e2e_how += "inputs double-buffered from pinned host memory on a copy stream"
        except Exception as e:  # pragma: no cover  (same code and shapes on every rank: all ranks take this path together)
            sys.stderr.write("rank %d: pipelined e2e loop failed (%r); timing the sequential loop\n" % (rank, e))
            torch.cuda.synchronize()
            e2e_s = e2e_sequential()
            e2e_how += "inputs copied from pinned host memory at the start of every step"
    e2e_t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = nnz_global / (e2e_t.item() / args.steps)
    if sampler:
        sampler.stop()

    # opt-in side measurement (GCNB_BENCH_DIST_BF16=1; written after the round's multi-GPU minutes were spent): the same
    # step with bf16 panels -- half the all-gather bytes, half the gathered bytes -- as eager launches, max over ranks
    bf16_tier = None
    if os.environ.get("GCNB_BENCH_DIST_BF16") == "1" and not partitioned:
        try:
            layer16 = DistGraphConvolution(fin, fout, exchange="nccl", precision="bf16").to(dev)
            layer16.inner.load_state_dict(layer.inner.state_dict())

            def step16():
                layer16.inner.weight.grad = None
                layer16.inner.bias.grad = None
                layer16(x, dgraph).backward(g)

            for _ in range(warm):
                flush()
                step16()
            torch.cuda.synchronize()
            dist.barrier()
            ev16 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for a_, b_ in ev16:
                flush()
                a_.record()
                step16()
                b_.record()
            torch.cuda.synchronize()
            t16 = torch.tensor([sum(a_.elapsed_time(b_) for a_, b_ in ev16)], device=dev, dtype=torch.float64)
            dist.all_reduce(t16, op=dist.ReduceOp.MAX)
            eager_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for a_, b_ in eager_ev:  # the fp32 step as eager launches too, for a like-for-like ratio
                flush()
                a_.record()
                eager_step()
                b_.record()
            torch.cuda.synchronize()
            t32 = torch.tensor([sum(a_.elapsed_time(b_) for a_, b_ in eager_ev)], device=dev, dtype=torch.float64)
            dist.all_reduce(t32, op=dist.ReduceOp.MAX)
            bf16_tier = {"ms_per_step_eager": t16.item() / args.steps, "fp32_ms_per_step_eager": t32.item() / args.steps,
                         "value": nnz_global / (t16.item() / args.steps * 1e-3), "unit": "edges/s", "tolerance": 0.02,
                         "what": "DistGraphConvolution(precision='bf16'): bf16 slots all-gathered (half the bytes), "
                                 "gcnb_spmm_bf16 over the gathered panel, everything else fp32"}
        except Exception as e:  # pragma: no cover  (same code on every rank)
            sys.stderr.write("rank %d: bf16 side tier failed (%r)\n" % (rank, e))
            torch.cuda.synchronize()

    # dominant kernel on rank 0: the SpMM over its diagonal block, timed alone
    blk = dgraph.fwd_diag if dgraph.split else dgraph.fwd_remote
    ops = CudaOps()
    sup = torch.randn(blk.n_cols, fout, device=dev)
    outb = torch.empty(n_local, fout, device=dev)
    tms = []
    for it in range(3 + args.steps):
        flush()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.spmm_block(blk, sup, outb, False)
        b_.record()
        if it >= 3:
            tms.append((a, b_))
    torch.cuda.synchronize()
    spmm_ms = sum(a.elapsed_time(b_) for a, b_ in tms) / len(tms)
    alg = B.algorithmic_bytes_spmm(blk.nnz, n_local, fout)
    peak, peak_src = B.peaks()
    achieved = alg / (spmm_ms * 1e-3) / 1e9
    # our kernels per step (the claim behind "gpu_launches"; NCCL's and torch's own kernels are not counted): pack W,
    # X.W, colsum(G), dW, split-K reduce = 5, plus per SpMM (forward, transposed) the products over the row block --
    # one, two for a split block, one per phase of the peer exchange -- and the peer protocol's epoch bump, push and a
    # wait + ack per remote source
    if exchange == "peer" and dgraph.phases is not None:
        per_spmm = len(dgraph.phases) + 2 + 2 * (world - 1)
    else:
        per_spmm = 2 if dgraph.split else 1
    launches_per_step = 5 + 2 * per_spmm
    if rank == 0:
        line = {
            "metric": "gcn_layer_fwd_bwd_edges_per_sec", "value": value, "unit": "edges/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if partitioned else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"] + ((" over %d GPUs (row-partitioned by nnz, partitioned build)" % world)
                                                  if partitioned else
                                                  " x %d GPUs (N=%d, row-partitioned by nnz)" % (world, n_global)),
                       "nnz": nnz_global, "n": n_global, "in_features": fin, "out_features": fout,
                       "input_requires_grad": False, "l2": "flushed between timed steps (512 MiB write)",
                       "cuda_graph": cg is not None, "graph_build_s": build_s, "bounds": dgraph.bounds,
                       "row_block_split": dgraph.split,
                       "exchange": ("peer: own push kernel over NVLink peer memory (P2P stores + release flags), one "
                                    "column block per source rank consumed as its slot lands; NCCL all-reduce of dW,db"
                                    if exchange == "peer" else
                                    "nvls: own multimem.st push of each rank's slot to the NVLS multicast address of a "
                                    "symmetric buffer, device barriers, then the SpMM over the row block; NCCL "
                                    "all-reduce of dW,db" if exchange == "nvls" else
                                    "halo: needed-rows-only exchange (selection SpMM packs the rows each peer reads, grouped "
                                    "send / recv into a compact panel), then the SpMM over the renumbered row block; NCCL "
                                    "all-reduce of dW,db" if exchange == "halo" else
                                    "nccl: all-gather of the X.W / G panels, then the SpMM over the row block; "
                                    "all-reduce of dW,db"),
                       "nccl_chunks": layer.nccl_chunks},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": "edges/s", "h2d_bytes_per_step": (x_host.numel() + g_host.numel()) * 4,
                    "d2h_bytes_per_step": (fin * fout + fout) * 4, "ms_per_step": e2e_t.item() / args.steps * 1e3,
                    "how": e2e_how},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "spmm_group_kernel<LPR=8,U=4,24 CTAs/SM,W=2,SE=16> on rank 0's %s" % ("diagonal block" if dgraph.split else "row block (all-gathered panel)"),
                         "algorithmic_bytes_per_launch": alg, "kernel_ms": spmm_ms, "peak_source": peak_src},
        }
        if bf16_tier is not None:
            line["bf16_tier"] = bf16_tier
        print(json.dumps(line), flush=True)
    # a captured graph that holds NCCL kernels must go before the communicator does
    del step, cg
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)  # skip communicator teardown: it can block behind graph-captured collectives
