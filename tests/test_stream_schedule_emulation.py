"""CPU replay of the entry-balanced schedule of the streaming SpMM (pygcn_b200/csrc/spmm_stream.cu +
graph_build.cu: stream_items_kernel, tag_row_end_kernel, tag_class_kernel, spmm_stream_fixup_kernel): the items,
the end-of-row tags, the head / tail partial rows and the ordered fix-up, step by step as the kernels do them, against
a plain CSR product.  The CUDA kernels themselves are checked on the GPU (tests/test_gpu_parity.py); this pins the
index logic -- rows that span many items, rows that start exactly on an item boundary, single-entry rows -- where no
GPU is available."""
import numpy as np
import pytest

ITEM = 1024          # kStreamItem
ROW_END = 1 << 31    # kPairRowEnd
COL_MASK = (1 << 27) - 1


def build_tags(rowptr, col, n_cols):
    """graph_build.cu::build_stream_schedule on the host."""
    nnz = len(col)
    cnt = np.bincount(col, minlength=n_cols)
    srt = np.sort(cnt)[::-1]
    th = [int(srt[1024 << c]) if (1024 << c) < n_cols else -1 for c in range(15)]
    cls = np.full(n_cols, 15, np.uint32)
    for j in range(14, -1, -1):
        cls[cnt > th[j]] = j
    word = col.astype(np.uint32) | (cls[col] << np.uint32(27))
    ends = rowptr[1:][rowptr[1:] > rowptr[:-1]] - 1
    word[ends] |= np.uint32(ROW_END)
    n_items = (nnz + ITEM - 1) // ITEM
    items = np.zeros(n_items, np.uint32)
    for i in range(n_items):
        e = i * ITEM
        r = int(np.searchsorted(rowptr, e, side="right")) - 1  # last r with rowptr[r] <= e
        r = min(r, len(rowptr) - 2)
        items[i] = r | (ROW_END if rowptr[r] < e else 0)
    return word, items, cls, cnt


def stream_spmm(rowptr, word, val, items, dense, n_warps=7):
    """spmm_stream_kernel + spmm_stream_fixup_kernel, warp by warp, entry by entry."""
    n, f = len(rowptr) - 1, dense.shape[1]
    nnz = len(word)
    n_items = len(items)
    out = np.full((n, f), np.nan, np.float64)
    partial = np.full((2 * n_items, f), np.nan, np.float64)
    for gw in range(n_warps):
        for item in range(gw, n_items, n_warps):
            row = int(items[item] & 0x7FFFFFFF)
            head = bool(items[item] >> 31)
            acc = np.zeros(f)
            opened = False
            for e in range(item * ITEM, min((item + 1) * ITEM, nnz)):
                acc += val[e] * dense[word[e] & COL_MASK]
                opened = True
                if word[e] & ROW_END:
                    if head:
                        partial[2 * item] = acc
                    else:
                        assert np.isnan(out[row]).all(), "row written twice"
                        out[row] = acc
                    head, opened = False, False
                    row += 1
                    acc = np.zeros(f)
            if opened:
                partial[2 * item + (0 if head else 1)] = acc
    for bnd in range(1, n_items):
        if not (items[bnd] >> 31):
            continue
        row = int(items[bnd] & 0x7FFFFFFF)
        if rowptr[row] < (bnd - 1) * ITEM:
            continue
        last = (rowptr[row + 1] - 1) // ITEM
        acc = partial[2 * (bnd - 1) + 1].copy()
        for i in range(bnd, last + 1):
            acc += partial[2 * i]
        assert np.isnan(out[row]).all(), "row written twice"
        out[row] = acc
    return out


def reference(rowptr, col, val, dense):
    n = len(rowptr) - 1
    out = np.zeros((n, dense.shape[1]))
    for r in range(n):
        s, e = rowptr[r], rowptr[r + 1]
        out[r] = (val[s:e, None] * dense[col[s:e]]).sum(0)
    return out


def make_csr(lengths, n_cols, rs):
    rowptr = np.concatenate([[0], np.cumsum(lengths)]).astype(np.int64)
    col = rs.integers(0, n_cols, rowptr[-1])
    val = rs.standard_normal(rowptr[-1])
    return rowptr, col, val


@pytest.mark.parametrize("case", ["power_law", "exact_boundaries", "one_hub", "ones", "tiny"])
def test_stream_schedule_matches_csr_product(case):
    rs = np.random.default_rng(11)
    if case == "power_law":
        lengths = np.maximum(1, (rs.pareto(1.1, 600) * 6).astype(np.int64))
        lengths[17] = 5000
        lengths[300] = 2048
    elif case == "exact_boundaries":  # rows starting and ending exactly on item boundaries, a whole-item row
        lengths = np.array([1024, 1024, 512, 512, 2048, 1, 1023, 3072, 5])
    elif case == "one_hub":
        lengths = np.array([3, 7000, 2])
    elif case == "ones":
        lengths = np.ones(2500, np.int64)
    else:
        lengths = np.array([4, 1, 9])
    n_cols = 3000
    rowptr, col, val = make_csr(lengths, n_cols, rs)
    word, items, cls, cnt = build_tags(rowptr, col, n_cols)
    assert ((word & COL_MASK) == col).all()
    dense = rs.standard_normal((n_cols, 5))
    out = stream_spmm(rowptr, word, val, items, dense)
    ref = reference(rowptr, col, val, dense)
    assert not np.isnan(out).any(), "a row was never written"
    np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12)


def test_heat_classes_are_nested_top_k_sets():
    rs = np.random.default_rng(5)
    n_cols = 40_000
    col = np.minimum((rs.pareto(0.9, 400_000) * 3).astype(np.int64), n_cols - 1)
    rowptr = np.arange(0, 400_001, 4)
    _, _, cls, cnt = build_tags(rowptr, col, n_cols)
    for c in range(15):
        hot = cls <= c
        k = 1024 << c
        assert hot.sum() <= max(k, 0) or k >= n_cols
        if hot.any() and (~hot).any():  # every hot column is referenced more often than every other column
            assert cnt[hot].min() > cnt[~hot].max() - 1e-9 or cnt[hot].min() >= cnt[~hot].max()
    assert (cls[cnt == 0] >= cls[cnt == cnt.max()].max()).all()
