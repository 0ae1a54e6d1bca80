"""Thread-by-thread emulation, on the CPU, of the (ReLU ->) fresh-BatchNorm kernels (pygcn_b200/csrc/batchnorm.cu):
they were written after round 1's GPU minutes were spent, so the index walk -- CTA row ranges, cw column-group lanes x rl
row lanes, the four-rows-in-flight loop and its tail, column tiles for wide panels, the per-CTA partial layout, the
warp-per-column finalise (lane p takes parts p, p + 32, ...), the grid-stride apply pass -- is replayed here statement
by statement with numpy and checked against the oracle (oracle/gcn_oracle.py::fresh_batchnorm_*) and against the
reference-generated fixture tests/golden/apply_bn.npz.  Every (row, column) must be visited exactly once per pass."""
import numpy as np
import pytest

from oracle import gcn_oracle as O

K_THREADS = 256
K_MAX_BLOCKS = 4 * 148


def bn_blocks(n):
    return max(1, min(-(-n // 128), K_MAX_BLOCKS))


def lanes_for(fv):
    cw = 1
    while cw < fv and cw < K_THREADS:
        cw <<= 1
    return cw


def stats_pass(mode, V, y, relu, g=None, mean=None, rstd=None):
    """bn_stats_kernel<MODE, V> + bn_finalize_kernel<MODE> sums (before the division): returns (S1, S2, visits)."""
    n, f = y.shape
    fv = f // V
    nb = bn_blocks(n)
    rpb = -(-n // nb)
    cw = lanes_for(fv)
    rl = K_THREADS // cw
    partial = np.zeros((nb, 2, f), dtype=np.float64)
    visits = np.zeros((n, f), dtype=np.int64)
    for b in range(nb):
        r0, r1 = b * rpb, min(b * rpb + rpb, n)
        for j0 in range(0, fv, cw):
            red = np.zeros((2, V, K_THREADS), dtype=np.float64)
            for tid in range(K_THREADS):
                tx, ty = tid % cw, tid // cw
                j = j0 + tx
                if j >= fv:
                    continue
                s1 = np.zeros(V)
                s2 = np.zeros(V)

                def add(r):
                    for t in range(V):
                        c = j * V + t
                        visits[r, c] += 1
                        a = np.float32(max(y[r, c], 0.0) if relu else y[r, c])
                        if mode == 0:
                            s1[t] += float(a)
                            s2[t] += float(a) * float(a)
                        else:
                            xh = np.float32(np.float32(a - mean[c]) * rstd[c])
                            s1[t] += float(g[r, c])
                            s2[t] += float(g[r, c]) * float(xh)

                r = r0 + ty
                while r + 3 * rl < r1:
                    for u in range(4):
                        add(r + u * rl)
                    r += 4 * rl
                while r < r1:
                    add(r)
                    r += rl
                red[0, :, tid] = s1
                red[1, :, tid] = s2
            for tx in range(cw):  # ty == 0 threads
                j = j0 + tx
                if j >= fv:
                    continue
                for t in range(V):
                    for k in range(rl):
                        partial[b, 0, j * V + t] += red[0, t, k * cw + tx]
                        partial[b, 1, j * V + t] += red[1, t, k * cw + tx]
    s = np.zeros((2, f))
    for c in range(f):  # one warp per column
        lane_sums = np.zeros((2, 32))
        for lane in range(32):
            for p in range(lane, nb, 32):
                lane_sums[:, lane] += partial[p, :, c]
        v = lane_sums.copy()
        o = 16
        while o > 0:  # xor tree
            v = v + v[:, np.arange(32) ^ o]
            o >>= 1
        s[:, c] = v[:, 0]
    return s[0], s[1], visits


def apply_pass(mode, V, y, relu, mean, rstd, g=None, gbar=None, gxbar=None):
    n, f = y.shape
    fv = f // V
    cw = lanes_for(fv)
    rl = K_THREADS // cw
    grid = max(1, min(-(-n // (rl * 4)), 16 * 148))
    out = np.full((n, f), np.nan, dtype=np.float32)
    visits = np.zeros((n, f), dtype=np.int64)
    for b in range(grid):
        for tid in range(K_THREADS):
            tx, ty = tid % cw, tid // cw
            for j0 in range(0, fv, cw):
                j = j0 + tx
                if j >= fv:
                    continue
                r = b * rl + ty
                while r < n:
                    for t in range(V):
                        c = j * V + t
                        visits[r, c] += 1
                        a = np.float32(max(y[r, c], 0.0) if relu else y[r, c])
                        xh = np.float32(np.float32(a - mean[c]) * rstd[c])
                        if mode == 0:
                            out[r, c] = xh
                        else:
                            da = np.float32(rstd[c] * np.float32(np.float32(g[r, c] - gbar[c]) - np.float32(xh * gxbar[c])))
                            out[r, c] = np.float32(0.0) if (relu and not y[r, c] > 0) else da
                    r += grid * rl
    return out, visits


def emulate(y, g, relu, V, eps=1e-5):
    n = y.shape[0]
    s1, s2, v1 = stats_pass(0, V, y, relu)
    mu = s1 / n
    var = np.maximum(s2 / n - mu * mu, 0.0)
    mean = mu.astype(np.float32)
    rstd = (1.0 / np.sqrt(var + eps)).astype(np.float32)
    out, v2 = apply_pass(0, V, y, relu, mean, rstd)
    t1, t2, v3 = stats_pass(1, V, y, relu, g, mean, rstd)
    gbar, gxbar = (t1 / n).astype(np.float32), (t2 / n).astype(np.float32)
    dy, v4 = apply_pass(1, V, y, relu, mean, rstd, g, gbar, gxbar)
    for v in (v1, v2, v3, v4):
        assert (v == 1).all(), "an element was skipped or visited twice"
    return out, dy


@pytest.mark.parametrize("n,f,V,relu", [(300, 8, 4, True), (300, 8, 1, True), (257, 7, 1, True), (2, 4, 4, False),
                                        (131, 12, 4, False), (1, 5, 1, True), (700, 20, 4, True)])
def test_emulated_kernels_match_the_oracle(n, f, V, relu):
    rs = np.random.default_rng(n * 31 + f)
    y = rs.standard_normal((n, f), dtype=np.float32) + np.float32(0.3)
    g = rs.standard_normal((n, f), dtype=np.float32)
    out, dy = emulate(y, g, relu, V)
    ref_out, _, _ = O.fresh_batchnorm_forward(y, relu)
    ref_dy = O.fresh_batchnorm_backward(y, g, relu)
    assert O.normwise_err(out, ref_out) < 1e-5
    if n > 2:  # (1- and 2-row batches differentiate through a cancellation; the forward is still exact)
        assert O.normwise_err(dy, ref_dy) < 1e-5


def test_wide_panel_walks_column_tiles():
    """More column groups than the CTA has lanes (f = 1040 scalar columns > 256): the j0 tile loop."""
    rs = np.random.default_rng(5)
    y = rs.standard_normal((9, 1040), dtype=np.float32)
    s1, s2, visits = stats_pass(0, 1, y, True)
    assert (visits == 1).all()
    a = np.maximum(y.astype(np.float64), 0)
    assert np.allclose(s1, a.sum(0), rtol=1e-12, atol=1e-12) and np.allclose(s2, (a * a).sum(0), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name,V", [("deadcol", 4), ("odd7", 1)])
def test_emulated_kernels_match_the_reference_fixture(golden, name, V):
    """Against the reference's own apply_bn(F.relu(y)) outputs (tests/golden/apply_bn.npz, pygcn/models.py:41-45)."""
    c = golden("apply_bn.npz")
    y, g, relu = c[name + "/y"], c[name + "/g"], bool(c[name + "/relu"])
    out, dy = emulate(y, g, relu, V)
    assert O.normwise_err(out, c[name + "/out"]) < 1e-5
    assert O.normwise_err(dy, c[name + "/dy"]) < 1e-5
    if name == "deadcol":
        assert not out[:, 3].any() and not dy[:, 3].any()
