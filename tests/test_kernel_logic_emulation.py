"""Lock-step emulation, on the CPU, of the staging / prefetch index logic of spmm_group_persistent_kernel
(pygcn_b200/csrc/spmm.cu, tuning variants 16 / 17): the kernel was written after round 1's GPU minutes were spent, so
its control flow -- even-aligned stage windows, the zeroed leading pair of an odd row start, zero-filled tails,
sit-out of finished groups, the next row set's first stage landing in the free buffer during the last stage iteration,
empty and long (skipped) rows, several sets per CTA -- is checked here statement by statement against a CSR reference.
cp.async is modelled as an immediate copy with zero fill (the wait-group accounting is argued in the kernel's
comments); a gather is b[col] * val.  The plain group kernel is the same code without the cross-set prefetch."""
import random

def run(n_rows, rowptr, cols, vals, b, LPR, U, W, SE_E, grid, skip_long=False, LONG=1025):
    G = 32 // LPR
    E = SE_E
    NC = E // (2 * LPR)
    assert E % U == 0 and NC >= 1
    nnz = rowptr[-1]
    last_pair = nnz & ~1
    pair = [(cols[i], vals[i]) for i in range(nnz)] + [(0, 0.0)] * 4  # array padded like the real one
    n_sets = (n_rows + W * G - 1) // (W * G)
    out = [None] * n_rows
    for block in range(grid):
        for warp in range(W):
            # per-lane state
            lanes = range(32)
            sub = [l % LPR for l in lanes]
            grp = [l // LPR for l in lanes]
            stage = [[[(0, 0.0)] * (E + 2) for _ in range(2)] for _ in range(G)]  # [group][buf][entry]
            cur = [0] * 32; nxt = [1] * 32
            def row_of(s, l): return (s * W + warp) * G + grp[l]
            def load_row(s, l):
                st = en = 0
                if s < n_sets:
                    r = row_of(s, l)
                    if r < n_rows:
                        st, en = rowptr[r], rowptr[r + 1]
                        if skip_long and en - st >= LONG: en = st
                return st, en
            def fetch_at(l, buf, frm, ahead):
                for k in range(NC):
                    a = ahead - 2 * LPR * k
                    nbytes = 16 if a >= 2 else (8 if a == 1 else 0)
                    src = min(max(frm + 2 * LPR * k, 0), last_pair)
                    dst = 2 * (sub[l] + LPR * k)   # entry index (16 bytes = 2 pairs)
                    p0 = pair[src] if nbytes >= 8 else (0, 0.0)
                    p1 = pair[src + 1] if nbytes >= 16 else (0, 0.0)
                    stage[grp[l]][buf][dst] = p0
                    stage[grp[l]][buf][dst + 1] = p1
            s = block
            e = [0] * 32; left = [0] * 32; lead = [0] * 32
            for l in lanes:
                st, en = load_row(s, l)
                lead[l] = st & 1
                e[l] = st - lead[l] + 2 * sub[l]
                left[l] = 0 if en == st else en - st + lead[l]
            for l in lanes:
                fetch_at(l, cur[l], e[l], left[l] - 2 * sub[l]); e[l] += E
            nrow = [load_row(s + grid, l) for l in lanes]
            while True:
                acc = [0.0] * 32   # per lane: one scalar feature chunk (b is a vector of scalars per column)
                first = True
                prefetched = False
                while not all(left[l] <= 0 for l in lanes):
                    if not all(left[l] - E <= 0 for l in lanes):
                        for l in lanes:
                            fetch_at(l, nxt[l], e[l], left[l] - E - 2 * sub[l]); e[l] += E
                    else:
                        for l in lanes:
                            nst, nen = nrow[l]
                            nlead = nst & 1
                            nleft = 0 if nen == nst else nen - nst + nlead
                            fetch_at(l, nxt[l], nst - nlead + 2 * sub[l], nleft - 2 * sub[l])
                        prefetched = True
                    if first:
                        for l in lanes:
                            if lead[l] and sub[l] == 0:
                                c, v = stage[grp[l]][cur[l]][0]
                                stage[grp[l]][cur[l]][0] = (c, 0.0)
                        first = False
                    cnt = [min(max(left[l], 0), E) for l in lanes]
                    cmax = max(cnt)
                    j = 0
                    while j < cmax:
                        for l in lanes:
                            if j < cnt[l]:
                                for u in range(U):
                                    c, v = stage[grp[l]][cur[l]][j + u]
                                    acc[l] += v * b[c]
                        j += U
                    for l in lanes:
                        left[l] -= E
                        cur[l], nxt[l] = nxt[l], cur[l]
                for l in lanes:
                    r = row_of(s, l)
                    store = r < n_rows and sub[l] == 0
                    if store and skip_long and rowptr[r + 1] - rowptr[r] >= LONG: store = False
                    if store:
                        assert out[r] is None, "row stored twice"
                        out[r] = acc[l]
                s += grid
                if s >= n_sets: break
                for l in lanes:
                    nst, nen = nrow[l]
                    lead[l] = nst & 1
                    e[l] = nst - lead[l] + 2 * sub[l]
                    left[l] = 0 if nen == nst else nen - nst + lead[l]
                if not prefetched:
                    for l in lanes: fetch_at(l, cur[l], e[l], left[l] - 2 * sub[l])
                for l in lanes: e[l] += E
                nrow = [load_row(s + grid, l) for l in lanes]
    return out

def test_persistent_group_kernel_index_logic_matches_csr_reference():
    rnd = random.Random(1)
    for trial in range(200):
        LPR = rnd.choice([2, 4, 8, 16])
        U = rnd.choice([2, 4, 8])
        W = rnd.choice([1, 2, 4])
        E = rnd.choice([16, 32])
        if E % U or E // (2 * LPR) < 1: continue
        n_rows = rnd.randint(1, 120)
        n_cols = rnd.randint(1, 50)
        rowptr = [0]
        for r in range(n_rows):
            kind = rnd.random()
            d = 0 if kind < 0.2 else (rnd.randint(1, 5) if kind < 0.5 else (rnd.randint(6, 70) if kind < 0.95 else rnd.randint(100, 140)))
            rowptr.append(rowptr[-1] + d)
        nnz = rowptr[-1]
        cols = [rnd.randrange(n_cols) for _ in range(nnz)]
        vals = [rnd.uniform(-1, 1) for _ in range(nnz)]
        b = [rnd.uniform(-1, 1) for _ in range(n_cols)]
        G = 32 // LPR
        n_sets = (n_rows + W * G - 1) // (W * G)
        grid = rnd.randint(1, max(1, n_sets))
        skip = rnd.random() < 0.3
        out = run(n_rows, rowptr, cols, vals, b, LPR, U, W, E, grid, skip_long=skip, LONG=100)
        for r in range(n_rows):
            d = rowptr[r + 1] - rowptr[r]
            if skip and d >= 100:
                assert out[r] is None, (trial, r)
                continue
            want = sum(vals[i] * b[cols[i]] for i in range(rowptr[r], rowptr[r + 1]))
            assert out[r] is not None and abs(out[r] - want) < 1e-9, (trial, r, out[r], want, LPR, U, W, E, grid, d, rowptr[r] & 1)
