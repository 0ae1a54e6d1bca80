"""CPU-side checks: the C-ABI library builds, loads and exports every symbol the header
declares; the host-side mirror of the reference interface behaves like the reference's
(init draws, repr, state-dict keys, pickling under the module path `layers`, error
behaviour on CPU tensors).  No kernel is launched here."""
import io
import os
import pickle
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from pygcn_b200 import _lib

    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    from pygcn_b200 import _lib

    header = open(os.path.join(ROOT, "include", "gcnb200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(gcnb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(lib, name), "libgcnb200.so does not export %s" % name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert lib.gcnb_version() == int(re.search(r"#define GCNB_VERSION (\d+)", header).group(1))


def test_library_is_sm100a_with_lineinfo():
    from pygcn_b200 import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_header_constants_match_binding():
    from pygcn_b200 import _lib

    header = open(os.path.join(ROOT, "include", "gcnb200.h")).read()
    consts = dict(re.findall(r"#define (GCNB_[A-Z0-9_]+) (\d+)", header))
    assert int(consts["GCNB_NUM_BINS"]) == _lib.NUM_BINS
    assert tuple(int(consts["GCNB_BIN_EDGE_%d" % i]) for i in range(1, 5)) == _lib.BIN_EDGES[1:]
    assert int(consts["GCNB_SPMM_TRANSPOSE"]) == _lib.SPMM_TRANSPOSE and int(consts["GCNB_SPMM_RELU"]) == _lib.SPMM_RELU
    assert int(consts["GCNB_LAYER_NEED_DB"]) == _lib.LAYER_NEED_DB
    # the oracle's schedule bins use the same edges
    from oracle import gcn_oracle as O
    import inspect

    assert inspect.signature(O.degree_bins).parameters["edges"].default == _lib.BIN_EDGES


def test_tuning_knobs_validate_their_values(lib):
    """gcnb_set_tuning touches host state only: valid values are accepted, anything else is refused with a message
    (the defaults -- auto kernel, auto variant, plain stream order -- are restored)."""
    from pygcn_b200 import _lib

    header = open(os.path.join(ROOT, "include", "gcnb200.h")).read()
    consts = dict(re.findall(r"#define (GCNB_TUNE_[A-Z0-9_]+) (\d+)", header))
    assert (int(consts["GCNB_TUNE_SPMM_KERNEL"]), int(consts["GCNB_TUNE_SPMM_GROUP_VARIANT"]), int(consts["GCNB_TUNE_PDL"])) == (
        _lib.TUNE_SPMM_KERNEL, _lib.TUNE_SPMM_GROUP_VARIANT, _lib.TUNE_PDL)
    try:
        for key, good, bad in ((_lib.TUNE_SPMM_KERNEL, (0, 1, 2, 3), (-1, 4)),
                               (_lib.TUNE_SPMM_GROUP_VARIANT, tuple(range(-1, 16)), (-2, 16)),
                               (_lib.TUNE_PDL, (0, 1), (-1, 2)),
                               (_lib.TUNE_SPMM_STREAM, (0, 1, 2), (-1, 3)),
                               (_lib.TUNE_STREAM_HOT_MB, (0, 48, 126), (-1, 127)),
                               (_lib.TUNE_STREAM_HINT, (0, 1, 2), (-1, 3)),
                               (_lib.TUNE_STREAM_MIN_ROW_BYTES, (16, 256, 1024), (8, 2048)),
                               (_lib.TUNE_STREAM_BATCH, (0, 2, 4, 8), (3, 16))):
            for v in good:
                assert lib.gcnb_set_tuning(key, v) == 0, (key, v, _lib.last_error())
            for v in bad:
                assert lib.gcnb_set_tuning(key, v) != 0 and "set_tuning" in _lib.last_error()
        assert lib.gcnb_set_tuning(99, 0) != 0 and "unknown key" in _lib.last_error()
    finally:
        lib.gcnb_set_tuning(_lib.TUNE_SPMM_KERNEL, 0)
        lib.gcnb_set_tuning(_lib.TUNE_SPMM_GROUP_VARIANT, -1)
        lib.gcnb_set_tuning(_lib.TUNE_PDL, 0)
        lib.gcnb_set_tuning(_lib.TUNE_SPMM_STREAM, 1)
        lib.gcnb_set_tuning(_lib.TUNE_STREAM_HOT_MB, 48)
        lib.gcnb_set_tuning(_lib.TUNE_STREAM_HINT, 0)
        lib.gcnb_set_tuning(_lib.TUNE_STREAM_MIN_ROW_BYTES, 256)
        lib.gcnb_set_tuning(_lib.TUNE_STREAM_BATCH, 0)


def test_halo_plan_entry_points_validate_on_the_host(lib):
    """gcnb_halo_create (the exchange step of the row-partitioned layer, csrc/dist.cu): argument checks run before any
    CUDA call, an empty plan needs no device memory, NCCL is bound at run time (the library has no link-time
    dependency on it: `ldd` must not list libnccl)."""
    import ctypes
    import subprocess

    from pygcn_b200 import _lib

    i64 = ctypes.c_int64 * 3
    h = ctypes.c_void_p()
    assert lib.gcnb_halo_create(1, 3, i64(0, 0, 0), None, i64(0, 0, 0), i64(0, 0, 0), None, ctypes.byref(h)) == 0
    assert lib.gcnb_halo_send_rows(h) == 0 and lib.gcnb_halo_recv_rows(h) == 0
    assert lib.gcnb_halo_pack(h, None, 8, 8, None, None) == 0  # nothing to send: a no-op
    lib.gcnb_halo_free(h)
    for args in ((3, 3, i64(0, 0, 0)), (0, 3, i64(5, 0, 0)), (0, 3, i64(0, -1, 0))):  # bad rank; rows to itself; negative
        assert lib.gcnb_halo_create(args[0], args[1], args[2], None, i64(0, 0, 0), i64(0, 0, 0), None, ctypes.byref(h)) != 0
        assert "halo_create" in _lib.last_error()
    assert lib.gcnb_halo_exchange(None, None, None, 8, None, None) != 0 and "halo_exchange" in _lib.last_error()
    assert lib.gcnb_halo_nccl_available() in (0, 1)
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "libnccl" not in out


def test_fresh_bn_entry_points_validate_before_launching(lib):
    """gcnb_fresh_bn_*: every argument check runs on the host before any kernel is launched, so refusals can be
    exercised without a GPU -- shapes, null operands, leading dimensions, workspace size and alignment; an empty batch
    is a no-op; the workspace size is the per-CTA fp64 partials [blocks][2][f]."""
    import ctypes

    from pygcn_b200 import _lib

    blocks = lambda n: max(1, min(-(-n // 128), 4 * 148))
    for n, f in ((1, 1), (100000, 32), (257, 7), (10 ** 7, 256)):
        assert lib.gcnb_fresh_bn_workspace_bytes(n, f) == blocks(n) * 2 * f * 8 + 16
    assert lib.gcnb_fresh_bn_workspace_bytes(0, 32) == 16
    buf = (ctypes.c_float * 4096)()
    a = ctypes.addressof(buf)
    fwd = lambda n, f, y, ldy, out, ldo, mean, rstd, ws, wsb, eps=1e-5: lib.gcnb_fresh_bn_forward(
        n, f, y, ldy, 1, eps, out, ldo, mean, rstd, ws, wsb, None)
    bwd = lambda n, f, y, ldy, g, ldg, mean, rstd, dy, lddy, gstat, ws, wsb: lib.gcnb_fresh_bn_backward(
        n, f, y, ldy, 1, g, ldg, mean, rstd, dy, lddy, gstat, ws, wsb, None)
    assert fwd(0, 8, None, 8, None, 8, None, None, None, 0) == 0             # empty batch: nothing to do
    assert bwd(0, 8, None, 8, None, 8, None, None, None, 8, None, None, 0) == 0
    for call, msg in ((lambda: fwd(4, 0, a, 8, a, 8, a, a, a, 4096), "shape out of range"),
                    (lambda: fwd(-1, 8, a, 8, a, 8, a, a, a, 4096), "shape out of range"),
                    (lambda: fwd(4, 8, a, 8, a, 8, a, a, a, 4096, eps=-1.0), "eps"),
                    (lambda: fwd(4, 8, None, 8, a, 8, a, a, a, 4096), "null operand"),
                    (lambda: fwd(4, 8, a, 7, a, 8, a, a, a, 4096), "leading dimension"),
                    (lambda: fwd(4, 8, a, 8, a, 8, a, a, a, 8), "workspace too small"),
                    (lambda: fwd(4, 8, a, 8, a, 8, a, a, a + 4, 4096), "misaligned"),
                    (lambda: bwd(4, 8, a, 8, None, 8, a, a, a, 8, a, a, 4096), "null operand"),
                    (lambda: bwd(4, 8, a, 8, a, 8, a, a, a, 7, a, a, 4096), "leading dimension"),
                    (lambda: bwd(4, 8, a, 8, a, 8, a, a, a, 8, a, a, 8), "workspace too small")):
        st = call()
        assert st == 1 and msg in _lib.last_error(), (st, msg, _lib.last_error())


def test_init_draws_match_reference(golden):
    import pygcn_b200 as P

    g = golden("init.npz")
    torch.manual_seed(42)
    gc = P.GraphConvolution(64, 32)
    assert np.array_equal(gc.weight.detach().numpy(), g["w_64_32"])
    assert np.array_equal(gc.bias.detach().numpy(), g["b_64_32"])
    assert repr(gc) == str(g["repr"])
    assert sorted(gc.state_dict().keys()) == list(g["state_keys"])
    torch.manual_seed(42)
    stack = [P.GraphConvolution(8, 32), P.GraphConvolution(32, 32), P.GraphConvolution(32, 32, bias=False)]
    for i, gg in enumerate(stack):
        assert np.array_equal(gg.weight.detach().numpy(), g["stack_w%d" % i])
        if gg.bias is not None:
            assert np.array_equal(gg.bias.detach().numpy(), g["stack_b%d" % i])
    assert stack[2].bias is None and "bias" in dict(stack[2].named_parameters(recurse=False)) or stack[2].bias is None


def test_cpu_tensors_raise_not_fallback():
    import pygcn_b200 as P

    gc = P.GraphConvolution(4, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        gc(torch.zeros(5, 4), torch.eye(5))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.spmm(torch.eye(5).to_sparse(), torch.zeros(5, 2))
    with pytest.raises(TypeError):
        P.spmm([[1.0]], torch.zeros(1, 1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.apply_bn(torch.zeros(5, 4), relu=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.load_adj(torch.zeros(5, 4))


def test_whole_model_pickle_under_reference_module_path():
    import pygcn_b200 as P

    mod = P.install_as_pygcn()
    assert sys.modules["layers"] is mod
    import layers  # what pygcn/models.py:4 does

    gc = layers.GraphConvolution(5, 2)
    buf = io.BytesIO()
    torch.save(gc, buf)  # gnn-over-mlp.py:489 saves whole modules
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    assert isinstance(back, P.GraphConvolution) and torch.equal(back.weight, gc.weight)
    # a state dict from the reference layer loads (same keys / shapes)
    sd = {"weight": torch.ones(5, 2), "bias": torch.zeros(2)}
    gc.load_state_dict(sd)
    assert pickle.loads(pickle.dumps(gc)).in_features == 5


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from pygcn_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.GcnbError, match="no CPU/PyTorch fallback"):
        _lib.load(build_if_missing=False)


def test_apply_bn_batched_layout_plumbing(monkeypatch):
    """apply_bn(x[B, N, F]): the host-side plumbing (node-major view, one 2-D call on [N, B*F], the permuted view back,
    autograd through the views) with a torch stand-in for the CUDA op -- the result is the per-sample loop of
    GCN_OVER_MLP.forward (pygcn/models.py:343-349), for batch-major storage and for the batched layer's own layout."""
    import pygcn_b200.functional as Fn

    calls = []

    class StandIn:
        @staticmethod
        def apply(x2, relu, eps):
            calls.append(tuple(x2.shape))
            assert x2.dim() == 2 and x2.is_contiguous()
            a = torch.relu(x2) if relu else x2
            return (a - a.mean(0)) / torch.sqrt(a.var(0, unbiased=False) + eps)

    monkeypatch.setattr(Fn, "_require_cuda", lambda t, what: None)
    monkeypatch.setattr(Fn, "_FreshBatchNormFn", StandIn)
    gen = torch.Generator().manual_seed(3)
    b, n, f = 4, 50, 6
    x = torch.randn(b, n, f, generator=gen, requires_grad=True)
    g = torch.randn(b, n, f, generator=gen)
    out = Fn.apply_bn(x, relu=True)
    out.backward(g)
    got_grad = x.grad.clone()
    x.grad = None
    want = torch.stack([torch.nn.BatchNorm1d(f)(torch.relu(x[i])) for i in range(b)])
    want.backward(g)
    assert out.shape == (b, n, f) and calls == [(n, b * f)]
    assert torch.allclose(out, want, atol=1e-5) and torch.allclose(got_grad, x.grad, atol=1e-5)
    node_major = torch.randn(n, b, f, generator=gen).permute(1, 0, 2)  # what the batched layer returns: a free view
    assert torch.allclose(Fn.apply_bn(node_major), torch.stack([torch.nn.BatchNorm1d(f)(node_major[i]) for i in range(b)]),
                          atol=1e-5)
    with pytest.raises(ValueError):
        Fn.apply_bn(torch.zeros(2, 1, 3))       # one node per sample: torch's training-mode error
    with pytest.raises(ValueError):
        Fn.apply_bn(torch.zeros(2, 2, 3, 4))
    assert Fn.apply_bn(torch.zeros(0, 5, 3)).shape == (0, 5, 3)
