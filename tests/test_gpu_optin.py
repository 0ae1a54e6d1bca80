"""GPU tests of the OPT-IN paths that were written after round 1's GPU minutes were spent and have therefore never
run on hardware: they are skipped unless GCNB_TEST_UNVERIFIED=1, so the default `pytest -m gpu` run covers exactly
what has been measured.  First thing to do with a GPU: `GCNB_TEST_UNVERIFIED=1 pytest tests/test_gpu_optin.py -m gpu`.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("GCNB_TEST_UNVERIFIED") != "1",
                                 reason="opt-in paths not yet run on hardware (set GCNB_TEST_UNVERIFIED=1)")]


def dev():
    return torch.device("cuda:0")


@pytest.fixture()
def pdl():
    """Switches programmatic dependent launch on for the test body (csrc/common.cuh), restores plain stream order."""
    from pygcn_b200 import _lib

    lib = _lib.load()

    def switch(on):
        _lib.check(lib.gcnb_set_tuning(_lib.TUNE_PDL, 1 if on else 0), "set_tuning")
    yield switch
    lib.gcnb_set_tuning(_lib.TUNE_PDL, 0)


def _graph(P, n, n_edges, seed):
    rs = np.random.default_rng(seed)
    src = torch.from_numpy(rs.integers(0, n, n_edges).astype(np.int32)).to(dev())
    dst = torch.from_numpy(rs.integers(0, n, n_edges).astype(np.int32)).to(dev())
    return P.Graph.from_edges(src, dst, n)


@pytest.mark.parametrize("n,fin,fout,relu", [(20000, 64, 32, False), (20000, 64, 32, True), (20000, 16, 7, True),
                                             (30000, 100, 256, False), (6000, 8, 32, True)])
def test_pdl_chain_is_bit_identical_to_plain_launches(pdl, n, fin, fout, relu):
    """The same kernels in the same order: with the PDL instantiations (launch_dependents first, wait before the first
    dependent access) out, dX, dW and db must equal the plain-launch results bit for bit, eagerly and as a CUDA-graph
    replay, over several back-to-back steps (the overlap only exists between consecutive kernels)."""
    import pygcn_b200 as P

    gr = _graph(P, n, 20 * n, seed=fout)
    gen = torch.Generator(device=dev()).manual_seed(fin)
    x = torch.randn(n, fin, generator=gen, device=dev())
    g = torch.randn(n, fout, generator=gen, device=dev())
    torch.manual_seed(42)
    layer = P.GraphConvolution(fin, fout, fuse_relu=relu).to(dev())

    def step():
        layer.weight.grad = None
        layer.bias.grad = None
        xt = x.clone().requires_grad_(True)
        out = layer(xt, gr)
        out.backward(g)
        return out.detach().clone(), xt.grad.clone(), layer.weight.grad.clone(), layer.bias.grad.clone()

    pdl(False)
    want = step()
    pdl(True)
    for _ in range(4):
        got = step()
        torch.cuda.synchronize()
        for a, b in zip(got, want):
            assert torch.equal(a, b)
    # graph replay with the programmatic edges captured
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(s)
    xs = x.clone()
    layer.weight.grad = None
    layer.bias.grad = None
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        o_static = layer(xs, gr)
        o_static.backward(g)
    for _ in range(3):
        cg.replay()
    torch.cuda.synchronize()
    assert torch.equal(o_static.detach(), want[0])
    assert torch.equal(layer.weight.grad, want[2]) and torch.equal(layer.bias.grad, want[3])
