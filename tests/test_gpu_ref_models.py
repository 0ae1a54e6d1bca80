"""The reference's REAL model classes on the B200 through the drop-in layer (north_star: "models.GCN, train.py and the
repo's policy-generator and regression scripts run unchanged").

oracle/_ref/ holds the reference's own models.py / layers.py / utils.py (copied there by oracle/make_ref.py, run by
__graft_entry__.build(); never committed).  Each test imports models.py TWICE -- once against the reference's own
`layers` module (stock: torch.mm / torch.spmm on CUDA), once with `pygcn_b200.layers` registered under the name
models.py:4 imports (`from layers import GraphConvolution`) -- builds the same model from the same seed, runs the call
pattern of the live scripts (policy-generator.py:389-420: dense `adj`, column-slice inputs, anomaly mode,
`backward(retain_graph=True)`), and compares outputs and every parameter gradient.  Bar: a third run of
the stock classes in DOUBLE precision is the arbiter -- the dense row-normalised adjacency makes every layer output
nearly constant over the nodes, and the fresh BatchNorm then divides by a column deviation ~1e-3 of the column mean, so
two correct fp32 evaluations differ by ~1e-3 after it: our error against fp64 must be <= max(1e-5, 8 x the stock fp32
run's error against fp64) (the layer's dense-adjacency route is a tcgen05 3xTF32 product, 1e-6 from fp64 where cuBLAS'
fp32 FMA is 2e-7 -- both far inside the layer's 1e-5 bar; the factor is what the BatchNorm leaves of that ratio).
  models.GCN           3 layers, F.relu + a fresh .cuda() BatchNorm1d after the first two   (models.py:17-71)
  models.GCN_OVER_MLP  the per-sample loop over B = 20 samples, pooling, MLP head           (models.py:333-355)
  models.Generator     GeneratorGCN + MLP head with BatchNorm + top-NN selection            (models.py:358-379)
"""
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import ref_runtime  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_runtime.available("layers.py", "models.py", "utils.py", "constants.py"),
                                 reason="oracle/_ref/ not made (python oracle/make_ref.py needs /root/reference)")]
TOL = 1e-5
N, FEAT, TOUCHED, HID = 2943, 10, 8, 32  # the fork's evaluator shape: 2943 CBGs, dense adjacency, 8 -> 32 -> 32 -> 32


def dev():
    return torch.device("cuda:0")


def nerr(a, b):
    a, b = a.detach().double(), b.detach().double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture(scope="module")
def both():
    """(stock models module, drop-in models module)."""
    import pygcn_b200 as P
    from pygcn_b200 import layers as our_layers

    _, _, stock = ref_runtime.load_reference_models()
    lay, _, dropin = ref_runtime.load_reference_models(layers_module=our_layers)
    assert lay is our_layers and dropin.GraphConvolution is P.GraphConvolution
    assert stock.GraphConvolution is not P.GraphConvolution
    return stock, dropin


@pytest.fixture(scope="module")
def inputs():
    gen = torch.Generator(device="cpu").manual_seed(3)
    visits = torch.rand(400, N, generator=gen) * (torch.rand(400, N, generator=gen) < 0.05)
    adj = (visits.T @ visits)  # utils.load_adj's V^T V: dense, non-negative (pygcn/utils.py:124-128)
    adj = adj / adj.sum(1, keepdim=True).clamp_min(1e-6)
    x = torch.rand(N, FEAT, generator=gen)
    x[:, -1] = (torch.rand(N, generator=gen) < 0.02).float()  # vaccination flag column (PoolLayer's mask)
    xb = torch.rand(20, N, FEAT, generator=gen)
    xb[:, :, -1] = x[:, -1]
    return adj.to(dev()), x.to(dev()), xb.to(dev())


def config(pooled=True):
    """pooled: GCN_OVER_MLP's PoolLayer drops the flag column (models.py:279); Generator's head sees all of them."""
    return types.SimpleNamespace(gcn_nfeat=TOUCHED, gcn_nhid=HID, gcn_nclass=HID, gcn_dropout=0.1, NN=70,
                                 linear_nin=HID + FEAT - TOUCHED - (1 if pooled else 0), linear_nhid1=32, linear_nhid2=32, linear_nout=1,
                                 linear_activation="relu", linear_bias=True, dim_touched=TOUCHED)


def build(mod, cls, *args):
    torch.manual_seed(11)
    return getattr(mod, cls)(*args).to(dev())


class double_default:
    """torch's default dtype = float64 inside the block: the reference's apply_bn builds a FRESH nn.BatchNorm1d per
    call (models.py:41-45), which takes the default dtype -- the arbiter run needs it in double like everything else."""

    def __enter__(self):
        self.prev = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)

    def __exit__(self, *exc):
        torch.set_default_dtype(self.prev)
        return False


def build64(mod, cls, *args):
    """The stock model in double precision with the fp32 model's initial values."""
    m32 = build(mod, cls, *args)
    with double_default():
        m64 = getattr(mod, cls)(*args)
    m64.load_state_dict({k: v.double() for k, v in m32.state_dict().items()})
    return m64.double().to(dev())


def within(ours, stock32, ref64, what, tol=TOL):
    e_ours, e_stock = nerr(ours, ref64), nerr(stock32, ref64)
    assert e_ours <= max(tol, 8 * e_stock), (what, e_ours, e_stock)


def grads_within(m_drop, m_stock, m64):
    ps, pd, p64 = dict(m_stock.named_parameters()), dict(m_drop.named_parameters()), dict(m64.named_parameters())
    assert ps.keys() == pd.keys() == p64.keys()
    for k in ps:
        assert (ps[k].grad is None) == (pd[k].grad is None), k
        if ps[k].grad is not None:
            within(pd[k].grad, ps[k].grad, p64[k].grad, k)


def test_models_gcn_with_fresh_cuda_batchnorm(both, inputs):
    stock, dropin = both
    adj, x, _ = inputs
    ms, md = build(stock, "GCN", TOUCHED, HID, HID, 0.1, 70), build(dropin, "GCN", TOUCHED, HID, HID, 0.1, 70)
    assert [k for k, _ in ms.named_parameters()] == [k for k, _ in md.named_parameters()]
    for (_, a), (_, b) in zip(ms.named_parameters(), md.named_parameters()):
        assert torch.equal(a, b)  # same draws in the same order (layers.py:23-29)
    m64 = build64(stock, "GCN", TOUCHED, HID, HID, 0.1, 70)
    with torch.autograd.set_detect_anomaly(True):
        outs = []
        for m in (ms, md):
            o = m(x[:, :TOUCHED], adj)  # column-slice view, as models.py:345 / :368 pass it
            o.square().mean().backward(retain_graph=True)
            outs.append(o)
        with double_default():
            o64 = m64(x.double()[:, :TOUCHED], adj.double())
            o64.square().mean().backward(retain_graph=True)
    within(outs[1], outs[0], o64, "out")
    grads_within(md, ms, m64)


def test_models_gcn_over_mlp_per_sample_loop(both, inputs):
    stock, dropin = both
    adj, _, xb = inputs
    ms, md = build(stock, "GCN_OVER_MLP", config()), build(dropin, "GCN_OVER_MLP", config())
    m64 = build64(stock, "GCN_OVER_MLP", config())
    with torch.autograd.set_detect_anomaly(True):
        outs = []
        for m in (ms, md):
            o = m(xb, adj)  # 20 samples x 3 layers = 60 layer calls on the same adj (models.py:343-349)
            assert o.shape == (20, 1)
            o.square().mean().backward(retain_graph=True)
            outs.append(o)
        with double_default():
            o64 = m64(xb.double(), adj.double())
            o64.square().mean().backward(retain_graph=True)
    within(outs[1], outs[0], o64, "out")
    grads_within(md, ms, m64)


def test_models_generator_forward_backward(both, inputs, capsys):
    stock, dropin = both
    adj, x, _ = inputs
    ms, md = build(stock, "Generator", config(False)), build(dropin, "Generator", config(False))
    m64 = build64(stock, "Generator", config(False))
    captured = []
    outs = []
    with torch.autograd.set_detect_anomaly(True):
        for m in (ms, md, m64):
            orig = m.MLPLayers.forward  # (models.py:370 calls .forward directly: module hooks do not fire)

            def spy(inp, orig=orig):
                out = orig(inp)
                captured.append(out.detach().clone())
                return out
            m.MLPLayers.forward = spy
            if m is m64:
                with double_default():
                    o = m(x.double(), adj.double())
                    o.sum().backward(retain_graph=True)
            else:
                o = m(x, adj)
                o.sum().backward(retain_graph=True)
            m.MLPLayers.forward = orig
            assert o.shape == (N, 1)
            outs.append(o)
    capsys.readouterr()  # models.py:371 prints statistics
    within(captured[1], captured[0], captured[2], "scores")  # the scores the head ranks
    # the NN CBGs above the threshold are a discrete choice: gradients are comparable when all three runs pick the same
    if torch.equal(outs[0] != 0, outs[1] != 0) and torch.equal(outs[0] != 0, outs[2] != 0):
        grads_within(md, ms, m64)
    else:
        assert ((outs[0] != 0) ^ (outs[1] != 0)).sum().item() <= 4


def test_generator_gcn_state_dict_and_pickle_cross_load(both, inputs, tmp_path):
    """A model trained on the stock layer loads into the drop-in (same parameter names), and a whole-model pickle
    written by the scripts (`torch.save(model)`, gnn-over-mlp.py:489) resolves `layers.GraphConvolution`."""
    stock, dropin = both
    adj, x, _ = inputs
    ms, md = build(stock, "GeneratorGCN", TOUCHED, HID, HID, 0.1, 70), build(dropin, "GeneratorGCN", TOUCHED, HID, HID, 0.1, 70)
    with torch.no_grad():
        for p in ms.parameters():
            p.add_(0.01)
    md.load_state_dict(ms.state_dict())
    m64 = build64(stock, "GeneratorGCN", TOUCHED, HID, HID, 0.1, 70)
    m64.load_state_dict({k: v.double() for k, v in ms.state_dict().items()})
    within(md(x[:, :TOUCHED], adj), ms(x[:, :TOUCHED], adj), m64(x.double()[:, :TOUCHED], adj.double()), "out")
