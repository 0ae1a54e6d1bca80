"""pygcn_b200 -- the GCN layer hot path of LinChen-65/pygcn on B200 (sm_100a).

Public surface (mirrors the reference's for this path):
    GraphConvolution      drop-in for pygcn/layers.py::GraphConvolution
    spmm                  drop-in for torch.spmm(adj, dense)
    gcn_layer             functional form of the layer
    Graph                 device-resident adjacency (replaces utils.normalize /
                          utils.sparse_mx_to_torch_sparse_tensor, built on the GPU)
    load_adj              the CBG adjacency avg^T avg of utils.load_adj, on the device
    apply_bn              (ReLU ->) fresh BatchNorm1d of models.GCN.apply_bn (opt-in, not yet run on hardware)
    io                    the on-disk formats either side of the layer: `.cites` edge lists, the npy / pickle levels
                          of utils.load_adj (host-side parsing in front of the device paths)
    install_as_pygcn      make `import layers` / `import pygcn.layers` resolve to this package
"""
import sys as _sys

from .functional import apply_bn, gcn_layer, load_adj, mm, spmm
from . import io
from .graph import Graph, as_graph, clear_cache
from .layers import GraphConvolution

__all__ = ["GraphConvolution", "Graph", "as_graph", "clear_cache", "gcn_layer", "spmm", "mm", "load_adj", "apply_bn", "io", "install_as_pygcn"]


def install_as_pygcn():
    """Register this package's layer module under the names the reference imports.

    pygcn/models.py:4 does `from layers import GraphConvolution`; whole-model pickles written
    by gnn-over-mlp.py:489 record the module path `layers`.  After this call both resolve to
    the B200 implementation without touching the reference's files.
    """
    from . import layers as _layers

    _sys.modules.setdefault("layers", _layers)
    _sys.modules.setdefault("pygcn.layers", _layers)
    return _layers
