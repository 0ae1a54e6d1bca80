"""Drop-in for `pygcn/layers.py::GraphConvolution` backed by libgcnb200.so.

Same constructor, attribute and parameter names (`weight` [in, out], `bias` [out]),
same initialisation calls in the same RNG order (pygcn/layers.py:23-29), same `forward(input,
adj)` contract and `__repr__` (pygcn/layers.py:40-43), so `models.GCN`, `GeneratorGCN`, ...
(pygcn/models.py:17-177) and state dicts / whole-model pickles keep working.  The arithmetic
of forward and backward runs in hand-written sm_100a kernels; CPU tensors raise.
"""
import math

import torch
from torch.nn.modules.module import Module
from torch.nn.parameter import Parameter

from .functional import gcn_layer


class GraphConvolution(Module):
    """GCN layer `adj @ (input @ weight) + bias` (https://arxiv.org/abs/1609.02907).

    Extra keyword-only options (defaults reproduce the reference exactly):
      fuse_relu  -- apply the ReLU every caller in pygcn/models.py applies, inside the SpMM
                    epilogue (a following F.relu is then a no-op, results are identical)
      dropout    -- p of a dropout fused after the (optional) ReLU, active in training mode only
                    (upstream pygcn applies F.dropout to the layer output; the fork commented it out)
      precision  -- "auto" (default: tcgen05 3xTF32 when the product is large enough, fp32 CUDA
                    cores otherwise), "tf32x3" or "fp32" for the dense products; "bf16" = the reduced
                    tier (<= 2e-2): the panels the SpMM gathers are rounded to bf16, everything else fp32
      association -- "auto" (default) computes (adj @ input) @ weight when in_features < out_features makes that the
                    order with less SpMM traffic (same result to fp32 rounding, functional._aggregate_first);
                    "reference" keeps adj @ (input @ weight) always, "aggregate_first" forces the other order
    """

    def __init__(self, in_features, out_features, bias=True, *, fuse_relu=False, dropout=0.0, precision="auto",
                 association="auto"):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.fuse_relu = fuse_relu
        self.dropout = dropout
        if dropout and precision == "bf16":
            raise ValueError("the fused dropout mask is not available in the bf16 panel tier (precision='bf16')")
        self.precision = precision
        self.association = association
        self.weight = Parameter(torch.empty(in_features, out_features, dtype=torch.float32))
        if bias:
            self.bias = Parameter(torch.empty(out_features, dtype=torch.float32))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        # same two draws, same order, same bounds as pygcn/layers.py:23-29:
        # kaiming_uniform_ on a [in, out] tensor takes fan_in = size(1) = out_features
        bound = 1.0 / math.sqrt(self.weight.size(1))
        torch.nn.init.kaiming_uniform_(self.weight)
        if self.bias is not None:
            self.bias.data.uniform_(-bound, bound)

    def forward(self, input, adj):
        """input [N, in_features] (the reference's call), or [B, N, in_features]: B feature matrices
        over the same adjacency in one pass (what GCN_OVER_MLP's per-sample loop computes,
        pygcn/models.py:343-349) -> [B, N, out_features]."""
        p = getattr(self, "dropout", 0.0)
        mask = None
        if p > 0.0 and self.training and input.dim() != 2:
            raise NotImplementedError("dropout > 0 with batched input [B, N, F]: the fused mask is per [N, F] output "
                                      "(apply F.dropout to the result, or call the layer per sample)")
        if p > 0.0 and self.training:
            n_rows = adj.shape[0]
            mask = torch.rand(n_rows, self.out_features, device=input.device) >= p
        return gcn_layer(input, adj, self.weight, self.bias, relu=getattr(self, "fuse_relu", False),
                         precision=getattr(self, "precision", "auto"), dropout_mask=mask, dropout_p=p,
                         association=getattr(self, "association", "auto"))

    def __repr__(self):
        return "%s (%s -> %s)" % (self.__class__.__name__, self.in_features, self.out_features)
