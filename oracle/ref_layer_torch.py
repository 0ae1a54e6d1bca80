"""The reference layer's CPU path restated on the library ops it calls (torch CPU).

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/gcn_oracle.py header).  The reference
executes this path as three PyTorch calls -- pygcn/layers.py:33 `torch.mm`, :34 `torch.spmm`
with the adjacency as `utils.sparse_mx_to_torch_sparse_tensor` builds it (uncoalesced fp32 COO,
int64 indices, pygcn/utils.py:407-414), :36 the bias add -- and autograd's backward of them.
/root/reference does not exist on the GPU box, so `bench.py --impl reference` and the
`cpu_baseline` leg time this restatement ("kind": "port") on the host cores.  It is checked
against the reference's own outputs by tests/test_oracle_golden.py::test_ref_port_matches_golden.
"""
import time

import torch


def reference_layer_fwdbwd(x, weight, bias, adj, grad_out):
    """One forward + backward of the reference layer on CPU tensors; returns (out, dW, db)."""
    w = weight.detach().clone().requires_grad_(True)
    b = bias.detach().clone().requires_grad_(True) if bias is not None else None
    support = torch.mm(x, w)            # layers.py:33
    out = torch.spmm(adj, support)      # layers.py:34
    if b is not None:
        out = out + b                   # layers.py:36
    out.backward(grad_out)
    return out.detach(), w.grad, (b.grad if b is not None else None)


def make_reference_adj(indices, values, shape, form="coo_as_written"):
    """Adjacency in the form the reference hands to torch.spmm.

    coo_as_written : uncoalesced COO exactly as utils.py:414 constructs it (the headline baseline)
    csr            : the same matrix as torch sparse CSR (torch's best CPU path)
    """
    adj = torch.sparse_coo_tensor(indices, values, shape, check_invariants=False)
    if form == "coo_as_written":
        return adj
    if form == "csr":
        return adj.coalesce().to_sparse_csr()
    raise ValueError(form)


def time_reference(x, weight, bias, adj, grad_out, steps, warmup):
    """Seconds per fwd+bwd step (mean over `steps`) after `warmup` untimed steps."""
    for _ in range(warmup):
        reference_layer_fwdbwd(x, weight, bias, adj, grad_out)
    t0 = time.perf_counter()
    for _ in range(steps):
        reference_layer_fwdbwd(x, weight, bias, adj, grad_out)
    return (time.perf_counter() - t0) / max(steps, 1)
