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


# ---------------------------------------------------------------------------- a stack of layers (bench.py workloads)
def build_reference_stack(dims, params):
    """The model of a bench workload on the reference's own class: one `GraphConvolution(fin, fout)` per consecutive
    pair of `dims`, parameters copied from `params` = [(weight, bias), ...].  Uses the UNMODIFIED class out of
    oracle/_ref/layers.py when oracle/make_ref.py has put it there ("reference"), else this file's port ("port")."""
    from . import ref_runtime

    mod = ref_runtime.load_reference_layers()
    if mod is None:
        return None, "port"
    layers = []
    for (fin, fout), (w, b) in zip(zip(dims[:-1], dims[1:]), params):
        gc = mod.GraphConvolution(fin, fout)
        with torch.no_grad():
            gc.weight.copy_(w)
            gc.bias.copy_(b)
        layers.append(gc)
    return layers, "reference"


def reference_stack_fwdbwd(layers, params, relu, x, adj, grad_out):
    """Forward + backward of the stack as the reference's models write it (`x = F.relu(self.gcK(x, adj))`,
    pygcn/models.py:102-111 GeneratorGCN.forward; relu=False: the bare layer).  Returns the output and the gradients
    [(dW, db), ...]."""
    h = x
    if layers is not None:
        for gc in layers:
            gc.weight.grad = None
            gc.bias.grad = None
            h = gc(h, adj)
            if relu:
                h = torch.nn.functional.relu(h)
        h.backward(grad_out)
        return h.detach(), [(gc.weight.grad, gc.bias.grad) for gc in layers]
    ws = [(w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)) for w, b in params]
    for w, b in ws:
        h = torch.spmm(adj, torch.mm(h, w)) + b
        if relu:
            h = torch.nn.functional.relu(h)
    h.backward(grad_out)
    return h.detach(), [(w.grad, b.grad) for w, b in ws]


def time_reference_stack(dims, params, relu, x, adj, grad_out, steps, warmup):
    """(seconds per fwd+bwd step, "reference" | "port")."""
    layers, kind = build_reference_stack(dims, params)
    for _ in range(warmup):
        reference_stack_fwdbwd(layers, params, relu, x, adj, grad_out)
    t0 = time.perf_counter()
    for _ in range(steps):
        reference_stack_fwdbwd(layers, params, relu, x, adj, grad_out)
    return (time.perf_counter() - t0) / max(steps, 1), kind
