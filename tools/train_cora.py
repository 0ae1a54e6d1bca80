"""The upstream 2-layer GCN training loop on the Cora citation graph through the drop-in layer:
the semantics of the reference's train.py (pygcn/train.py:36-47 defaults: Adam lr 0.01, weight decay 5e-4, dropout 0.5,
200 epochs; train() / test() of upstream pygcn: log_softmax + NLL on the training nodes) with the loader the fork
commented out (pygcn/utils.py:348-382): `cora.cites` -> max-symmetrised adjacency + I -> D^-1 normalisation, here built
on the device (Graph.from_edges).

`cora.content` (features, labels) is not shipped with the fork (data/cora holds the citation list only), so features
and labels are PLANTED: 7 classes drawn per node, labels smoothed over the graph so that neighbours tend to agree, 1433
sparse binary features whose probabilities depend on the class.  What the run shows is that the layer trains -- loss
falls, accuracy on held-out nodes rises well above the 1/7 chance level -- with the fused ReLU + dropout epilogue and
the library's backward; its loss curve against torch's own ops is pinned step by step in
tests/test_gpu_parity.py::test_cora_two_layer_training_follows_torch_reference.

    python tools/train_cora.py [--cites path/to/cora.cites] [--epochs 200] [--hidden 16] [--seed 42]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.nn.functional as F

import pygcn_b200 as P


class GCN(torch.nn.Module):
    """upstream pygcn models.GCN: gc1 -> relu -> dropout -> gc2 -> log_softmax (the fork's models.py:17-71 is this class
    with the dropout commented out and a third layer added)."""

    def __init__(self, nfeat, nhid, nclass, dropout):
        super().__init__()
        self.gc1 = P.GraphConvolution(nfeat, nhid, fuse_relu=True, dropout=dropout)  # ReLU + dropout in the SpMM epilogue
        self.gc2 = P.GraphConvolution(nhid, nclass)

    def forward(self, x, adj):
        return F.log_softmax(self.gc2(self.gc1(x, adj), adj), dim=1)


def planted_data(graph, n, nfeat, nclass, rs, dev):
    lab = torch.from_numpy(rs.integers(0, nclass, n)).to(dev)
    onehot = F.one_hot(lab, nclass).float()
    for _ in range(3):  # neighbours agree: three rounds of A @ onehot, then argmax
        onehot = P.spmm(graph, onehot) + 0.3 * F.one_hot(lab, nclass).float()
    lab = onehot.argmax(1)
    proto = torch.from_numpy((rs.random((nclass, nfeat)) < 0.04).astype(np.float32)).to(dev)
    p = 0.004 + 0.15 * proto[lab]
    x = (torch.rand(n, nfeat, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) < p).float()
    x = x / x.sum(1, keepdim=True).clamp_min(1.0)  # row-normalised features, like utils.normalize(features)
    return x, lab


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cites", default=None, help="edge list `cited citing` per line (data/cora/cora.cites); default: the "
                                                  "edges of tests/golden/cora_pipeline.npz (the same file, parsed)")
    ap.add_argument("--epochs", type=int, default=200)
    ap.add_argument("--hidden", type=int, default=16)
    ap.add_argument("--dropout", type=float, default=0.5)
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--weight_decay", type=float, default=5e-4)
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rs = np.random.default_rng(args.seed)
    torch.manual_seed(args.seed)
    if args.cites:
        graph, ids = P.io.graph_from_cites(args.cites, dev)
        n = graph.n_rows
    else:
        g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "cora_pipeline.npz"))
        n = int(g["n"])
        e = torch.from_numpy(g["edges"]).to(dev)
        graph = P.Graph.from_edges(e[:, 0], e[:, 1], n)
    x, labels = planted_data(graph, n, 1433, 7, rs, dev)
    idx_train, idx_val, idx_test = torch.arange(140, device=dev), torch.arange(200, 500, device=dev), torch.arange(500, 1500, device=dev)
    model = GCN(1433, args.hidden, 7, args.dropout).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.weight_decay)
    t0 = time.time()
    for epoch in range(args.epochs):
        model.train()
        opt.zero_grad()
        out = model(x, graph)
        loss = F.nll_loss(out[idx_train], labels[idx_train])
        loss.backward()
        opt.step()
        if epoch % 20 == 0 or epoch == args.epochs - 1:
            model.eval()
            with torch.no_grad():
                out = model(x, graph)
                acc_val = (out[idx_val].argmax(1) == labels[idx_val]).float().mean().item()
            print("epoch %4d loss_train %.4f acc_val %.4f time %.2fs" % (epoch + 1, loss.item(), acc_val, time.time() - t0), flush=True)
    model.eval()
    with torch.no_grad():
        out = model(x, graph)
        acc = (out[idx_test].argmax(1) == labels[idx_test]).float().mean().item()
        loss_test = F.nll_loss(out[idx_test], labels[idx_test]).item()
    print("Test set results: loss= %.4f accuracy= %.4f (chance 0.1429; graph %r)" % (loss_test, acc, graph))
    return 0 if acc > 0.5 else 1


if __name__ == "__main__":
    sys.exit(main())
