"""First hardware run of the (ReLU ->) fresh-BatchNorm op (pygcn_b200.apply_bn, csrc/batchnorm.cu):
parity with torch's own ReLU + BatchNorm1d + autograd, then forward+backward time of both, as the models use them
(pygcn/models.py:41-45, 49, 53) on the hidden panels of the BASELINE shapes.

    python tools/bn_probe.py [reps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pygcn_b200 as P  # noqa: E402


def timed(fn, reps, flush):
    for _ in range(3):
        fn()
    evs = []
    for _ in range(reps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in evs) / reps * 1e3  # us


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device("cuda:0")
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    for name, n, f in (("CBG hidden", 100000, 32), ("Reddit hidden", 232965, 256), ("products hidden", 2449029, 256)):
        gen = torch.Generator(device=dev).manual_seed(f)
        y = (torch.randn(n, f, generator=gen, device=dev) + 0.25).requires_grad_(True)
        g = torch.randn(n, f, generator=gen, device=dev)

        def ours():
            y.grad = None
            P.apply_bn(y, relu=True).backward(g)

        def theirs():
            y.grad = None
            torch.nn.functional.batch_norm(torch.relu(y), None, None, None, None, True, 0.1, 1e-5).backward(g)

        ours()
        dy = y.grad.clone()
        out = P.apply_bn(y, relu=True).detach()
        theirs()
        ref = torch.nn.functional.batch_norm(torch.relu(y), None, None, None, None, True, 0.1, 1e-5).detach()
        e_out = float((out - ref).abs().max() / ref.abs().max())
        e_dy = float((dy - y.grad).abs().max() / y.grad.abs().max())
        for label, flush in (("L2 flushed", lambda: flush_buf.fill_(1)), ("input L2-warm (as after the SpMM)", lambda: None)):
            print("%-16s [%d x %d] %s: apply_bn fwd+bwd %.1f us, torch relu + batch_norm fwd+bwd %.1f us; parity out %.1e dy %.1e"
                  % (name, n, f, label, timed(ours, reps, flush), timed(theirs, reps, flush), e_out, e_dy))
        bytes_min = n * f * 4 * (2 + 3)  # forward: read y, write out; backward: read y, g, write dy
        print("    algorithmic bytes fwd+bwd %.1f MB (statistics passes re-read y / g out of L2 when they fit)" % (bytes_min / 1e6))


if __name__ == "__main__":
    main()
