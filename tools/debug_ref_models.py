"""Per-layer outputs of the reference's models.GCN on the stock layer (fp32), the drop-in layer and the stock layer in
fp64: where along the model the two fp32 runs leave the fp64 one.  Development tool for tests/test_gpu_ref_models.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch

import test_gpu_ref_models as T


def main():
    import pygcn_b200 as P
    from pygcn_b200 import layers as our_layers
    from oracle import ref_runtime

    _, _, stock = ref_runtime.load_reference_models()
    _, _, dropin = ref_runtime.load_reference_models(layers_module=our_layers)
    gen = torch.Generator(device="cpu").manual_seed(3)
    N, FEAT, TOUCHED, HID = T.N, T.FEAT, T.TOUCHED, T.HID
    visits = torch.rand(400, N, generator=gen) * (torch.rand(400, N, generator=gen) < 0.05)
    adj = (visits.T @ visits)
    adj = (adj / adj.sum(1, keepdim=True).clamp_min(1e-6)).cuda()
    x = torch.rand(N, FEAT, generator=gen).cuda()
    ms, md = T.build(stock, "GCN", TOUCHED, HID, HID, 0.1, 70), T.build(dropin, "GCN", TOUCHED, HID, HID, 0.1, 70)
    m64 = T.build64(stock, "GCN", TOUCHED, HID, HID, 0.1, 70)
    caps = {}
    for name, m in (("stock", ms), ("ours", md), ("f64", m64)):
        caps[name] = {}
        for ln in ("gc1", "gc2", "gc3"):
            getattr(m, ln).register_forward_hook(lambda mod, inp, out, name=name, ln=ln: caps[name].__setitem__(ln, (inp[0].detach().clone(), out.detach().clone())))
    ms(x[:, :TOUCHED], adj)
    md(x[:, :TOUCHED], adj)
    with T.double_default():
        m64(x.double()[:, :TOUCHED], adj.double())
    for ln in ("gc1", "gc2", "gc3"):
        i64, o64 = caps["f64"][ln]
        for name in ("stock", "ours"):
            i_, o_ = caps[name][ln]
            print("%s %-5s input err %.2e  output err %.2e   | out col std/mean %.2e" % (
                ln, name, T.nerr(i_, i64), T.nerr(o_, o64), float((o64.std(0) / o64.mean(0).abs().clamp_min(1e-30)).median())))
    # the layer alone on identical inputs
    for ln in ("gc1", "gc2", "gc3"):
        i64, o64 = caps["f64"][ln]
        xin = i64.float()
        o_s = getattr(ms, ln)(xin, adj)
        o_d = getattr(md, ln)(xin, adj)
        o_6 = getattr(m64, ln)(i64, adj.double())
        print("%s alone on the fp64 run's input: stock %.2e  ours %.2e" % (ln, T.nerr(o_s, o_6), T.nerr(o_d, o_6)))
    for prec in ("fp32", "tf32x3"):
        l = P.GraphConvolution(HID, HID, precision=prec).cuda()
        l.load_state_dict(md.gc2.state_dict())
        i64, o64 = caps["f64"]["gc2"]
        print("gc2 alone precision=%s: %.2e" % (prec, T.nerr(l(i64.float(), adj), m64.gc2(i64, adj.double()))))


if __name__ == "__main__":
    main()
