"""Sweep NGP_MLP_SLOTS_FW / NGP_MLP_SLOTS_BW on the two MLPs of the headline model: time + max error vs the
bf16-operand oracle.  Usage: python tools/mlp_sweep.py [log2_samples=23]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "instant-ngp-pp_b200")]
import torch
from ngp_b200 import tcnn
from oracle import tcnn_oracle

torch.manual_seed(0)
N = int(2 ** float(sys.argv[1])) if len(sys.argv) > 1 else 9_500_000
dev = "cuda"
cfg = lambda nh, oa: {"otype": "FullyFusedMLP", "activation": "ReLU", "output_activation": oa, "n_neurons": 64, "n_hidden_layers": nh}
sig = tcnn.Network(32, 16, cfg(1, "None")).to(dev)
rgb = tcnn.Network(32, 3, cfg(2, "Sigmoid")).to(dev)
y = torch.randn(N, 32, device=dev) * 0.5
d = torch.randn(N, 3, device=dev)
h = torch.randn(N, 16, device=dev) * 0.5
dh = torch.randn(N, 16, device=dev)
drgb = torch.randn(N, 3, device=dev)
dsig = torch.randn(N, device=dev)


def tm(fn, it=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it


def ref_grads(net, nin, nh, nout, oa, x, dout, n=20000):
    p = net.params.detach().clone().requires_grad_(True)
    xx = x[:n].clone().requires_grad_(True)
    o = tcnn_oracle.mlp_forward(xx, p, nin, 64, nh, nout, "ReLU", oa, operand_dtype=torch.bfloat16)
    o.backward(dout[:n])
    return o.detach(), p.grad, xx.grad


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


sh = lambda dd: tcnn_oracle.sh_encode((torch.nn.functional.normalize(dd, dim=-1) + 1) / 2)
n_chk = 20000
x_rgb = torch.cat([sh(d[:n_chk]), h[:n_chk]], 1)
o_s, gp_s, gx_s = ref_grads(sig, 32, 1, 16, "None", y, dh)
o_r, gp_r, gx_r = ref_grads(rgb, 32, 2, 3, "Sigmoid", x_rgb, drgb)

import itertools
for bulk, sf in itertools.product((0, 1), (2, 3, 4, 5, 6, 8)):
    os.environ["NGP_MLP_SLOTS_FW"] = str(sf); os.environ["NGP_MLP_BULK"] = str(bulk)
    t1 = tm(lambda: tcnn.mlp_forward([(y, 32, 0)], sig.params.detach(), sig.mlp))
    t2 = tm(lambda: tcnn.mlp_forward([(d, 16, 1), (h, 16, 0)], rgb.params.detach(), rgb.mlp))
    e1 = rel(tcnn.mlp_forward([(y[:n_chk], 32, 0)], sig.params.detach(), sig.mlp), o_s)
    e2 = rel(tcnn.mlp_forward([(d[:n_chk], 16, 1), (h[:n_chk], 16, 0)], rgb.params.detach(), rgb.mlp), o_r)
    print(f"FW bulk={bulk} slots={sf}: sigma {t1:.3f} ms  rgb {t2:.3f} ms   relerr {e1:.2e} {e2:.2e}", flush=True)
for bulk, sb in itertools.product((0, 1), (2, 3, 4, 5, 6)):
    os.environ["NGP_MLP_SLOTS_BW"] = str(sb); os.environ["NGP_MLP_BULK"] = str(bulk)
    t1 = tm(lambda: tcnn.mlp_backward([(y, 32, 0)], sig.params.detach(), sig.mlp, dh, [True]))
    t2 = tm(lambda: tcnn.mlp_backward([(d, 16, 1), (h, 16, 0)], rgb.params.detach(), rgb.mlp, drgb, [False, True]))
    dp1, ds1 = tcnn.mlp_backward([(y[:n_chk], 32, 0)], sig.params.detach(), sig.mlp, dh[:n_chk], [True])
    dp2, ds2 = tcnn.mlp_backward([(d[:n_chk], 16, 1), (h[:n_chk], 16, 0)], rgb.params.detach(), rgb.mlp, drgb[:n_chk], [False, True])
    print(f"BW bulk={bulk} slots={sb}: sigma {t1:.3f} ms  rgb {t2:.3f} ms   relerr dW {rel(dp1, gp_s):.2e} {rel(dp2, gp_r):.2e}  dX {rel(ds1[0], gx_s):.2e} {rel(ds2[1], gx_r[:, 16:]):.2e}", flush=True)
