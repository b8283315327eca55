"""Feasibility probe for a two-term FP16 split GEMM ('what would move it further'): over one backbone forward,
the largest |activation| and |weight| entering every Conv/Linear (fp16 overflows at 65504), and the error of the emulated split
product  x = h1 + 2^-11 h2,  w = g1 + 2^-11 g2,  y ~ h1 g1 + 2^-11 (h1 g2 + h2 g1)  against float64 on sampled rows."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
model = bench.build_model(dev).eval()
b = {k: v.to(dev) for k, v in bench.make_pool(1, 1000)[0].items()}
rows = []


def split(t):
    h1 = t.half().float()
    h2 = ((t - h1) * 2048.0).half().float()
    return h1, h2


def hook(mod, inp, out):
    x = inp[0].detach()
    w = mod.weight.detach()
    rec = {"layer": type(mod).__name__ + str(tuple(w.shape)), "x_max": x.abs().max().item(), "w_max": w.abs().max().item(), "err": None}
    if w.dim() == 2 or (w.dim() >= 3 and all(s == 1 for s in w.shape[2:])):  # 1x1 conv / linear: emulate on sampled rows
        w2 = w.reshape(w.shape[0], -1)
        xr = x if x.dim() == 2 else x.transpose(1, -1).reshape(-1, x.shape[1]) if x.dim() > 2 and x.shape[1] == w2.shape[1] else None
        if xr is not None and xr.shape[-1] == w2.shape[1]:
            xr = xr[torch.randperm(xr.shape[0], device=dev)[:4096]]
            h1, h2 = split(xr)
            g1, g2 = split(w2)
            y = h1 @ g1.t() + (h1 @ g2.t() + h2 @ g1.t()) / 2048.0
            want = xr.double() @ w2.double().t()
            rec["err"] = ((y.double() - want).abs().max() / want.abs().max().clamp_min(1e-30)).item()
    rows.append(rec)


for m in model.modules():
    if isinstance(m, (torch.nn.Conv2d, torch.nn.Conv1d, torch.nn.Linear)):
        m.register_forward_hook(hook)
with torch.no_grad():
    model(b["points"], b["image"], b["xy"].clone())
print("layers: %d   max |x| = %.3g   max |w| = %.3g   (fp16 max 65504)" % (len(rows), max(r["x_max"] for r in rows), max(r["w_max"] for r in rows)))
errs = [r["err"] for r in rows if r["err"] is not None]
print("emulated split product, max abs err / output scale over %d 1x1 layers: worst %.2e, median %.2e" % (len(errs), max(errs), sorted(errs)[len(errs) // 2]))
for r in sorted(rows, key=lambda r: -r["x_max"])[:6]:
    print("  %-34s |x| <= %.3g  |w| <= %.3g  err %s" % (r["layer"], r["x_max"], r["w_max"], "%.2e" % r["err"] if r["err"] is not None else "-"))
