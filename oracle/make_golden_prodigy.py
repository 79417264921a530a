"""Generates tests/golden/prodigy.pt by running the UNMODIFIED reference optimizer (ldm/prodigy.py) in the build
container:  PYTHONPATH=/root/reference PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_prodigy.py"""
import importlib.util
import os
import sys

import torch

spec = importlib.util.spec_from_file_location("ref_prodigy", "/root/reference/ldm/prodigy.py")
mod = importlib.util.module_from_spec(spec)
spec.loader.exec_module(mod)

out = {}
for name, kw in {"default": {}, "decay_biascorr": {"weight_decay": 0.01, "use_bias_correction": True, "safeguard_warmup": True},
                 "coupled_decay_growth": {"weight_decay": 0.02, "decouple": False, "growth_rate": 1.5, "d_coef": 2.0}}.items():
    g = torch.Generator().manual_seed(1)
    params = [torch.nn.Parameter(torch.randn(33, 7, generator=g)), torch.nn.Parameter(torch.randn(129, generator=g))]
    init = [p.detach().clone() for p in params]
    opt = mod.Prodigy(params, lr=1.0, **kw)
    # gradient of 0.5 * |p - target|^2 plus noise, evaluated at the CURRENT parameters (so that d adapts)
    targets = [torch.randn(p.shape, generator=g) for p in params]
    noises, ds = [], []
    for step in range(12):
        ns = [torch.randn(p.shape, generator=g) * 0.1 for p in params]
        for p, t, n in zip(params, targets, ns):
            p.grad = (p.detach() - t) + n
        opt.step()
        noises.append(ns)
        ds.append(opt.param_groups[0]["d"])
    out[name] = {"kw": kw, "init": init, "targets": targets, "noises": noises,
                 "final": [p.detach().clone() for p in params], "d": ds}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "prodigy.pt")
torch.save(out, path)
print("wrote", path, {k: v["d"][-1] for k, v in out.items()})
