"""Diagnostic: per-loss relative error of two optimize_parameters calls (fp32 mode) against the golden fixtures of the
real reference, for opt.net in {'1', '2', 'SRdens'}."""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import srcgan_oracle as O  # noqa: E402
from srcgan_b200 import nn as snn, trainer  # noqa: E402

snn.set_precision("fp32")
G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
gray = torch.load(os.path.join(G, "step_gray_tiny.pt"), weights_only=False)
for net in ("2", "SRdens"):
    fx = gray[net]
    random.seed(5)
    opt = trainer.params()
    opt.device = torch.device("cuda:0")
    opt.mode, opt.net = "x4", net
    m = trainer.SRCycleGAN(opt)
    for name, sd in O.default_states(0, net).items():
        getattr(m, "net" + name).load_state_dict(sd, strict=True)
    for it, rec in enumerate(fx["steps"]):
        a, b = O.synthetic_gray_batch(2, lr=16, scale=4, seed=1234 + it)
        m.optimize_parameters(a.cuda(), b.cuda())
        got = m.current_losses()
        print(net, it, {n: "%.2e" % (abs(got[n] - v) / max(abs(v), 1e-30)) for n, v in rec["losses"].items()})
