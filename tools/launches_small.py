"""driver for an ncu launch list (gpu__time_duration only) of the paths whose helper kernels are not in
the bench's launch list: (1) BASELINE config 1 (4-D Gaussian, neval=1e4) through the one-call iteration
(k_plan / k_scan / k_super_items / k_engine / k_finalize / k_map_adapt / k_tot), (2) the same integrand through
the callback path (k_sample_x, the user's torch kernels, k_reduce), (3) a 6-parameter PDFIntegrator with f(p)
on the device (k_sample_x, k_pdf_map, k_pdf_weight, k_reduce), (4) AdaptiveMap.map / invmap / add_training_data
(k_map, k_invmap, k_add_training).   ncu ... python tools/launches_small.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vegas_b200 as vegas
from vegas_b200._gv import gv

f = vegas.integrands.GaussMix([4 * [0.5]], 100., 1013.2118364296088)
integ = vegas.Integrator([[-1., 1.]] + 3 * [[0., 1.]], neval=1e4, seed=1)
print('config 1, fused:', integ(f, nitn=3))


@vegas.devicebatchintegrand
def fdev(x):
    return torch.exp(-100. * ((x - 0.5) ** 2).sum(dim=1)) * 1013.2118364296088


integ = vegas.Integrator([[-1., 1.]] + 3 * [[0., 1.]], neval=1e4, seed=1)
print('config 1, callback path:', integ(fdev, nitn=3))

rng = np.random.default_rng(11)
a = rng.normal(size=(6, 6))
pint = vegas.PDFIntegrator(gv.gvar(rng.normal(size=6), a @ a.T + 0.5 * np.eye(6)), neval=1e5, seed=12)
pint(nitn=2)
fp = vegas.devicebatchintegrand(lambda p: torch.stack([p[:, 0], p[:, 0] * p[:, 1], p[:, 2] ** 2], dim=1))
print('PDFIntegrator:', np.asarray(pint(fp, nitn=2, adapt=False)))

m = vegas.AdaptiveMap([[0., 1.], [-1., 3.]], ninc=64)
y = rng.uniform(size=(4096, 2))
x, jac = np.empty_like(y), np.empty(4096)
m.map(y, x, jac)
m.add_training_data(y, jac * jac)
m.adapt(alpha=1.0)
m.invmap(x, y, jac)
torch.cuda.synchronize()
