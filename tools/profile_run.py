"""small driver for ncu captures: a few vegas+ iterations of one workload
   python tools/profile_run.py [ridge1000|ridge1|gauss|pathint_unfused] [neval]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vegas_b200 as vegas

what = sys.argv[1] if len(sys.argv) > 1 else 'ridge1000'
neval = float(sys.argv[2]) if len(sys.argv) > 2 else 1e7
F = vegas.integrands
if what.startswith('ridge'):
    n = int(what[5:])
    f = F.Ridge(8, N=n, lo=0.5 if n == 1 else 0.4, hi=0.5 if n == 1 else 0.6)
    integ = vegas.Integrator(8 * [[0., 1.]], neval=neval, seed=3)
    integ(f, nitn=4)
elif what == 'gauss':
    f = F.GaussMix([4 * [0.5]], 100., 1013.2118364296088)
    integ = vegas.Integrator([[-1., 1.]] + 3 * [[0., 1.]], neval=neval, seed=3)
    integ(f, nitn=4)
elif what == 'pathint_unfused':
    f = F.PathIntegral(T=4., ndT=10, x0list=np.linspace(0, 2., 6))
    NORM, NORM0, X0L = f.norm, f.norm_x0, [float(v) for v in f.x0list]

    @vegas.devicebatchintegrand
    def fdev(theta):
        x = torch.tan(theta)
        Vx = 0.5 * x * x
        a, m_2a = 0.4, 1.25
        jf = 1.0 + x * x
        jac = NORM * jf.prod(dim=1)
        jac0 = NORM0 * jf[:, 1:].prod(dim=1)
        Smid = a * Vx[:, -1] + (m_2a * (x[:, 2:] - x[:, 1:-1]) ** 2 + a * Vx[:, 1:-1]).sum(dim=1)
        out = torch.empty((x.shape[0], 7), dtype=torch.float64, device=x.device)
        for i in range(7):
            e = x[:, 0] if i == 0 else torch.full_like(x[:, 0], X0L[i - 1])
            Ve = 0.5 * e * e
            S = Smid + m_2a * ((x[:, 1] - e) ** 2 + (e - x[:, -1]) ** 2) + a * Ve
            out[:, i] = (jac if i == 0 else jac0) * torch.exp(-S)
        return out
    integ = vegas.Integrator(f.region, neval=neval, seed=3, alpha=0.1)
    integ(fdev, nitn=3)
    fused_f, f = f, fdev
elif what == 'pdf':
    from vegas_b200._gv import gv
    rng = np.random.default_rng(1)
    a = rng.normal(size=(6, 6))
    g = gv.gvar(rng.normal(size=6), a @ a.T + 0.5 * np.eye(6))
    integ = vegas.PDFIntegrator(g, neval=neval, seed=3)
    integ(nitn=3)
    f = vegas.devicebatchintegrand(lambda p: torch.stack([p[:, 0], p[:, 0] * p[:, 1], p[:, 2] ** 2], dim=1))
elif what == 'restratify':
    f = F.GaussMix([[0.5, 0.5, 0.5, 0.5]], 100., 1013.2118364296088)
    integ = vegas.Integrator(4 * [[0., 1.]], neval=neval, seed=3)
    integ(f, nitn=4)
    vegas.restratify(integ, f, nitn=2, ndy=8)
torch.cuda.synchronize()
r = integ(f, nitn=2)
torch.cuda.synchronize()
print(what, neval, r if not hasattr(r, 'keys') else r['exp(-E0*T)'], 'launches', integ.gpu_launches)
