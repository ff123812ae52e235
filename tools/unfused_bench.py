"""HBM-bound path: sample -> device batch callback -> reduce (10-D path integral, nf = 7).
   python tools/unfused_bench.py [neval]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vegas_b200 as vegas

neval = float(sys.argv[1]) if len(sys.argv) > 1 else 3e7
f = vegas.integrands.PathIntegral(T=4., ndT=10, x0list=np.linspace(0, 2., 6))
x0 = torch.tensor(f.x0list, dtype=torch.float64, device='cuda')


@vegas.devicebatchintegrand
def fdev(theta):
    x = torch.tan(theta)
    Vx = 0.5 * x * x
    a, m_2a = f.T / f.ndT, f.m / 2. / (f.T / f.ndT)
    jf = 1.0 + x * x
    jac = f.norm * jf.prod(dim=1)
    jac0 = f.norm_x0 * jf[:, 1:].prod(dim=1)
    Smid = a * Vx[:, -1] + (m_2a * (x[:, 2:] - x[:, 1:-1]) ** 2 + a * Vx[:, 1:-1]).sum(dim=1)
    e = torch.cat([x[:, :1], x0[None, :].expand(x.shape[0], 6)], dim=1)
    Ve = 0.5 * e * e
    S = Smid[:, None] + m_2a * ((x[:, 1:2] - e) ** 2 + (e - x[:, -1:]) ** 2) + a * Ve
    pref = torch.cat([jac[:, None], jac0[:, None].expand(x.shape[0], 6)], dim=1)
    return pref * torch.exp(-S)


integ = vegas.Integrator(f.region, neval=neval, seed=3, alpha=0.1, max_batch=1 << 23)
integ(fdev, nitn=3)
integ._timing = []
integ._unfused_events = []
r = integ(fdev, nitn=3)
torch.cuda.synchronize()
rows = sum(n for _, n in integ._unfused_events)
ts = sum(ev[0].elapsed_time(ev[1]) for ev, _ in integ._unfused_events)
tc = sum(ev[1].elapsed_time(ev[2]) for ev, _ in integ._unfused_events)
tr = sum(ev[2].elapsed_time(ev[3]) for ev, _ in integ._unfused_events)
D, nf = 10, 7
print('rows %d  sample %.2f ms (%.0f GB/s of %d B/row)  callback %.2f ms  reduce %.2f ms (%.0f GB/s of %d B/row)' % (
    rows, ts, rows * (8 * D + 8) / ts / 1e6, 8 * D + 8, tc, tr, rows * (8 * nf + 8) / tr / 1e6, 8 * nf + 8))
print('engine (sample+reduce) %.3e samples/s ; with torch callback %.3e samples/s' % (rows / ((ts + tr) * 1e-3), rows / ((ts + tc + tr) * 1e-3)))
print('E0 =', -np.log(r[0].mean) / 4., r[0], 'Q=%.2f' % r.Q)
ff = integ(f, nitn=3)
print('fused result', ff['exp(-E0*T)'])
