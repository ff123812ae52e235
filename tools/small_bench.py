"""per-iteration latency at the reference's everyday sizes (BASELINE config 1: 4-D Gaussian, neval=1e4)
   python tools/small_bench.py"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vegas_b200 as vegas

f = vegas.integrands.GaussMix([4 * [0.5]], 100., 1013.2118364296088)
fn = vegas.lbatchintegrand(lambda x: np.exp(-100. * np.sum((x - 0.5) ** 2, axis=1)) * 1013.2118364296088)
for neval in (1e4, 1e5, 1e6):
    for name, g in (('device functor', f), ('numpy lbatch', fn)):
        integ = vegas.Integrator([[-1., 1.]] + 3 * [[0., 1.]], neval=neval, seed=1)
        integ(g, nitn=5)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = integ(g, nitn=10)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        print('neval=%.0e %-14s %.3f ms/iteration  %.3e samples/s  %s Q=%.2f' % (neval, name, dt * 1e3, r.sum_neval / 10 / dt, r, r.Q))
