"""fixed cost of one Integrator.__call__ (config 1, nitn=10 per call, as the reference's examples use it)"""
import os, sys, time, cProfile, pstats
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import vegas_b200 as vegas
f = vegas.integrands.GaussMix([4 * [0.5]], 100., 1013.2118364296088)
integ = vegas.Integrator([[-1., 1.]] + 3 * [[0., 1.]], neval=1e4, seed=1)
integ(f, nitn=5)
torch.cuda.synchronize()
for nitn in (1, 10, 100):
    t0 = time.perf_counter()
    for _ in range(20):
        r = integ(f, nitn=nitn)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print('nitn=%3d: %.3f ms per call, %.4f ms per iteration' % (nitn, dt * 1e3, dt * 1e3 / nitn))
pr = cProfile.Profile(); pr.enable()
for _ in range(50):
    r = integ(f, nitn=1)
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(50)
