import os, sys, time, cProfile, pstats
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import vegas_b200 as vegas
f = vegas.integrands.GaussMix([4 * [0.5]], 100., 1013.2118364296088)
integ = vegas.Integrator([[-1., 1.]] + 3 * [[0., 1.]], neval=1e4, seed=1)
integ(f, nitn=5)
torch.cuda.synchronize()
t0 = time.perf_counter(); r = integ(f, nitn=50); torch.cuda.synchronize(); dt=(time.perf_counter()-t0)/50
print('%.3f ms/iteration' % (dt*1e3))
pr = cProfile.Profile(); pr.enable(); r = integ(f, nitn=200); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(45)
pstats.Stats(pr).sort_stats('tottime').print_stats(25)
