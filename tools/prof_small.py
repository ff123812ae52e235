"""where a small-problem iteration goes (BASELINE config 1: 4-D Gaussian, neval=1e4, one-call iterations):
wall time per iteration, the split launch / overlapped bookkeeping / wait / rest, and a cProfile table"""
import os, sys, time, cProfile, pstats
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import vegas_b200 as vegas
from vegas_b200 import _lib
f = vegas.integrands.GaussMix([4 * [0.5]], 100., 1013.2118364296088)
integ = vegas.Integrator([[-1., 1.]] + 3 * [[0., 1.]], neval=1e4, seed=1)
integ(f, nitn=5)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); r = integ(f, nitn=200); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 200
    print('%.4f ms/iteration' % (dt * 1e3))
acc = dict(begin=0., end=0.)
b0, e0 = _lib.Context.iteration_begin, _lib.Context.iteration_end
def begin(self, *a):
    t = time.perf_counter(); b0(self, *a); acc['begin'] += time.perf_counter() - t
def end(self, *a):
    t = time.perf_counter(); e0(self, *a); acc['end'] += time.perf_counter() - t
_lib.Context.iteration_begin, _lib.Context.iteration_end = begin, end
t0 = time.perf_counter(); r = integ(f, nitn=200); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('per iteration: total %.1f us = launch %.1f + wait %.1f + Python (incl. overlapped bookkeeping) %.1f'
      % (dt / 200 * 1e6, acc['begin'] / 200 * 1e6, acc['end'] / 200 * 1e6, (dt - acc['begin'] - acc['end']) / 200 * 1e6))
os.environ['VB200_NO_DEFER'] = '1'
acc.update(begin=0., end=0.)
t0 = time.perf_counter(); r = integ(f, nitn=200); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print('not overlapped: total %.1f us = launch %.1f + wait %.1f + Python %.1f'
      % (dt / 200 * 1e6, acc['begin'] / 200 * 1e6, acc['end'] / 200 * 1e6, (dt - acc['begin'] - acc['end']) / 200 * 1e6))
del os.environ['VB200_NO_DEFER']
_lib.Context.iteration_begin, _lib.Context.iteration_end = b0, e0
pr = cProfile.Profile(); pr.enable(); r = integ(f, nitn=200); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(30)
