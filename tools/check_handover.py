"""the pre-pass left by a call's last iteration, picked up by the next call (one-call iteration path) against the
general path on the same seed; and the cost of one-iteration calls"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
import torch
import vegas_b200 as vegas
f = vegas.integrands.GaussMix([4 * [0.5]], 100., 1013.2118364296088)
out = []
for general in (False, True):
    if general:
        os.environ['VB200_NO_FAST_ITERATION'] = '1'
    integ = vegas.Integrator([[-1., 1.]] + 3 * [[0., 1.]], neval=1e4, seed=3)
    rs = [integ(f, nitn=n) for n in (3, 1, 2, 1)]
    out.append(([x.mean for r in rs for x in r.itn_results], [x.sdev for r in rs for x in r.itn_results], integ.map.grid.copy(),
                [r.sum_neval for r in rs]))
    os.environ.pop('VB200_NO_FAST_ITERATION', None)
a, b = out
assert a[3] == b[3], (a[3], b[3])
np.testing.assert_allclose(a[0], b[0], rtol=1e-11)
np.testing.assert_allclose(a[1], b[1], rtol=1e-8)
np.testing.assert_allclose(a[2], b[2], rtol=1e-11, atol=1e-14)
print('hand-over across calls == general path:', len(a[0]), 'iterations')
integ = vegas.Integrator([[-1., 1.]] + 3 * [[0., 1.]], neval=1e4, seed=1)
integ(f, nitn=5)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    integ(f, nitn=1)
torch.cuda.synchronize()
print('nitn=1: %.3f ms per call' % ((time.perf_counter() - t0) / 50 * 1e3))
