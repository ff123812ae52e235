"""quick A/B timing of the fused ridge kernel:  python tools/quick_bench.py [N] [neval] [shifted]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vegas_b200 as vegas
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
neval = float(sys.argv[2]) if len(sys.argv) > 2 else 1e8
shifted = len(sys.argv) > 3 and sys.argv[3] == '1'
f = vegas.integrands.Ridge(8, N=N, lo=0.5 if N == 1 else 0.4, hi=0.5 if N == 1 else 0.6, shifted=shifted)
integ = vegas.Integrator(8 * [[0., 1.]], neval=neval, seed=3)
integ(f, nitn=5)
integ._timing = []
r = integ(f, nitn=4)
torch.cuda.synchronize()
kms = [ev[1].elapsed_time(ev[2]) for ev, _ in integ._timing]
tot = sum(t for _, t in integ._timing)
from vegas_b200 import _lib
pk = _lib.fp64_peak(0, 20000)[0]
print('[fp64 peak %.1f TF/s, launch %s] %s N=%d shifted=%d neval=%.0e: kernel %.2f ms/itn  %.4e samples/s  result %s' % (
    pk, integ._ctx.last_launch(), os.environ.get('VB200_LIB', 'default'), N, shifted, neval, np.mean(kms), tot / (sum(kms) * 1e-3), r))
