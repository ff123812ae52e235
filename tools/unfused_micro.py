"""kernel-level timing of the callback path's two kernels on one batch (no callback in between):
   python tools/unfused_micro.py [pathint|gauss8] [neval] [max_batch]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vegas_b200 as vegas
from vegas_b200 import _lib

what = sys.argv[1] if len(sys.argv) > 1 else 'pathint'
neval = float(sys.argv[2]) if len(sys.argv) > 2 else 1e8
mb = int(float(sys.argv[3])) if len(sys.argv) > 3 else 1 << 23
F = vegas.integrands
if what == 'pathint':
    f = F.PathIntegral(T=4., ndT=10, x0list=np.linspace(0, 2., 6)); limits = f.region; kw = dict(alpha=0.1); nf = 7
else:
    f = F.Ridge(8, N=1, lo=0.5, hi=0.5); limits = 8 * [[0., 1.]]; kw = {}; nf = 1
if os.environ.get('MAXNH'):
    kw['max_neval_hcube'] = int(os.environ['MAXNH'])
integ = vegas.Integrator(limits, neval=neval, seed=3, **kw)
integ(f, nitn=4)                      # adapt with the fused kernel
ctx, _ = integ._engine()
total, nmax, adaptive = integ._plan(ctx)
batches = integ._batches(ctx, mb)
c0, c1, rows = batches[len(batches) // 2]
dim = integ.dim
dev = ctx.device
x = torch.empty((rows, dim), dtype=torch.float64, device=dev); wgt = torch.empty(rows, dtype=torch.float64, device=dev)
bins = torch.empty((rows, dim), dtype=torch.int16, device=dev)
fx = torch.rand((rows, nf), dtype=torch.float64, device=dev)
hs = int(integ.map.inc.shape[1])
acc = torch.zeros(nf + nf * (nf + 1) // 2 + 1, dtype=torch.float64, device=dev)
sum_f = torch.zeros((dim, hs), dtype=torch.float64, device=dev); n_f = torch.zeros((dim, hs), dtype=torch.int64, device=dev)
status = torch.zeros(1, dtype=torch.int32, device=dev)
sig = integ._sigf_dev.clone()
flags = integ._flags(nf)

def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

print('%s: batch of %d rows (%d chunks), dim %d nf %d, range %s' % (what, rows, c1 - c0, dim, nf, list(integ.neval_hcube_range)))
t = timeit(lambda: ctx.sample(7, c0, c1, x, wgt));                 print('sample x,wgt        %.3f ms  %.3e rows/s  %.0f GB/s' % (t, rows / t * 1e3, rows * (8 * dim + 8) / t / 1e6))
t = timeit(lambda: ctx.sample(7, c0, c1, x, wgt, bins=bins));      print('sample x,wgt,bins   %.3f ms  %.3e rows/s  %.0f GB/s' % (t, rows / t * 1e3, rows * (10 * dim + 8) / t / 1e6))
xt = torch.empty((dim, rows), dtype=torch.float64, device=dev)
t = timeit(lambda: ctx.sample(7, c0, c1, xt, wgt, transposed=True)); print('sample x^T,wgt      %.3f ms  %.3e rows/s' % (t, rows / t * 1e3))
y = torch.empty_like(x)
t = timeit(lambda: ctx.sample(7, c0, c1, x, wgt, y=y));            print('sample x,wgt,y (generic kernel) %.3f ms  %.3e rows/s' % (t, rows / t * 1e3))
for name, b, fl in (('replay', None, flags), ('bins', bins, flags), ('no training', None, flags & ~_lib.TRAIN),
                    ('no train/corr', None, flags & ~_lib.TRAIN & ~_lib.CORRELATE), ('bins, no corr', bins, flags & ~_lib.CORRELATE),
                    ('bins, no sigf', bins, flags & ~_lib.UPDATE_SIGF)):
    t = timeit(lambda: ctx.reduce(7, integ.beta, fl, c0, c1, fx, nf, wgt, sig, acc, sum_f, n_f, hs, status, bins=b))
    print('reduce %-12s %.3f ms  %.3e rows/s  %.0f GB/s   launch %s' % (name, t, rows / t * 1e3, rows * (8 * nf + 8 + (2 * dim if b is not None else 0)) / t / 1e6, ctx.last_launch()))
