"""one-GPU throughput of the other BASELINE.json configurations (fused kernels):
   python tools/cfg_bench.py [genz|peaks20|pathint|all] [neval]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vegas_b200 as vegas

what = sys.argv[1] if len(sys.argv) > 1 else 'all'
F = vegas.integrands
rng = np.random.default_rng(0x5eed + 3)
CFG = {
    'genz': (lambda: F.Genz('product_peak', 2 + 3 * rng.random(10), rng.random(10)), 10 * [[0., 1.]], dict(neval=1e9, max_mem=1e10)),
    'genz_osc': (lambda: F.Genz('oscillatory', rng.random(10), rng.random(10)), 10 * [[0., 1.]], dict(neval=1e9, max_mem=1e10)),
    'peaks20': (lambda: F.GaussMix([5 * [c] + 15 * [0.45] for c in (.23, .39, .74)], 100., 356047712484621.56), 20 * [[0., 1.]],
                dict(neval=5e8, nstrat=5 * [30] + 15 * [1], max_mem=1e10)),
    'peaks20_full': (lambda: F.GaussMix([5 * [c] + 15 * [0.45] for c in (.23, .39, .74)], 100., 356047712484621.56), 20 * [[0., 1.]],
                     dict(neval=1e10, nstrat=5 * [60] + 15 * [1], max_mem=1e11)),
    'pathint': (lambda: F.PathIntegral(T=4., ndT=10, x0list=np.linspace(0, 2., 6)), 10 * [[-np.pi / 2, np.pi / 2]],
                dict(neval=1e8, alpha=0.1)),
}
for name, (mk, limits, kw) in CFG.items():
    if what != name and not (what == 'all' and name != 'peaks20_full'):
        continue
    if len(sys.argv) > 2:
        kw['neval'] = float(sys.argv[2])
    f = mk()
    if os.environ.get('CFG_ALPHA'):
        kw['alpha'] = float(os.environ['CFG_ALPHA'])
    integ = vegas.Integrator(limits, seed=5, **kw)
    integ(f, nitn=int(os.environ.get('CFG_ADAPT', 5)))
    integ._timing = []
    r = integ(f, nitn=3)
    torch.cuda.synchronize()
    kms = [ev[1].elapsed_time(ev[2]) for ev, _ in integ._timing]
    tot = sum(t for _, t in integ._timing)
    r0 = r if not hasattr(r, 'keys') else r['exp(-E0*T)']
    print('%-8s neval=%.0e nhcube=%d range=%s launch=%s: kernel %.2f ms/itn  %.4e samples/s  result %s Q=%.2f' % (
        name, kw['neval'], integ.nhcube, list(integ.neval_hcube_range), integ._ctx.last_launch(), np.mean(kms),
        tot / (sum(kms) * 1e-3), r0, r.Q), flush=True)
