"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference
(oracle/_ref = /root/reference/src/vegas/_vegas.pyx compiled by oracle/Makefile, imported with the
test-only gvar stand-in).  Run in the build container (needs /root/reference):

    make -C oracle ref && python tests/golden/make_golden.py

Each case drives ``vegas.Integrator`` for a few iterations with uniforms injected through the
reference's own ``ran_array_generator`` hook (numpy default_rng(seed), one call per batch), and
records through the ``analyzer`` hook what the reference computed: strata/increment integers,
per-iteration mean / covariance, ``sigf``, ``sum_sigf``, ``last_neval``, ``neval_hcube_range`` and the
adapted grid.  tests/test_oracle_golden.py replays the same uniform stream through the CPU oracle.
Also records map / invmap / jac1d / adapt outputs of ``AdaptiveMap`` on fixed inputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'gvar_shim'))
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
sys.path.insert(0, ROOT)
import gvar            # noqa: E402  (the shim)
import vegas           # noqa: E402  (the compiled reference)

from tests.golden.cases import CASES, integrand      # noqa: E402


class Recorder(object):
    def __init__(self):
        self.rows = []

    def begin(self, itn, integ):
        self.integ = integ

    def end(self, itn_result, result):
        I = self.integ
        r = np.asarray(itn_result, dtype=object).reshape(-1)
        mean = np.array([x.mean for x in r], float)
        cov = gvar.evalcov(r)
        self.rows.append(dict(
            mean=mean, cov=np.array(cov, float), sigf=np.array(I.sigf, float), sum_sigf=float(I.sum_sigf),
            last_neval=int(I.last_neval), range=np.array(I.neval_hcube_range, np.int64),
            grid=np.array(I.map.grid, float)))


def run_case(name, spec):
    gvar.ranseed(1)
    rng = np.random.default_rng(spec['seed'])
    calls = []

    def gen(shape):
        calls.append(int(shape[0]))
        return rng.random(shape)

    rec = Recorder()
    kw = dict(spec['kw'])
    integ = vegas.Integrator(spec['limits'], ran_array_generator=gen, analyzer=rec, **kw)
    f = vegas.lbatchintegrand(integrand(spec['f']))
    integ(f, nitn=spec['nitn'])
    out = dict(nstrat=np.array(integ.nstrat, np.int64), ninc=np.array(integ.map.ninc, np.int64),
               nhcube=np.int64(integ.nhcube), min_neval_hcube=np.int64(integ.min_neval_hcube),
               batch_rows=np.array(calls, np.int64))
    for i, row in enumerate(rec.rows):
        for k, v in row.items():
            out['itn%d_%s' % (i, k)] = v
    np.savez_compressed(os.path.join(HERE, 'ref_%s.npz' % name), **out)
    print('%-14s nstrat=%s nhcube=%d batches=%d  itn0 mean=%s' % (
        name, list(integ.nstrat), integ.nhcube, len(calls), rec.rows[0]['mean']))


def run_map():
    rng = np.random.default_rng(2024)
    grid = [np.sort(np.concatenate([[0., 1.], rng.random(n - 1)])) * s + o
            for n, s, o in [(12, 1., 0.), (7, 3., -1.), (30, 1e-3, 5.)]]
    m = vegas.AdaptiveMap(grid)
    y = rng.random((400, 3))
    y[:4] = [[0, 0, 0], [1, 1, 1], [0.5, 0.25, 0.75], [1 - 1e-16, 1e-300, 0.999999999999]]
    x = np.empty_like(y)
    jac = np.empty(len(y))
    m.map(y, x, jac)
    out = dict(grid0=np.array(grid[0]), grid1=np.array(grid[1]), grid2=np.array(grid[2]), y=y, x=x, jac=jac,
               jac1d=np.array(m.jac1d(y)))
    # invmap: equal node counts on every axis (the reference searches the whole padded row,
    # _vegas.pyx:404, which is only meaningful when no row is padded)
    gridi = [np.sort(np.concatenate([[0., 1.], rng.random(19)])) * s + o for s, o in [(1., 0.), (3., -1.), (1e-3, 5.)]]
    mi = vegas.AdaptiveMap(gridi)
    xi = np.empty_like(y)
    ji = np.empty(len(y))
    mi.map(y, xi, ji)
    xi[4] = [-0.5, 2.5, 5.0005]          # below / above the region, inside
    y2 = np.empty_like(y)
    jac2 = np.empty(len(y))
    mi.invmap(xi, y2, jac2)
    out.update(gridi0=gridi[0], gridi1=gridi[1], gridi2=gridi[2], inv_x=xi, inv_y=y2, inv_jac=jac2)
    # training + adapt, three alphas, mixed ninc
    m = vegas.AdaptiveMap([[0, 2], [-1, 1]], ninc=[50, 33])
    for i, alpha in enumerate((1.5, 0.5, -1.0)):
        yt = rng.random((3000, 2))
        yt[0] = [0.0, 1.0]
        ft = rng.standard_normal(3000) ** 2 * np.exp(-3 * yt[:, 0])
        m.add_training_data(yt, ft)
        out['train%d_y' % i], out['train%d_f' % i] = yt, ft
        out['train%d_sum_f' % i], out['train%d_n_f' % i] = np.array(m.sum_f), np.array(m.n_f)
        m.adapt(alpha=alpha)
        out['train%d_grid' % i] = np.array(m.grid)
    m.adapt(ninc=[20, 7])
    out['regrid'] = np.array(m.grid)
    np.savez_compressed(os.path.join(HERE, 'ref_map.npz'), **out)
    print('map fixtures written')


if __name__ == '__main__':
    print('reference version', vegas.__version__)
    for name, spec in CASES.items():
        run_case(name, spec)
    run_map()
