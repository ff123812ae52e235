"""Parity harness: drive the CUDA engine and the CPU oracle through the same iterations with
the same Philox uniforms and compare them piece by piece.

The oracle (``oracle/oracle.py`` over ``oracle/vegas_oracle.c``) is the checker only."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _oracle():
    from oracle import oracle
    return oracle


def _as_lbatch(f):
    """numpy twin returning f[n, nf]"""
    if hasattr(f, 'eval_array'):
        return f.eval_array
    return lambda x: np.asarray(f(x), dtype=float).reshape(x.shape[0], -1)


def run_engine_iterations(limits, f, nitn, neval, seed, fused=True, **kw):
    """list of per-iteration dicts from the CUDA engine (state BEFORE map.adapt, plus the sigf and
    grid the next iteration starts from)"""
    import torch
    import vegas_b200 as vegas
    integ = vegas.Integrator(limits, neval=neval, seed=seed, fused=fused, **kw)
    out = []
    integ._trace = lambda rec: out.append(rec)
    for i in range(nitn):
        ctx, _ = integ._engine()
        sigf_in = integ.sigf.copy()
        sum_sigf_in = float(integ.sum_sigf)
        grid_in = integ.map.grid.copy()
        nh = torch.zeros(integ._nlocal, dtype=torch.int32, device=ctx.device)
        integ._plan(ctx, nh)
        integ(f, nitn=1)
        rec = out[-1]
        rec.update(neval_hcube=nh.cpu().numpy().astype(np.int64), sigf_in=sigf_in, sum_sigf_in=sum_sigf_in,
                   grid_in=grid_in, sigf_out=integ.sigf.copy(), grid_out=integ.map.grid.copy(),
                   ninc=np.array(integ.map.ninc), nstrat=np.array(integ.nstrat), launch=integ._ctx.last_launch(),
                   range=tuple(int(v) for v in integ.neval_hcube_range))
    return out


def run_oracle_iterations(limits, f, nitn, neval, seed, engine=None, **kw):
    """the same iterations on the CPU oracle.  If ``engine`` (the list from
    run_engine_iterations) is given, the oracle starts every iteration from the ENGINE's sigf /
    sum_sigf / grid, so that each iteration is compared on identical inputs (the north_star's
    'bit-exact given the same sigf')."""
    O = _oracle()
    kw = dict(kw)
    kw.pop('fused', None)
    v = O.Vegas(limits, neval=neval, **kw)
    fl = _as_lbatch(f)
    out = []
    for i in range(nitn):
        itn = engine[i]['itn'] if engine is not None else i + 1
        if engine is not None:
            if len(v.sigf):
                v.sigf = engine[i]['sigf_in'].copy()
                v.sum_sigf = engine[i]['sum_sigf_in']
            g = engine[i]['grid_in']
            v.map.grid = g[:, :v.map.grid.shape[1]].copy()
        nh_all, rng = v.allocation()
        mean, var = v.iterate(fl, lambda h0, nh: O.philox_uniforms(seed, itn, v.dim, h0, nh))
        rec = dict(itn=itn, mean=mean, var=var, sum_sigf=float(v.sum_sigf), last_neval=v.last_neval,
                   neval_hcube=nh_all, sigf_out=v.sigf.copy(),
                   sum_f=None if v.map.sum_f is None else v.map.sum_f.copy(),
                   n_f=None if v.map.n_f is None else v.map.n_f.copy(), range=tuple(int(r) for r in rng))
        v.adapt_map()
        rec['grid_out'] = v.map.grid.copy()
        out.append(rec)
    return out


def _close(a, b, rtol, what, atol=0.0):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    scale = np.maximum(np.abs(a), np.abs(b))
    err = np.abs(a - b)
    bad = err > rtol * scale + atol
    assert not bad.any(), '%s: max rel err %.3e at %s (engine %r, oracle %r)' % (
        what, float((err / np.maximum(scale, 1e-300)).max()), np.argwhere(bad)[:3].tolist(),
        a[bad][:3], b[bad][:3])


def compare_iterations(eng, ora, rtol=1e-12, var_rtol=None):
    """integers exact; fp64 sums within rtol"""
    var_rtol = rtol if var_rtol is None else var_rtol
    for i, (e, o) in enumerate(zip(eng, ora)):
        tag = 'itn %d ' % i
        assert np.array_equal(e['neval_hcube'], o['neval_hcube']), tag + 'neval_hcube differs in %d cubes' % int(
            (e['neval_hcube'] != o['neval_hcube']).sum())
        assert e['last_neval'] == o['last_neval'], tag + 'last_neval'
        _close(e['mean'], o['mean'], rtol, tag + 'mean')
        ov = np.asarray(o['var'])
        _close(np.asarray(e['var']).reshape(ov.shape), ov, var_rtol, tag + 'var', atol=1e-300)
        if len(o['sigf_out']) and (e['flags'] & 1):          # sigf is only updated when adaptive_strat
            _close(e['sum_sigf'], o['sum_sigf'], rtol, tag + 'sum_sigf')
            _close(e['sigf_out'], o['sigf_out'], 1e-8, tag + 'sigf', atol=1e-300)   # per-cube variances are ill-conditioned (cancellation in few-sample cubes)
        if o['n_f'] is not None:
            nb = o['n_f'].shape[1]
            cnt = np.rint(o['n_f']).astype(np.int64)
            assert np.array_equal(e['n_f'][:, :nb], cnt), tag + 'training counts n_f differ'
            _close(e['sum_f'][:, :nb], o['sum_f'], rtol, tag + 'sum_f', atol=1e-300)
        ge, go = e['grid_out'], o['grid_out']
        nb = min(ge.shape[1], go.shape[1])
        for d in range(ge.shape[0]):
            n = int(e['ninc'][d]) + 1
            _close(ge[d, :n], go[d, :n], 1e-10, tag + 'adapted grid axis %d' % d, atol=1e-14)
