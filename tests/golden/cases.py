"""The cases of the golden fixtures: shared by make_golden.py (reference run) and the tests."""
import numpy as np


def integrand(name):
    """numpy lbatch integrands, f(x[n, D]) -> f[n] or f[n, nf]"""
    if name == 'gauss':
        return lambda x: np.exp(-100. * np.sum((x - 0.5) ** 2, axis=1)) * 1013.2118364296088
    if name == 'poly':
        return lambda x: 0.5 + x[:, 0] ** 2 + 2. * x[:, 1] ** 3
    if name == 'osc':
        return lambda x: np.cos(2. + x.dot(np.arange(1, x.shape[1] + 1) * 0.7))
    if name == 'vec':
        def f(x):
            g = np.exp(-30. * np.sum((x - 0.4) ** 2, axis=1))
            return np.stack([g, g * x[:, 0], x[:, 1] ** 2], axis=1)
        return f
    if name == 'two_axes':
        return lambda x: np.exp(-100. * ((x[:, 0] - 0.5) ** 2 + (x[:, 1] - 0.3) ** 2)) * (1. + 0.01 * x[:, 2] + 0.01 * x[:, 3])
    if name == 'const':
        return lambda x: 7. * np.ones(x.shape[0])
    raise KeyError(name)


CASES = {
    # name: limits, integrand, Integrator kwargs, iterations, uniform seed
    'gauss4': dict(limits=[[-1., 1.]] + 3 * [[0., 1.]], f='gauss', kw=dict(neval=3000), nitn=3, seed=11),
    'poly2_beta0': dict(limits=2 * [[0., 2.]], f='poly', kw=dict(neval=1500, beta=0.0), nitn=3, seed=12),
    'osc3_errors': dict(limits=3 * [[0., 1.]], f='osc', kw=dict(neval=2000, adapt_to_errors=True), nitn=3, seed=13),
    'vec2': dict(limits=[[0., 1.], [0., 2.]], f='vec', kw=dict(neval=2500, alpha=0.3), nitn=3, seed=14),
    'vec2_nocorr': dict(limits=[[0., 1.], [0., 2.]], f='vec', kw=dict(neval=2500, correlate_integrals=False),
                        nitn=2, seed=15),
    'gauss2_batches': dict(limits=2 * [[0., 1.]], f='gauss', kw=dict(neval=4000, min_neval_batch=700), nitn=3, seed=16),
    'gauss2_bigcubes': dict(limits=2 * [[0., 1.]], f='gauss', kw=dict(neval=3000, nstrat=[3, 2]), nitn=3, seed=17),
    'const1': dict(limits=[[0., 1.]], f='const', kw=dict(neval=100, alpha=0.0), nitn=3, seed=18),
    'gauss3_clamp': dict(limits=3 * [[0., 1.]], f='gauss', kw=dict(neval=4000, max_neval_hcube=15), nitn=3, seed=20),
    'gauss2_frac50': dict(limits=2 * [[0., 1.]], f='gauss', kw=dict(neval=3000, neval_frac=0.5), nitn=3, seed=21),
    'gauss3_uniform': dict(limits=3 * [[0., 1.]], f='gauss', kw=dict(neval=5000, uniform_nstrat=True), nitn=3, seed=22),
    'gauss2_maxinc': dict(limits=2 * [[0., 1.]], f='gauss', kw=dict(neval=3000, maxinc_axis=40, alpha=1.2), nitn=3, seed=23),
    'gauss3_noadapt': dict(limits=3 * [[0., 1.]], f='gauss', kw=dict(neval=2000, adapt=False), nitn=2, seed=19),
}


# vegas.restratify (src/vegas/__init__.py:1313-1592): adapt for nitn_adapt iterations, then restratify
RESTRATIFY = {
    'plain': dict(limits=4 * [[0., 1.]], f='two_axes', kw=dict(neval=20000), nitn_adapt=3, nitn=2, ndy=5, opt={}, seed=31),
    'six_dims': dict(limits=6 * [[0., 1.]], f='two_axes', kw=dict(neval=40000), nitn_adapt=3, nitn=1, ndy=8, opt={}, seed=33),
    'uneven': dict(limits=4 * [[0., 1.]], f='two_axes', kw=dict(neval=20000, nstrat=[12, 2, 5, 3]), nitn_adapt=3, nitn=1, ndy=5,
                   opt=dict(below_avg_nstrat=1), seed=34),
    'damped': dict(limits=4 * [[0., 1.]], f='two_axes', kw=dict(neval=20000), nitn_adapt=3, nitn=1, ndy=4,
                   opt=dict(gamma=0.5, below_avg_nstrat=2), seed=32),
}


# ---- PDFIntegrator (tests/golden/make_golden_pdf.py -> ref_pdf.npz; replayed by tests/test_gpu_golden.py)
PDF_CASES = {
    'corr3': dict(mean=[1.0, 2.0, -0.5], cov=[[1.0, 0.3, -0.2], [0.3, 4.0, 0.5], [-0.2, 0.5, 0.25]], scale=1.0, limit=100.,
                  adapt_to_pdf=True, seed=321, nitn=3, kw=dict(neval=3000)),
    'corr3_fpdf': dict(mean=[1.0, 2.0, -0.5], cov=[[1.0, 0.3, -0.2], [0.3, 4.0, 0.5], [-0.2, 0.5, 0.25]], scale=1.5, limit=20.,
                       adapt_to_pdf=False, seed=322, nitn=2, kw=dict(neval=2000, alpha=0.2)),
}


def pdf_f(p):
    """dictionary-valued lbatch function of the parameters p[i, d]"""
    import numpy as np
    return dict(a=p[:, 0] + p[:, 1], b=np.stack([p[:, 1] * p[:, 2], p[:, 0] ** 2], axis=1))


# ----------------------------------------------------------------------------- Integrator.settings() strings
# (limits are given as a recipe so that the same file describes them to the reference and to vegas_b200)
SETTINGS = {
    'flat2': dict(limits=('list', [[0., 1.], [-1., 1.]]), kw=dict(neval=254, nitn=123, neval_frac=0.75)),
    'flat1_grid': dict(limits=('list', [[0., 2.]]), kw=dict(neval=1000), ngrid=4),
    'flat25_two_columns': dict(limits=('list', [[-0.5 * d, 1. + d] for d in range(25)]), kw=dict(neval=4e4, nitn=7)),
    'flat21_odd_split': dict(limits=('list', [[0., 1. + d] for d in range(21)]), kw=dict(neval=1e5, rtol=1e-3, atol=1e-9)),
    'dict_mixed': dict(limits=('dict', [('x', [[0., 1.]]), ('y', [-1., 1.]), ('zz', [[[0., 3.], [1., 2.]]])]),
                       kw=dict(neval=2000, alpha=0.2, beta=0.5)),
    'index_array': dict(limits=('list', [[[0., 1.]], [[-1., 1.]]]), kw=dict(neval=500)),
    'no_adapt': dict(limits=('list', [[0., 1.], [0., 1.], [0., 1.]]), kw=dict(neval=1e4, adapt=False, max_neval_hcube=1e4)),
    'adapt_to_errors': dict(limits=('list', [[0., 1.], [0., 1.]]), kw=dict(neval=1e3, adapt_to_errors=True), ngrid=2),
    'beta0': dict(limits=('list', [[0., 1.], [0., 1.]]), kw=dict(neval=1e3, beta=0., min_neval_batch=2000)),
    'nstrat': dict(limits=('list', [[0., 1.], [0., 1.], [0., 1.]]), kw=dict(nstrat=[6, 2, 1], neval_frac=0.5)),
    'tiny_numbers': dict(limits=('list', [[1.23456789e-7, 9.87654321e5], [-3.3333333, 7.7777777]]), kw=dict(neval=1e6)),
}


def settings_limits(recipe):
    import collections
    kind, val = recipe
    return val if kind == 'list' else collections.OrderedDict(val)
