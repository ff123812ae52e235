"""GPU parity tests: the CUDA engine (through the C ABI / Python host layer) against the CPU oracle
on the same seeded inputs.  Integers bit-exact; fp64 within the stated relative tolerance."""
import numpy as np
import pytest

from tests.parity import run_engine_iterations, run_oracle_iterations, compare_iterations, _oracle

pytestmark = pytest.mark.gpu

RTOL = 1e-12          # north_star: fp64 results of deterministic pieces within 1e-12 relative


def _vegas():
    import vegas_b200
    return vegas_b200


# ----------------------------------------------------------------------------- AdaptiveMap
def test_map_known_answers():
    """reference tests/test_vegas.py:76-112 (map / jac / invmap on a 2-increment grid)"""
    vegas = _vegas()
    m = vegas.AdaptiveMap(grid=[[0, 1, 3], [-2, 0, 6]])
    y = np.array([[0, 0], [0.25, 0.25], [0.5, 0.5], [0.75, 0.75], [1.0, 1.0]])
    x = np.empty_like(y)
    jac = np.empty(len(y))
    m.map(y, x, jac)
    np.testing.assert_allclose(x, [[0, -2], [0.5, -1], [1, 0], [2, 3], [3, 6]], rtol=1e-15, atol=0)
    np.testing.assert_allclose(jac, [8, 8, 48, 48, 48], rtol=1e-15)
    np.testing.assert_allclose(m(y), x)
    np.testing.assert_allclose(m.jac(y), jac)
    np.testing.assert_allclose(m.jac1d(y), [[2, 4], [2, 4], [4, 12], [4, 12], [4, 12]])
    y2 = np.empty_like(y)
    jac2 = np.empty(len(y))
    m.invmap(x, y2, jac2)
    np.testing.assert_allclose(y2, y, rtol=1e-15, atol=1e-16)
    np.testing.assert_allclose(jac2, jac)


def test_map_vs_oracle_random():
    vegas = _vegas()
    O = _oracle()
    rng = np.random.default_rng(5)
    grid = [np.sort(np.concatenate([[0., 1.], rng.random(n - 1)])) * s + o
            for n, s, o in [(100, 1., 0.), (37, 3., -1.), (1000, 1e-3, 5.)]]
    m = vegas.AdaptiveMap(grid)
    om = O.Map(grid)
    y = rng.random((200000, 3))
    y[:5] = [[0, 0, 0], [1, 1, 1], [0.5, 0.5, 0.5], [1 - 1e-16, 1e-300, 0.999999999999], [0.01, 0.37, 1.0]]
    x = np.empty_like(y)
    jac = np.empty(len(y))
    m.map(y, x, jac)
    xo, jo = om.map(y)
    assert np.array_equal(x, xo) and np.array_equal(jac, jo)          # same ops, same order: bit-exact
    assert np.array_equal(m.jac1d(y), om.jac1d(y))
    y2 = np.empty_like(y)
    j2 = np.empty(len(y))
    m.invmap(x, y2, j2)
    yo, jo2 = om.invmap(x)
    assert np.array_equal(y2, yo) and np.array_equal(j2, jo2)


def test_add_training_data_and_adapt_vs_oracle():
    vegas = _vegas()
    O = _oracle()
    rng = np.random.default_rng(11)
    m = vegas.AdaptiveMap([[0, 2], [-1, 1]], ninc=[50, 33])
    om = O.Map([[0, 2], [-1, 1]], ninc=[50, 33])
    for alpha in (1.5, 0.5, -1.0):
        y = rng.random((50000, 2))
        y[0] = [0.0, 1.0]              # boundary points are skipped (pyx:460)
        f = rng.standard_normal(50000) ** 2 * np.exp(-3 * y[:, 0])
        m.add_training_data(y, f)
        om.add_training_data(y, f)
        assert np.array_equal(np.rint(m.n_f), np.rint(om.n_f))
        np.testing.assert_allclose(m.sum_f, om.sum_f, rtol=1e-12)
        m.adapt(alpha=alpha)
        om.adapt(alpha=alpha)
        for d in range(2):
            np.testing.assert_allclose(m.grid[d, :m.ninc[d] + 1], om.grid[d, :om.ninc[d] + 1], rtol=1e-11, atol=1e-15)


# ----------------------------------------------------------------------------- RNG + allocation
def test_adapt_to_samples():
    """AdaptiveMap.adapt_to_samples (pyx:728-804; reference test tests:131-140): invmap, training and
    adapt on the device/host pair reproduce the oracle's loop, and the map concentrates where the
    samples are"""
    vegas = _vegas()
    O = _oracle()
    rng = np.random.default_rng(5)
    x = rng.normal(0, .1, (1000, 2))
    Fx = np.exp(-np.sum(x ** 2, axis=1) * 100 / 2)
    m1 = vegas.AdaptiveMap([[0, 2], [0, 1]])
    m1.adapt_to_samples(x, Fx, nitn=5)
    assert list(m1.ninc) == [1000, 1000]
    mo = O.Map([[0., 2.], [0., 1.]])
    mo.adapt(ninc=100)
    for _ in range(5):
        y, jac = mo.invmap(x)
        mo.add_training_data(y, (jac * Fx) ** 2)
        mo.adapt(alpha=1.0, ninc=100)
    mo.adapt(ninc=1000)
    np.testing.assert_allclose(m1.grid, mo.grid[:, :m1.grid.shape[1]], rtol=1e-10, atol=1e-14)
    m1.adapt(ninc=2)
    np.testing.assert_allclose(np.asarray(m1.grid), [[0., 0.071, 2.0], [0., 0.073, 1.]], rtol=0.4)


def test_philox_uniforms_bit_exact():
    import torch
    vegas = _vegas()
    O = _oracle()
    integ = vegas.Integrator(5 * [[0, 1]], neval=3000, seed=987654321012345)
    ctx, _ = integ._engine()
    nh = torch.zeros(integ._nlocal, dtype=torch.int32, device=ctx.device)
    total, nmax, _ = integ._plan(ctx, nh)
    u = torch.empty((total, 5), dtype=torch.float64, device=ctx.device)
    ctx.uniforms(7, 0, integ._nchunks, u)
    uo = O.philox_uniforms(987654321012345, 7, 5, 0, nh.cpu().numpy().astype(np.int64))
    assert np.array_equal(u.cpu().numpy(), uo)
    assert uo.min() >= 0 and uo.max() < 1


@pytest.mark.parametrize('nh,neval,mx', [(100000, 1e6, 50000), (1000, 1e7, 200), (777, 5e3, 50000)])
def test_allocation_bit_exact(nh, neval, mx):
    """neval_hcube from sigf: integer results bit-exact against the reference formula
    (pyx:1692-1706) for heavy-tailed sigf, including the max_neval_hcube clamp"""
    import torch
    from vegas_b200 import _lib
    O = _oracle()
    rng = np.random.default_rng(nh)
    sigf = np.abs(rng.standard_cauchy(nh)) ** 0.75
    sigf[::97] = 0.0
    neval_sigf = 0.75 * neval / sigf.sum()
    ctx = _lib.Context()
    ctx.set_strata([nh], _lib.CHUNK)
    sd = torch.from_numpy(sigf).cuda()
    out = torch.zeros(nh, dtype=torch.int32, device='cuda')
    total, nmin, nmax, nchunks = ctx.plan(sd, neval_sigf, 2, mx, 0, out)
    ref = np.empty(nh, np.int64)
    rng2 = np.array([2, 2], np.int64)
    tot = O.lib().vo_alloc_neval(O._dp(sigf), nh, neval_sigf, 2, mx, O._ip(ref), O._ip(rng2))
    assert np.array_equal(out.cpu().numpy(), ref)
    assert total == tot and nmax == ref.max() and nmin == ref.min()
    off = ctx.chunk_offsets(nchunks + 1)
    assert off[0] == 0 and off[-1] == tot
    assert np.array_equal(np.diff(off), np.add.reduceat(ref, np.arange(0, nh, _lib.CHUNK)))


# ----------------------------------------------------------------------------- full iterations
def _cases():
    import vegas_b200 as vegas
    F = vegas.integrands
    rng = np.random.default_rng(3)
    return {
        'poly2': (2 * [[0., 2.]], F.Poly(0.5, [1.0, 2.0], [2, 3]), dict(neval=4000)),
        'gauss4': ([[-1., 1.]] + 3 * [[0., 1.]], F.GaussMix([4 * [0.5]], 100., 1013.2118364296088), dict(neval=10000)),
        # the settings of the extra golden fixtures (tests/golden/cases.py): allocation clamp, neval_frac,
        # uniform_nstrat, few increments with strong damping
        'gauss3_clamp': (3 * [[0., 1.]], F.GaussMix([3 * [0.5]], 100., 1.0), dict(neval=4000, max_neval_hcube=15)),
        'gauss2_frac50': (2 * [[0., 1.]], F.GaussMix([2 * [0.5]], 100., 1.0), dict(neval=3000, neval_frac=0.5)),
        'gauss3_uniform': (3 * [[0., 1.]], F.GaussMix([3 * [0.5]], 100., 1.0), dict(neval=5000, uniform_nstrat=True)),
        'gauss2_maxinc': (2 * [[0., 1.]], F.GaussMix([2 * [0.5]], 100., 1.0), dict(neval=3000, maxinc_axis=40, alpha=1.2)),
        'ridge8': (8 * [[0., 1.]], F.Ridge(8, N=17), dict(neval=60000)),
        'ridge8_shifted': (8 * [[0., 1.]], F.Ridge(8, N=21, shifted=True), dict(neval=60000)),
        'ridge6_pad': (6 * [[0., 1.]], F.Ridge(6, N=9), dict(neval=20000)),
        'ridge4_nomap': (4 * [[0., 1.]], F.Ridge(4, N=5), dict(neval=5000, alpha=0.0)),
        'genz10_pp': (10 * [[0., 1.]], F.Genz('product_peak', 2 + 3 * rng.random(10), rng.random(10)), dict(neval=50000)),
        'genz10_osc': (10 * [[0., 1.]], F.Genz('oscillatory', rng.random(10), rng.random(10)), dict(neval=50000, beta=0.0)),
        'genz3_corner': (3 * [[0., 1.]], F.Genz('corner_peak', 1 + rng.random(3), rng.random(3)), dict(neval=8000, adapt_to_errors=True)),
        'peaks20': (20 * [[0., 1.]], F.GaussMix([5 * [c] + 15 * [0.45] for c in (.23, .39, .74)], 100., 356047712484621.56),
                    dict(neval=200000, nstrat=5 * [6] + 15 * [1])),
        'pathint10': (10 * [[-np.pi / 2, np.pi / 2]], F.PathIntegral(T=4., ndT=10, x0list=np.linspace(0, 2., 6)),
                      dict(neval=40000, alpha=0.1)),
        'pathint8_nocorr': (8 * [[-np.pi / 2, np.pi / 2]], F.PathIntegral(T=4., ndT=8, x0list=np.linspace(0, 2., 6)),
                            dict(neval=30000, correlate_integrals=False)),
    }


CASES = ['poly2', 'gauss4', 'gauss3_clamp', 'gauss2_frac50', 'gauss3_uniform', 'gauss2_maxinc', 'ridge8', 'ridge8_shifted', 'ridge6_pad', 'ridge4_nomap', 'genz10_pp', 'genz10_osc', 'genz3_corner', 'peaks20',
         'pathint10', 'pathint8_nocorr']


@pytest.mark.parametrize('name', CASES)
def test_fused_iterations_vs_oracle(name):
    """three adapting iterations of the fused kernel vs the oracle on the same uniforms"""
    limits, f, kw = _cases()[name]
    eng = run_engine_iterations(limits, f, nitn=3, seed=1000 + len(name), **kw)
    ora = run_oracle_iterations(limits, f, nitn=3, seed=1000 + len(name), engine=eng, **kw)
    compare_iterations(eng, ora, rtol=1e-11 if name.startswith('pathint') else RTOL, var_rtol=1e-10)


@pytest.mark.parametrize('name', ['poly2', 'gauss4', 'genz10_pp', 'pathint10', 'genz3_corner'])
def test_unfused_iterations_vs_oracle(name):
    """sample -> host numpy integrand -> reduce (the callback path) vs the oracle"""
    limits, f, kw = _cases()[name]
    eng = run_engine_iterations(limits, f, nitn=3, seed=77, fused=False, **kw)
    ora = run_oracle_iterations(limits, f, nitn=3, seed=77, engine=eng, **kw)
    compare_iterations(eng, ora, rtol=RTOL, var_rtol=1e-10)


@pytest.mark.parametrize('name', ['gauss4', 'pathint10', 'torch_gauss4'])
def test_device_callback_route_vs_oracle(name):
    """the on-device callback route (sampler -> @devicebatchintegrand on the HBM buffers -> reduce; samples and
    integrand values never leave the GPU) vs the oracle on the same uniforms: a library functor run on the buffers
    (``DeviceIntegrand.device_twin`` = vb200_eval_integrand), and a callback written in torch"""
    import torch
    vegas = _vegas()
    if name == 'torch_gauss4':
        limits, f, kw = _cases()['gauss4']

        @vegas.devicebatchintegrand
        def fdev(x):
            assert x.is_cuda and x.dtype == torch.float64
            return torch.exp(-100. * ((x - 0.5) ** 2).sum(dim=1)) * 1013.2118364296088

        fnp = vegas.lbatchintegrand(lambda x: np.exp(-100. * np.sum((x - 0.5) ** 2, axis=1)) * 1013.2118364296088)
    else:
        limits, fnp, kw = _cases()[name]
        fdev = fnp.device_twin(len(limits))
    eng = run_engine_iterations(limits, fdev, nitn=3, seed=78, fused=False, **kw)
    ora = run_oracle_iterations(limits, fnp, nitn=3, seed=78, engine=eng, **kw)
    compare_iterations(eng, ora, rtol=1e-11 if name.startswith('pathint') else RTOL, var_rtol=1e-10)


def test_unfused_philox_replay_equals_bins():
    """reduce kernel: training bins re-derived from the Philox counter (train_bins=False) == bins
    handed over by the sampler"""
    limits, f, kw = _cases()['pathint10']
    a = run_engine_iterations(limits, f, nitn=2, seed=5, fused=False, train_bins=True, **kw)
    b = run_engine_iterations(limits, f, nitn=2, seed=5, fused=False, train_bins=False, **kw)
    for ra, rb in zip(a, b):
        assert np.array_equal(ra['neval_hcube'], rb['neval_hcube'])
        assert np.array_equal(ra['n_f'], rb['n_f'])
        np.testing.assert_allclose(ra['mean'], rb['mean'], rtol=1e-12)
        np.testing.assert_allclose(ra['sum_f'], rb['sum_f'], rtol=1e-10, atol=1e-300)


def test_fused_equals_unfused_streams():
    """the fused kernel and the unfused path consume the same samples"""
    limits, f, kw = _cases()['gauss4']
    a = run_engine_iterations(limits, f, nitn=2, seed=5, fused=True, **kw)
    b = run_engine_iterations(limits, f, nitn=2, seed=5, fused=False, **kw)
    for ra, rb in zip(a, b):
        assert np.array_equal(ra['neval_hcube'], rb['neval_hcube'])
        assert np.array_equal(ra['n_f'], rb['n_f'])
        np.testing.assert_allclose(ra['mean'], rb['mean'], rtol=1e-12)


def test_random_batch_matches_oracle():
    """Integrator.random_batch: x, y, wgt, hcube in hypercube order (pyx:1601-1634)"""
    vegas = _vegas()
    O = _oracle()
    integ = vegas.Integrator([[0, 1], [-1, 3], [0, 10]], neval=5000, seed=42, min_neval_batch=1500)
    parts = list(integ.random_batch(yield_hcube=True, yield_y=True))
    assert len(parts) > 1
    x = np.concatenate([p[0] for p in parts]); y = np.concatenate([p[1] for p in parts])
    w = np.concatenate([p[2] for p in parts]); hc = np.concatenate([p[3] for p in parts])
    assert len(x) == integ.last_neval and np.all(np.diff(hc) >= 0)
    v = O.Vegas([[0, 1], [-1, 3], [0, 10]], neval=5000)
    nh, _ = v.allocation()
    u = O.philox_uniforms(42, 1, 3, 0, nh)
    yo = np.empty_like(u); hco = np.empty(len(u), np.int64)
    O.lib().vo_stratify(O._ip(v.nstrat), 3, 0, len(nh), O._ip(nh), O._dp(u), O._dp(yo), O._ip(hco))
    xo, jo = v.map.map(yo)
    O.lib().vo_weights(O._dp(jo), O._ip(nh), len(nh), 1. / v.nhcube)
    assert np.array_equal(hc, hco) and np.array_equal(y, yo)
    assert np.array_equal(x, xo)                 # same ops in the same order, no FMA contraction
    np.testing.assert_allclose(w, jo, rtol=1e-15)
    assert abs(w.sum() - 40.0) < 1e-9            # sum of weights = volume


# ----------------------------------------------------------------------------- integrals
def test_integrals_statistically_correct():
    """reference-style checks (tests/test_vegas.py:791-838): |mean - exact| < 5 sigma, Q > 1e-3"""
    vegas = _vegas()
    F = vegas.integrands
    rng = np.random.default_rng(8)
    cases = [
        ([[-1., 1.]] + 3 * [[0., 1.]], F.GaussMix([4 * [0.5]], 100., 1013.2118364296088), 1.0, dict(neval=20000)),
        (8 * [[0., 1.]], F.Ridge(8, N=30), None, dict(neval=400000)),
        (6 * [[0., 1.]], F.Genz('gaussian', 2 + 2 * rng.random(6), 0.3 + 0.4 * rng.random(6)), 'exact', dict(neval=100000)),
        (5 * [[0., 1.]], F.Genz('c0', 1 + 2 * rng.random(5), rng.random(5)), 'exact', dict(neval=100000)),
        (4 * [[0., 1.]], F.Genz('discontinuous', rng.random(4), 0.2 + 0.6 * rng.random(4)), 'exact', dict(neval=100000)),
        (7 * [[0., 1.]], F.Genz('corner_peak', 0.2 + rng.random(7), rng.random(7)), 'exact', dict(neval=100000)),
    ]
    for limits, f, exact, kw in cases:
        integ = vegas.Integrator(limits, seed=31, **kw)
        integ(f, nitn=6)
        r = integ(f, nitn=8)
        if exact == 'exact':
            exact = f.exact()
        if exact is None:
            from scipy.special import erf
            # ridge: mean over k of prod_d Gaussian integral over [0,1]
            x0 = f.x0
            one = 0.5 * (erf(10 * (1 - x0)) + erf(10 * x0))
            exact = float(np.mean(one ** f.dim))
        assert abs(r.mean - exact) < 5 * r.sdev, (type(f).__name__, getattr(f, 'kind', ''), r.mean, r.sdev, exact)
        assert r.Q > 1e-3, (type(f).__name__, r.Q)
        assert r.sdev < 0.02 * abs(exact)


def test_constant_and_zero_integrands():
    """reference tests/test_vegas.py:774-789, 1272-1296: EPSILON clamp and the sum_sigf == 0 reset"""
    vegas = _vegas()
    F = vegas.integrands
    integ = vegas.Integrator([[-1, 1], [0, 4]], seed=3)
    r = integ(F.Poly(2.0))
    np.testing.assert_allclose(r.mean, 16, rtol=1e-6)
    assert r.sdev < 1e-6
    integ = vegas.Integrator([(0, 1)], neval=100, alpha=0, seed=4)
    res = integ(F.Poly(7.0), nitn=4)
    assert abs(res.itn_results[0].sdev / res.itn_results[1].sdev - 1.0) < 1e-7
    assert abs(res.itn_results[0].sdev / res.sdev - 2.0) < 1e-7
    assert abs(res.mean - 7.0) < 1e-7
    res = integ(F.Poly(7.0), nitn=4, neval=1e2, adapt=False)
    assert abs(res.itn_results[0].sdev / res.sdev - 2.0) < 1e-7
    integ = vegas.Integrator([(0, 1)], seed=5)
    res = integ(F.Poly(0.0), nitn=4, neval=100)
    assert res.mean == 0.0


def test_python_integrands_through_gpu_sampler():
    """scalar / lbatch / rbatch / dict-valued python integrands (pyx:2959-3383) on GPU samples"""
    vegas = _vegas()
    integ = vegas.Integrator([[0, 1], [0, 2]], neval=4000, seed=11)

    def fs(x):
        return x[0] * x[1]

    @vegas.lbatchintegrand
    def fl(x):
        return x[:, 0] * x[:, 1]

    @vegas.rbatchintegrand
    def fr(x):
        return dict(a=x[0] * x[1], b=[x[0], x[1] ** 2])

    for f in (fs, fl):
        r = integ(f, nitn=5)
        assert abs(r.mean - 1.0) < 5 * r.sdev and r.sdev < 0.01
    r = integ(fr, nitn=5)
    assert abs(r['a'].mean - 1.0) < 5 * r['a'].sdev
    assert abs(r['b'][0].mean - 1.0) < 5 * r['b'][0].sdev
    assert abs(r['b'][1].mean - 8. / 3.) < 5 * r['b'][1].sdev

    def bad(x):
        return float('nan')

    with pytest.raises(ValueError):
        integ(bad, nitn=1)

    def boom(x):
        return 1 / 0

    with pytest.raises(ZeroDivisionError):
        integ(boom, nitn=1)


@pytest.mark.parametrize('dim', [1, 3, 25, 32])
def test_odd_and_maximum_dimensions_callback_path(dim):
    """dimensions outside the fused instantiations (odd, above 20, the ABI maximum of 32) through
    the callback path: engine == oracle on the same uniforms"""
    vegas = _vegas()
    c = np.linspace(0.3, 0.7, dim)
    twin = lambda x: np.exp(-3. * np.sum((x - c) ** 2, axis=1)) * (1. + 0.1 * x[:, 0])
    f = vegas.lbatchintegrand(lambda x: twin(x))
    kw = dict(neval=6000)
    eng = run_engine_iterations(dim * [[0., 1.]], f, nitn=2, seed=300 + dim, **kw)
    ora = run_oracle_iterations(dim * [[0., 1.]], twin, nitn=2, seed=300 + dim, engine=eng, **kw)
    compare_iterations(eng, ora, rtol=RTOL, var_rtol=1e-10)


def test_fused_falls_back_when_no_instantiation():
    """a built-in functor in a dimension its family is not compiled for (path integral, 20 time
    slices: instantiated up to 16) runs through the callback path instead of failing"""
    vegas = _vegas()
    f = vegas.integrands.PathIntegral(T=4., ndT=20, x0list=np.linspace(0, 2., 6))
    kw = dict(neval=30000, alpha=0.1)
    eng = run_engine_iterations(20 * [[-np.pi / 2, np.pi / 2]], f, nitn=2, seed=808, **kw)
    ora = run_oracle_iterations(20 * [[-np.pi / 2, np.pi / 2]], f, nitn=2, seed=808, engine=eng, **kw)
    compare_iterations(eng, ora, rtol=1e-11, var_rtol=1e-9)


def test_device_batch_callback():
    """@devicebatchintegrand: torch CUDA tensors in HBM, no host round trip"""
    import torch
    vegas = _vegas()

    @vegas.devicebatchintegrand
    def f(x):
        assert x.is_cuda and x.dtype == torch.float64
        return torch.exp(-100. * ((x - 0.5) ** 2).sum(dim=1)) * 1013.2118364296088

    integ = vegas.Integrator([[-1., 1.]] + 3 * [[0., 1.]], neval=50000, seed=2)
    integ(f, nitn=5)
    r = integ(f, nitn=5)
    assert abs(r.mean - 1.0) < 5 * r.sdev and r.sdev < 0.01


def test_large_cubes_and_giant_cube_paths():
    """cubes above VB_WARP_CUBE samples (warp reduce) and above the staging capacity (global
    scratch) agree with the oracle"""
    vegas = _vegas()
    F = vegas.integrands
    f = F.GaussMix([[0.3, 0.6]], 400., 1.0)
    for kw in (dict(neval=40000, nstrat=[4, 3]), dict(neval=30000, nstrat=[1, 2]), dict(neval=9000, nstrat=[1, 1])):
        eng = run_engine_iterations(2 * [[0., 1.]], f, nitn=2, seed=9, **kw)
        ora = run_oracle_iterations(2 * [[0., 1.]], f, nitn=2, seed=9, engine=eng, **kw)
        compare_iterations(eng, ora, rtol=1e-11, var_rtol=1e-9)


@pytest.mark.parametrize('name,fused', [('gauss4', True), ('peaks20', True), ('pathint10', True), ('gauss4', False),
                                        ('pathint10', False), ('genz3_corner', False)])
def test_split_chunks_vs_oracle(name, fused, monkeypatch):
    """work items: chunks cut into several items (forced here with a 256-sample item size; in
    production only chunks the vegas+ allocation piled > 4096 samples onto) give the same sums"""
    monkeypatch.setenv('VB200_ITEM', '256')
    limits, f, kw = _cases()[name]
    eng = run_engine_iterations(limits, f, nitn=3, seed=321, fused=fused, **kw)
    ora = run_oracle_iterations(limits, f, nitn=3, seed=321, engine=eng, **kw)
    compare_iterations(eng, ora, rtol=1e-11 if name.startswith('pathint') else RTOL, var_rtol=1e-10)


def test_split_chunks_big_cubes_and_batches(monkeypatch):
    """item splitting with cubes larger than an item / than the staging buffer, and the sampler
    (random_batch) under splitting: identical rows"""
    vegas = _vegas()
    f = vegas.integrands.GaussMix([[0.3, 0.6]], 400., 1.0)
    monkeypatch.setenv('VB200_ITEM', '256')
    for kw in (dict(neval=40000, nstrat=[4, 3]), dict(neval=200000, nstrat=[20, 30])):
        eng = run_engine_iterations(2 * [[0., 1.]], f, nitn=3, seed=9, **kw)
        ora = run_oracle_iterations(2 * [[0., 1.]], f, nitn=3, seed=9, engine=eng, **kw)
        compare_iterations(eng, ora, rtol=1e-11, var_rtol=1e-9)
    out = []
    base = vegas.Integrator(2 * [[0., 1.]], neval=200000, nstrat=[20, 30], seed=11)
    base(f, nitn=3)
    for item in ('256', '1000000'):
        monkeypatch.setenv('VB200_ITEM', item)
        integ = vegas.Integrator(base, seed=11)          # same map and sigf
        integ._itn_counter = 100
        parts = list(integ.random_batch(yield_hcube=True, yield_y=True))
        out.append([np.concatenate([p[i] for p in parts]) for i in range(4)])
    for a, b in zip(*out):
        assert np.array_equal(a, b)


# ----------------------------------------------------------------------------- restratify
@pytest.mark.parametrize('case', ['functor', 'host', 'bigcubes', 'split'])
def test_stratification_profile_vs_oracle(case, monkeypatch):
    """I and the profile dI[mu][i] of vegas.restratify (vb200_reduce + vb200_dy_profile on the callback
    path's buffers) vs the oracle's restatement of the reference's auxiliary integrand
    (src/vegas/__init__.py:1390-1419; pinned to the reference by tests/test_oracle_golden.py)"""
    vegas = _vegas()
    O = _oracle()
    from vegas_b200._restratify import stratification_profile
    from tests.golden.cases import integrand
    limits, ndy, seed, kw = 4 * [[0., 1.]], 5, 515, dict(neval=20000)
    if case == 'host':
        twin = integrand('two_axes')
        f = vegas.lbatchintegrand(lambda x: twin(x))
    else:
        f = vegas.integrands.GaussMix([[0.5, 0.3, 0.5, 0.5], [0.2, 0.7, 0.5, 0.5]], 60., 3.0)
        twin = f
    if case == 'bigcubes':
        kw = dict(neval=10000, nstrat=[4, 3, 2, 1])          # ~100 samples per cube and more: the warp-per-cube path
        ndy = 7
    if case == 'split':
        monkeypatch.setenv('VB200_ITEM', '256')
    integ = vegas.Integrator(limits, seed=seed, **kw)
    integ(f, nitn=3)
    for it in range(2):
        sigf_in, sum_sigf_in, grid_in = integ.sigf.copy(), float(integ.sum_sigf), integ.map.grid.copy()
        res = stratification_profile(integ, f, nitn=1, ndy=ndy)
        v = O.Vegas(limits, alpha=0.0, correlate_integrals=False, **kw)
        v.map.grid = grid_in[:, :v.map.grid.shape[1]].copy()
        v.sigf, v.sum_sigf = sigf_in, sum_sigf_in
        pitn = integ._itn_counter
        mean, var = v.iterate(O.profile_integrand(v.map, twin, ndy), lambda h0, nh: O.philox_uniforms(seed, pitn, 4, h0, nh))
        var = np.asarray(var).reshape(-1)
        r = res.itn_results[0]
        assert integ.last_neval == v.last_neval
        np.testing.assert_allclose(r['I'].mean, mean[0], rtol=1e-12)
        np.testing.assert_allclose(r['I'].sdev ** 2, var[0], rtol=1e-10)
        dI = np.asarray(r['dI'])
        np.testing.assert_allclose([[g.mean for g in row] for row in dI], mean[1:].reshape(4, ndy), rtol=1e-11, atol=1e-300)
        np.testing.assert_allclose([[g.sdev ** 2 for g in row] for row in dI], var[1:].reshape(4, ndy), rtol=1e-9, atol=1e-300)
        np.testing.assert_allclose(integ.sigf, v.sigf, rtol=1e-8, atol=1e-300)
        np.testing.assert_allclose(integ.sum_sigf, v.sum_sigf, rtol=1e-12)
        np.testing.assert_allclose(sum(g.mean for g in dI[0]), r['I'].mean, rtol=1e-12)      # every sample is in one bin per axis


def test_restratify_end_to_end():
    """vegas.restratify on an integrand whose structure lives on two of six axes: the strata move
    there, the hypercube count stays close, and the new integrator integrates correctly"""
    vegas = _vegas()
    f = vegas.integrands.Genz('gaussian', [12., 12., .2, .2, .2, .2], [.5, .3, .5, .5, .5, .5])
    exact = f.exact()
    integ = vegas.Integrator(6 * [[0., 1.]], neval=200000, seed=77)
    integ(f, nitn=5)
    new = vegas.restratify(integ, f, nitn=2, ndy=5)
    assert isinstance(new, vegas.Integrator) and len(new.weight) == 6 and np.shape(new.dI) == (6, 5)
    assert 0.5 * integ.nhcube < new.nhcube <= 1.05 * integ.nhcube
    assert min(new.nstrat[:2]) > max(new.nstrat[2:]), new.nstrat
    assert abs(new.I.mean - exact) < 5 * new.I.sdev
    r = new(f, nitn=5)
    assert abs(r.mean - exact) < 5 * r.sdev and r.Q > 1e-3, (r, exact)
    # the device functor and its numpy twin on the host path give the same profile
    host = vegas.restratify(integ, vegas.lbatchintegrand(lambda x: f(x)), nitn=2, ndy=5)
    assert list(host.nstrat) == list(new.nstrat)


# ----------------------------------------------------------------------------- light geometry
@pytest.mark.parametrize('name', ['gauss4', 'genz10_pp', 'peaks20'])
def test_split_chunks_light_geometry(name, monkeypatch):
    """work items in the light geometry (512-cube chunks cut into items of 2 x 256 samples)"""
    monkeypatch.setenv('VB200_ITEM', '256')
    monkeypatch.setenv('VB200_LIGHT', '1')
    limits, f, kw = _cases()[name]
    eng = run_engine_iterations(limits, f, nitn=3, seed=99, **kw)
    assert all(r['launch']['threads'] == 256 for r in eng), eng[0]['launch']
    ora = run_oracle_iterations(limits, f, nitn=3, seed=99, engine=eng, **kw)
    compare_iterations(eng, ora, rtol=RTOL, var_rtol=1e-10)


@pytest.mark.parametrize('name', ['poly2', 'gauss4', 'ridge4_nomap', 'ridge8', 'genz10_pp', 'genz10_osc', 'genz3_corner', 'peaks20'])
def test_light_geometry_vs_oracle(name, monkeypatch):
    """the light geometry (256-thread CTAs, 512-cube chunks, grid + histogram windows in shared memory), forced on problems that would normally be too small for it, vs the oracle"""
    monkeypatch.setenv('VB200_LIGHT', '1')
    limits, f, kw = _cases()[name]
    eng = run_engine_iterations(limits, f, nitn=3, seed=4000 + len(name), **kw)
    assert all(r['launch']['threads'] == 256 and r['launch']['chunk_cubes'] == 512 for r in eng), eng[0]['launch']
    ora = run_oracle_iterations(limits, f, nitn=3, seed=4000 + len(name), engine=eng, **kw)
    compare_iterations(eng, ora, rtol=RTOL, var_rtol=1e-10)


def test_light_equals_heavy(monkeypatch):
    """both geometries consume the same Philox stream: integer results identical, sums to 1e-12"""
    limits, f, kw = _cases()['genz10_pp']
    monkeypatch.setenv('VB200_LIGHT', '0')
    a = run_engine_iterations(limits, f, nitn=2, seed=6, **kw)
    monkeypatch.setenv('VB200_LIGHT', '1')
    b = run_engine_iterations(limits, f, nitn=2, seed=6, **kw)
    assert a[0]['launch']['threads'] == 128 and b[0]['launch']['threads'] == 256
    for ra, rb in zip(a, b):
        assert np.array_equal(ra['neval_hcube'], rb['neval_hcube'])
        assert np.array_equal(ra['n_f'], rb['n_f'])
        np.testing.assert_allclose(ra['mean'], rb['mean'], rtol=1e-12)
        np.testing.assert_allclose(ra['var'], rb['var'], rtol=1e-10)
        np.testing.assert_allclose(ra['sum_f'], rb['sum_f'], rtol=1e-12, atol=1e-300)
        np.testing.assert_allclose(ra['sigf_out'], rb['sigf_out'], rtol=1e-9, atol=1e-300)


# ----------------------------------------------------------------------------- BASELINE.json full sizes
def _check_full_size(integ, f, exact, nitn_adapt, nitn, neval_lo=0.85):
    """size-independent properties at a BASELINE configuration: every sample is counted exactly once
    per axis by the training histogram; the allocation meets neval; sigf stays finite and
    non-negative with sum(sigf) == sum_sigf; the integral agrees with the exact value"""
    recs = []
    integ._trace = lambda rec: recs.append(dict(last_neval=rec['last_neval'], n_f=rec['n_f'], sum_sigf=rec['sum_sigf'],
                                                sum_f=rec['sum_f'], mean=rec['mean'], var=rec['var']))
    integ(f, nitn=nitn_adapt)
    r = integ(f, nitn=nitn)
    for rec in recs:
        assert np.array_equal(rec['n_f'].sum(axis=1), np.full(integ.dim, rec['last_neval']))
        assert neval_lo * integ.neval < rec['last_neval'] <= 1.001 * integ.neval     # int() truncation per cube (pyx:1696)
        assert np.isfinite(rec['sum_f']).all() and (rec['sum_f'] >= 0).all()
    import torch
    sg = integ._sigf_dev
    assert bool(torch.isfinite(sg).all()) and float(sg.min()) >= 0.0
    np.testing.assert_allclose(float(sg.sum()), recs[-1]['sum_sigf'], rtol=1e-10)
    assert abs(r.mean - exact) < 5 * r.sdev, (r.mean, r.sdev, exact)
    assert r.Q > 1e-4, r.Q
    return r


def test_full_size_config2_ridge():
    """BASELINE config 2: 8-D ridge (N=1000), vegas+ beta=0.75, neval=1e8, one GPU"""
    from scipy.special import erf
    vegas = _vegas()
    f = vegas.integrands.Ridge(8, N=1000)
    integ = vegas.Integrator(8 * [[0., 1.]], neval=1e8, seed=20, nitn=1)
    assert [int(v) for v in integ.nstrat] == [8, 8, 8, 8, 8, 7, 7, 7] and integ.nhcube == 11239424
    one = 0.5 * (erf(10 * (1 - f.x0)) + erf(10 * f.x0))
    r = _check_full_size(integ, f, float(np.mean(one ** 8)), 3, 3)
    assert r.sdev < 1e-4


def test_full_size_config3_genz():
    """BASELINE config 3: 10-D Genz product peak, neval=1e9 (all on one GPU here; sharded in bench/test_gpu_multi)"""
    vegas = _vegas()
    rng = np.random.default_rng(0x5eed + 3)
    f = vegas.integrands.Genz('product_peak', 2 + 3 * rng.random(10), rng.random(10))
    integ = vegas.Integrator(10 * [[0., 1.]], neval=1e9, seed=21, nitn=1, max_mem=1e10)
    assert [int(v) for v in integ.nstrat] == [7, 7, 7, 7, 6, 6, 6, 6, 6, 6] and integ.nhcube == 112021056
    r = _check_full_size(integ, f, f.exact(), 3, 2)
    assert integ._ctx.last_launch()['threads'] == 256            # cheap integrand: light geometry
    assert r.sdev < 1e-4 * abs(f.exact())


def test_large_config5_three_peaks():
    """BASELINE config 5 (20-D three-peak Gaussian, nstrat = 5 x [n] + 15 x [1]) at 1/20 of its size"""
    vegas = _vegas()
    f = vegas.integrands.GaussMix([5 * [c] + 15 * [0.45] for c in (.23, .39, .74)], 100., 356047712484621.56)
    integ = vegas.Integrator(20 * [[0., 1.]], nstrat=5 * [30] + 15 * [1], neval=5e8, seed=22, nitn=1, max_mem=1e10)
    assert integ.nhcube == 30 ** 5
    # sharp peaks: a few hypercubes hit the max_neval_hcube=50000 clamp (pyx:1704), so the total stays below neval.
    # Exact value on the unit cube (the tails of the peaks at 0.23 and 0.74 are cut, and the reference's
    # normalisation constant is 1.000295 x (100/pi)^10 / 3): 0.99914660
    from scipy.special import erf
    one = lambda c: 0.5 * (erf(10 * (1 - c)) + erf(10 * c))
    exact = np.mean([one(c) ** 5 * one(0.45) ** 15 for c in (.23, .39, .74)]) * 356047712484621.56 * 3 / (100 / np.pi) ** 10
    assert abs(exact - 0.9991466) < 1e-7
    _check_full_size(integ, f, exact, 5, 3, neval_lo=0.1)


def test_device_adapt_equals_host_adapt(monkeypatch):
    """AdaptiveMap.adapt on the device (the Integrator's fast path: no analyzer / trace hook) against the host step:
    five adapting iterations from the same seed give the same grid (log / sqrt are the device's instead of glibc's:
    1e-12 on the nodes), the same results, and integ.map.grid / inc are refreshed from the device on first access;
    a NaN leaves the map untouched"""
    vegas = _vegas()
    f = vegas.integrands.Ridge(4, N=7)

    def run(host, alpha=0.5, **kw):
        if host:
            monkeypatch.setenv('VB200_HOST_ADAPT', '1')
        else:
            monkeypatch.delenv('VB200_HOST_ADAPT', raising=False)
        integ = vegas.Integrator(4 * [[0., 1.]], neval=20000, seed=99, alpha=alpha, **kw)
        r = integ(f, nitn=5)
        return integ, r

    for alpha, kw in ((0.5, {}), (0.8, dict(maxinc_axis=50)), (1.0, {})):
        ih, rh = run(True, alpha, **kw)
        idv, rd = run(False, alpha, **kw)
        assert idv.map._device_owner is not None and ih.map._device_owner is None
        g = idv.map.grid
        assert idv.map._device_owner is None
        np.testing.assert_allclose(g, ih.map.grid, rtol=1e-12, atol=1e-14)
        np.testing.assert_allclose(idv.map.inc, ih.map.inc, rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(np.diff(g, axis=1)[:, :idv.map.ninc[0]], idv.map.inc[:, :idv.map.ninc[0]], rtol=0, atol=0)
        assert abs(rd.mean - rh.mean) < 1e-10 and abs(rd.sdev - rh.sdev) < 1e-8 * rh.sdev + 1e-14
        assert rd.sum_neval == rh.sum_neval
        # pickling and copies see the adapted grid (an unpickled integrator regrids through its defaults, as the
        # reference's does: compare the two routes with each other)
        import pickle
        i2, ih2 = pickle.loads(pickle.dumps(idv)), pickle.loads(pickle.dumps(ih))
        np.testing.assert_allclose(i2.map.grid, ih2.map.grid, rtol=1e-9, atol=1e-12)
        r2 = vegas.Integrator(idv)(f, nitn=1)
        assert abs(r2.mean - 1) < 5 * r2.sdev
    # NaN: the reference raises before adapting (pyx:2133-2134)
    monkeypatch.delenv('VB200_HOST_ADAPT', raising=False)
    integ = vegas.Integrator(4 * [[0., 1.]], neval=5000, seed=3)
    integ(f, nitn=2)
    before = integ.map.grid.copy()
    with pytest.raises(ValueError, match='nan'):
        integ(vegas.integrands.Poly(float('nan'), [1.], [1]), nitn=1)
    np.testing.assert_array_equal(integ.map.grid, before)


def test_one_call_iteration_equals_general_path(monkeypatch):
    """vb200_iteration (zero, engine, device adapt, next pre-pass, small copy back in one library call) against the
    general path of Integrator.__call__ on the same seed: identical sample counts, results and maps; also without
    adaptation (no training), with adapt_to_errors (host adapt: general path is taken by itself) and with beta = 0"""
    vegas = _vegas()
    f = vegas.integrands.GaussMix([4 * [0.5]], 100., 1013.2118364296088)
    for kw in (dict(), dict(adapt=False), dict(beta=0.), dict(adapt_to_errors=True), dict(alpha=0.)):
        out = []
        for fast in (True, False):
            if fast:
                monkeypatch.delenv('VB200_NO_FAST_ITERATION', raising=False)
            else:
                monkeypatch.setenv('VB200_NO_FAST_ITERATION', '1')
            integ = vegas.Integrator([[-1., 1.]] + 3 * [[0., 1.]], neval=10000, seed=31, **kw)
            r = integ(f, nitn=6)
            out.append((r, integ))
        (ra, ia), (rb, ib) = out
        assert ra.sum_neval == rb.sum_neval and tuple(ia.neval_hcube_range) == tuple(ib.neval_hcube_range)
        np.testing.assert_allclose([x.mean for x in ra.itn_results], [x.mean for x in rb.itn_results], rtol=1e-11)
        np.testing.assert_allclose([x.sdev for x in ra.itn_results], [x.sdev for x in rb.itn_results], rtol=1e-8)
        np.testing.assert_allclose(ia.map.grid, ib.map.grid, rtol=1e-11, atol=1e-14)
        np.testing.assert_allclose(ia.sigf, ib.sigf, rtol=1e-8)
        assert abs(ra.mean - 1) < 5 * ra.sdev


def test_deferred_bookkeeping_changes_nothing(monkeypatch):
    """iteration i booked behind the kernels of i+1 (vb200_iteration_begin/_end) against booking right away: the same
    per-iteration results; tolerances (rtol) switch the overlap off and still stop at the same iteration as the general
    path; a NaN still raises; _end without _begin is an error, not a hang"""
    vegas = _vegas()
    from vegas_b200 import _lib
    f = vegas.integrands.GaussMix([4 * [0.5]], 100., 1013.2118364296088)
    runs = {}

    def run(key, env, kw):
        for k in ('VB200_NO_DEFER', 'VB200_NO_FAST_ITERATION'):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        integ = vegas.Integrator([[-1., 1.]] + 3 * [[0., 1.]], neval=10000, seed=5)
        runs[key] = integ(f, nitn=8, **kw)

    run('defer', {}, {})
    run('now', {'VB200_NO_DEFER': '1'}, {})
    # a tolerance the weighted average reaches at the 4th iteration (or earlier), from the run just made
    first4 = runs['defer'].itn_results[:4]
    rtol = 1.2 / np.sqrt(sum(1. / x.sdev ** 2 for x in first4)) / abs(runs['defer'].mean)
    run('rtol', {}, dict(rtol=rtol))
    run('rtol_general', {'VB200_NO_FAST_ITERATION': '1'}, dict(rtol=rtol))
    a, b = runs['defer'], runs['now']
    assert len(a.itn_results) == len(b.itn_results) == 8
    # (not bit-identical: the order of the histogram's fp64 atomics differs from run to run)
    np.testing.assert_allclose([x.mean for x in a.itn_results], [x.mean for x in b.itn_results], rtol=1e-11)
    np.testing.assert_allclose([x.sdev for x in a.itn_results], [x.sdev for x in b.itn_results], rtol=1e-8)
    np.testing.assert_allclose([a.mean, a.sdev], [b.mean, b.sdev], rtol=1e-9)
    assert a.sum_neval == b.sum_neval
    c, d = runs['rtol'], runs['rtol_general']
    assert 1 <= len(c.itn_results) == len(d.itn_results) <= 4 and c.sdev < rtol * abs(c.mean)
    for k in ('VB200_NO_DEFER', 'VB200_NO_FAST_ITERATION'):
        monkeypatch.delenv(k, raising=False)
    integ = vegas.Integrator([[0., 1.]], neval=1000, seed=1)
    integ(vegas.integrands.Poly(1., [1.], [1]), nitn=3)
    with pytest.raises(ValueError, match='nan'):
        integ(vegas.integrands.Poly(float('nan'), [1.], [1]), nitn=3)
    r = integ(vegas.integrands.Poly(1., [1.], [1]), nitn=3)          # the integrator is usable afterwards
    assert abs(r.mean - 1.5) < 5 * r.sdev + 1e-9
    ctx = _lib.Context(None)
    with pytest.raises(_lib.VegasB200Error):
        ctx.iteration_end(np.empty(10))
