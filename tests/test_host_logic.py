"""CPU tests of the host-side mirror of the reference API (no GPU needed): the integer set-up of
``Integrator.set``, ``settings()``, pickling, ``AdaptiveMap`` construction / regrid (the library's
host ``vb200_map_adapt``), the result accumulators, the C-ABI export list, and the block-cyclic
sharding arithmetic.  Ported from the reference's own tests (tests/test_vegas.py, lines cited)."""
import ctypes
import os
import pickle
import re

import numpy as np
import pytest

import vegas_b200 as vegas
from vegas_b200 import _lib
from vegas_b200._gv import gv
from vegas_b200._integrator import _local_cubes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ----------------------------------------------------------------------------- C ABI
def test_library_exports_every_declared_symbol():
    """the shared library loads without a GPU and exports exactly what include/vegas_b200.h declares"""
    hdr = open(os.path.join(ROOT, 'include', 'vegas_b200.h')).read()
    declared = sorted(set(re.findall(r'\b(vb200_[a-z0-9_]+)\s*\(', hdr)))
    assert declared and set(declared) == set(_lib.SYMBOLS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert L.vb200_abi_version() == 3
    _lib.load()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_lib.VegasB200Error):
        _lib.Context()
    integ = vegas.Integrator([[0, 1]])
    with pytest.raises(_lib.VegasB200Error):
        integ(lambda x: x[0])
    m = vegas.AdaptiveMap([[0, 1]])
    with pytest.raises(_lib.VegasB200Error):
        m(np.array([[0.5]]))


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'vegas_b200')):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, fn)).read()
                assert 'import oracle' not in src and 'from oracle' not in src and 'vegas_oracle' not in src, fn


# ----------------------------------------------------------------------------- AdaptiveMap (host parts)
def test_map_init():
    """tests/test_vegas.py:39-63"""
    m = vegas.AdaptiveMap(grid=[[0, 1], [2, 4]])
    np.testing.assert_allclose(m.grid, [[0, 1], [2, 4]])
    np.testing.assert_allclose(m.inc, [[1], [2]])
    np.testing.assert_allclose(m.ninc, [1, 1])
    m = vegas.AdaptiveMap(grid=[[0, 1], [-2, 4]], ninc=2)
    np.testing.assert_allclose(m.grid, [[0, 0.5, 1.], [-2., 1., 4.]])
    np.testing.assert_allclose(m.inc, [[0.5, 0.5], [3., 3.]])
    assert m.dim == 2
    m = vegas.AdaptiveMap([[0, 0.4, 1], [-2, 0., 4]], ninc=4)
    np.testing.assert_allclose(m.grid, [[0, 0.2, 0.4, 0.7, 1.], [-2., -1., 0., 2., 4.]])
    np.testing.assert_allclose(m.inc, [[0.2, 0.2, 0.3, 0.3], [1, 1, 2, 2]])
    np.testing.assert_allclose(m.ninc, [4, 4])
    m = vegas.AdaptiveMap([[0, 1], [2, 2.5, 4]])
    np.testing.assert_allclose(m.ninc, [1, 2])
    np.testing.assert_allclose(m.grid[0, :2], [0, 1])
    np.testing.assert_allclose(m.grid[1, :3], [2, 2.5, 4])
    np.testing.assert_allclose(m.inc[0, :1], [1])
    np.testing.assert_allclose(m.inc[1, :2], [.5, 1.5])
    with pytest.raises(ValueError):
        vegas.AdaptiveMap([[0]])


def test_map_pickle_region_settings():
    """tests/test_vegas.py:65-74, 114-129"""
    m1 = vegas.AdaptiveMap(grid=[[0, 1, 3], [-2, 0, 6]])
    m2 = pickle.loads(pickle.dumps(m1))
    np.testing.assert_allclose(m2.grid, m1.grid)
    np.testing.assert_allclose(m2.inc, m1.inc)
    np.testing.assert_allclose(m1.region(0), [0, 3])
    np.testing.assert_allclose(m1.region(), [[0, 3], [-2, 6]])
    m = vegas.AdaptiveMap(grid=[[0, 1, 3], [-2, 0, 6]], ninc=4)
    out = "    grid[ 0] = [ 0.   0.5  1.   2.   3. ]\n    grid[ 1] = [-2. -1.  0.  3.  6.]\n"
    assert m.settings(5).replace(' ', '') == out.replace(' ', '')
    out = "    grid[ 0] = [ 0.5  2. ]\n    grid[ 1] = [-1.  3.]\n"
    assert m.settings(2).replace(' ', '') == out.replace(' ', '')


def test_map_adapt_host_matches_golden():
    """vb200_map_adapt (the product's host step) against the reference's adapt recorded in
    tests/golden/ref_map.npz -- training sums taken from the fixture"""
    G = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_map.npz'))
    m = vegas.AdaptiveMap([[0, 2], [-1, 1]], ninc=[50, 33])
    for i, alpha in enumerate((1.5, 0.5, -1.0)):
        m.sum_f = np.array(G['train%d_sum_f' % i])
        m.n_f = np.array(G['train%d_n_f' % i])
        m.adapt(alpha=alpha)
        g = G['train%d_grid' % i]
        for d in range(2):
            n = m.ninc[d] + 1
            np.testing.assert_allclose(m.grid[d, :n], g[d, :n], rtol=1e-14, atol=1e-16)
            np.testing.assert_array_equal(m.inc[d, :n - 1], m.grid[d, 1:n] - m.grid[d, :n - 1])
    m.adapt(ninc=[20, 7])
    for d, n in enumerate((21, 8)):
        np.testing.assert_allclose(m.grid[d, :n], G['regrid'][d, :n], rtol=1e-14, atol=1e-16)
    m.adapt(ninc=1)
    np.testing.assert_allclose(m.grid, [[0, 2], [-1, 1]])
    m.make_uniform(ninc=[4, 2])
    np.testing.assert_allclose(m.grid[0], [0, 0.5, 1, 1.5, 2])


# ----------------------------------------------------------------------------- Integrator set-up
def test_integrator_init():
    """tests/test_vegas.py:561-595"""
    I = vegas.Integrator([[0., 1.], [-1., 1.]], neval=234, nitn=123, neval_frac=0.75)
    assert I.neval == 234 and I.nitn == 123
    for k in vegas.Integrator.defaults:
        if k in ['neval', 'nitn']:
            assert getattr(I, k) != vegas.Integrator.defaults[k]
        elif k not in ['map', 'xparam']:
            assert getattr(I, k) == vegas.Integrator.defaults[k]
    np.testing.assert_allclose([I.map.grid[0, 0], I.map.grid[0, I.map.ninc[0]]], [0., 1.])
    np.testing.assert_allclose([I.map.grid[1, 0], I.map.grid[1, I.map.ninc[1]]], [-1., 1.])
    assert list(I.map.ninc) == [20, 20] and list(I.nstrat) == [5, 5]
    I = vegas.Integrator([[0., 1.], [-1., 1.]], nstrat=[1, 1], neval=1000)
    assert list(I.map.ninc) == [100, 100] and list(I.nstrat) == [1, 1]
    assert I.neval == 1000 and I.min_neval_hcube == 1000
    I = vegas.Integrator([[0., 1.], [-1., 1.]], nstrat=[10, 11], neval_frac=0.75)
    assert list(I.map.ninc) == [80, 88] and list(I.nstrat) == [10, 11]
    assert I.neval == 880 and I.min_neval_hcube == 2


def test_integrator_set():
    """tests/test_vegas.py:703-772"""
    new_defaults = dict(
        map=vegas.AdaptiveMap([[1, 2], [0, 1]]), neval=100, maxinc_axis=100, min_neval_batch=10,
        max_neval_hcube=1e1, max_mem=229, nitn=100, alpha=0.35, beta=0.25, adapt_to_errors=True,
        rtol=0.1, atol=0.2, analyzer=vegas.reporter(5))
    I = vegas.Integrator([[1, 2]])
    old = I.set(**new_defaults)
    for k in new_defaults:
        if k == 'map':
            np.testing.assert_allclose([[I.map.grid[0, 0], I.map.grid[0, I.map.ninc[0]]],
                                        [I.map.grid[1, 0], I.map.grid[1, I.map.ninc[1]]]], new_defaults['map'].grid)
        else:
            assert getattr(I, k) == new_defaults[k]
    assert old['neval'] == 1000 and old['alpha'] == 0.5
    I = vegas.Integrator([[1, 1.3, 2], [0, 1]], maxinc_axis=1000, neval_frac=0.75)
    I.set(nstrat=[22, 13])
    assert I.neval == 22 * 13 * 2 / 0.25 and I.min_neval_hcube == 2
    I.set(nstrat=[7, 9], neval=2000)
    assert I.neval == 2000 and list(I.nstrat) == [7, 9] and list(I.map.ninc) == [196, 198]
    assert I.min_neval_hcube == 7
    I.set(nstrat=[7, 9])
    assert I.neval == 7 * 9 * 2 / 0.25 and I.min_neval_hcube == 2
    with pytest.raises(ValueError):
        I.set(nstrat=[7, 9], neval=20)
    I.set(neval=3100)
    assert list(I.nstrat) == [20, 19] and I.min_neval_hcube == 2
    with pytest.raises(ValueError):
        I.set(nstrat=[2, 3, 5])
    I.set(neval=3500)
    assert list(I.nstrat) == [21, 20]
    with pytest.raises(ValueError):
        I.set(nstrat=[10, 12], neval=120 * 4)
    I.set(nstrat=[10, 12], neval=120 * 8)
    assert I.neval == 120 * 8 and I.min_neval_hcube == 2
    I.set(neval=3.5e8)
    nstrat, ninc = np.array(I.nstrat), np.array(I.map.ninc)
    assert np.all(np.round(nstrat / ninc) * ninc == nstrat)
    I.set(neval=300)
    nstrat, ninc = np.array(I.nstrat), np.array(I.map.ninc)
    assert np.all(np.round(ninc / nstrat) * nstrat == ninc)
    I.set(sigf=[1.])
    assert len(I.sigf) != 1
    with pytest.raises(AttributeError):
        I.set(no_such_parameter=1)
    I.set(nhcube_batch=10)          # legacy key: ignored


def test_strata_match_golden_and_survey_sizes():
    """integer outputs of set() for the BASELINE configs and the golden cases"""
    from tests.golden.cases import CASES
    for name, spec in CASES.items():
        G = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_%s.npz' % name))
        I = vegas.Integrator(spec['limits'], **spec['kw'])
        assert list(I.nstrat) == list(G['nstrat']) and list(I.map.ninc) == list(G['ninc']), name
        assert I.nhcube == int(G['nhcube']) and I.min_neval_hcube == int(G['min_neval_hcube']), name
    I = vegas.Integrator(4 * [[0, 1]], neval=1e4)
    assert list(I.nstrat) == [6, 6, 6, 5] and list(I.map.ninc) == [996, 996, 996, 1000] and I.nhcube == 1080
    I = vegas.Integrator(8 * [[0, 1]], neval=1e8)
    assert list(I.nstrat) == [8, 8, 8, 8, 8, 7, 7, 7] and I.nhcube == 11239424 and I.min_neval_hcube == 2
    I = vegas.Integrator(10 * [[0, 1]], neval=1e9)
    assert list(I.nstrat) == [7, 7, 7, 7, 6, 6, 6, 6, 6, 6] and I.nhcube == 112021056
    with pytest.raises(MemoryError):
        vegas.Integrator(20 * [[0, 1]], neval=1e10)
    I = vegas.Integrator(20 * [[0, 1]], neval=1e10, nstrat=5 * [60] + 15 * [1], max_mem=2e10)
    assert I.nhcube == 777600000 and I.min_neval_hcube == 3


def test_settings_strings():
    """tests/test_vegas.py:597-673"""
    head = [
        "Integrator Settings:",
        "    {neval} (approx) integrand evaluations in each of 123 iterations",
        "    number of: strata/axis = [{nstrat0} {nstrat1}]",
        "               increments/axis = [{ninc0} {ninc1}]",
        "               h-cubes = {nhcube}  processors = 1",
        "               evaluations/batch >= {min_neval_batch:.2g}",
        "               {min_neval_hcube} <= evaluations/h-cube <= {max_neval_hcube:.2g}",
        "    minimize_mem = False  adapt_to_errors = False  adapt = True",
        "    accuracy: relative = 0  absolute = 0",
        "    damping: alpha = {alpha}  beta= {beta}",
        ""]
    tails = [
        ([[0., 1.], [-1., 1.]], ["    axis    integration limits", "    --------------------------",
                                 "       0            (0.0, 1.0)", "       1           (-1.0, 1.0)\n"]),
        (dict(x=[[0., 1.]], y=[-1., 1.]), ["    key/index    axis    integration limits",
                                            "    ---------------------------------------",
                                            "          x 0       0            (0.0, 1.0)",
                                            "            y       1           (-1.0, 1.0)\n"]),
        ([[[0., 1.]], [[-1., 1.]]], ["    key/index    axis    integration limits",
                                     "    ---------------------------------------",
                                     "          0,0       0            (0.0, 1.0)",
                                     "          1,0       1           (-1.0, 1.0)\n"]),
    ]
    for limits, tail in tails:
        I = vegas.Integrator(limits, neval=254, nitn=123, neval_frac=0.75)
        out = '\n'.join(head + tail).format(
            neval=I.neval, nstrat0=I.nstrat[0], nstrat1=I.nstrat[1], ninc0=I.map.ninc[0], ninc1=I.map.ninc[1],
            nhcube=I.nhcube, min_neval_hcube=I.min_neval_hcube, alpha=I.alpha, beta=I.beta,
            min_neval_batch=I.min_neval_batch, max_neval_hcube=float(I.max_neval_hcube))
        assert out == I.settings()


def test_integrator_pickle():
    """tests/test_vegas.py:675-701"""
    I1 = vegas.Integrator([[0., 1.], [-1., 1.]], neval=234)
    I2 = pickle.loads(pickle.dumps(I1))
    assert isinstance(I2, vegas.Integrator)
    for k in vegas.Integrator.defaults:
        if k == 'map':
            np.testing.assert_allclose(I1.map.ninc, I2.map.ninc)
            for d in range(I1.dim):
                n = I1.map.ninc[d] + 1
                np.testing.assert_allclose(I1.map.grid[d, :n] + 1e-8, I2.map.grid[d, :n] + 1e-8, rtol=0.01)
        elif k != 'ran_array_generator':
            assert getattr(I1, k) == getattr(I2, k)
    assert list(I2.nstrat) == list(I1.nstrat) and len(I2.sigf) == len(I1.sigf)
    I3 = vegas.Integrator(I1, alpha=0.1)
    assert I3.alpha == 0.1 and list(I3.nstrat) == list(I1.nstrat) and I3.neval == I1.neval


# ----------------------------------------------------------------------------- results
def test_ravg_known_answers():
    """tests/test_vegas.py:216-236"""
    a = vegas.RAvg()
    a.add(gv.gvar(1, 1))
    a.add(gv.gvar(2, 2))
    a.add(gv.gvar(3, 3))
    np.testing.assert_allclose(a.mean, 1.346938775510204)
    np.testing.assert_allclose(a.sdev, 0.8571428571428571)
    assert a.dof == 2
    np.testing.assert_allclose(a.chi2, 0.5306122448979592)
    np.testing.assert_allclose(a.Q, 0.7669711269557102)
    assert str(a) == '1.35(86)'
    s = ["itn   integral        wgt average     chi2/dof        Q",
         "-------------------------------------------------------",
         "  1   1.0(1.0)        1.0(1.0)            0.00     1.00",
         "  2   2.0(2.0)        1.20(89)            0.20     0.65",
         "  3   3.0(3.0)        1.35(86)            0.27     0.77", ""]
    assert a.summary() == '\n'.join(s)
    b = pickle.loads(pickle.dumps(a))
    assert str(b) == '1.35(86)' and b.dof == 2
    a.extend(b)
    assert a.nitn == 6
    u = vegas.RAvg(weighted=False)
    for m, s in ((1, 1), (2, 2), (3, 3)):
        u.add(gv.gvar(m, s))
    np.testing.assert_allclose(u.mean, 2.0)
    np.testing.assert_allclose(u.sdev, (14. / 9.) ** 0.5)


def test_ravgarray_ravgdict_known_answers():
    """tests/test_vegas.py:312-361"""
    a = vegas.RAvgArray((1, 2))
    a.add([[gv.gvar(1, 1), gv.gvar(10, 10)]])
    a.add([[gv.gvar(2, 2), gv.gvar(20, 20)]])
    a.add([[gv.gvar(3, 3), gv.gvar(30, 30)]])
    assert a.shape == (1, 2)
    np.testing.assert_allclose(a[0, 0].mean, 1.346938775510204)
    np.testing.assert_allclose(a[0, 0].sdev, 0.8571428571428571)
    assert a.dof == 4
    np.testing.assert_allclose(a.chi2, 2 * 0.5306122448979592)
    np.testing.assert_allclose(a.Q, 0.900374555485)
    assert str(a[0, 0]) == '1.35(86)' and str(a[0, 1]) == '13.5(8.6)'
    s = ["itn   integral        wgt average     chi2/dof        Q",
         "-------------------------------------------------------",
         "  1   1.0(1.0)        1.0(1.0)            0.00     1.00",
         "  2   2.0(2.0)        1.20(89)            0.20     0.82",
         "  3   3.0(3.0)        1.35(86)            0.27     0.90", ""]
    assert a.summary() == '\n'.join(s)
    d = vegas.RAvgDict(dict(s=1.0, a=[[2.0, 3.0]]))
    d.add(dict(s=gv.gvar(1, 1), a=[[gv.gvar(1, 1), gv.gvar(10, 10)]]))
    d.add(dict(s=gv.gvar(2, 2), a=[[gv.gvar(2, 2), gv.gvar(20, 20)]]))
    d.add(dict(s=gv.gvar(3, 3), a=[[gv.gvar(3, 3), gv.gvar(30, 30)]]))
    assert d['a'].shape == (1, 2)
    np.testing.assert_allclose(d['a'][0, 0].mean, 1.346938775510204)
    assert str(d['a'][0, 1]) == '13.5(8.6)' and str(d['s']) == '1.35(86)'
    assert d.dof == 6
    np.testing.assert_allclose(d.chi2, 3 * 0.5306122448979592)
    np.testing.assert_allclose(d.Q, 0.953162484587)


# ----------------------------------------------------------------------------- integrand adapters
def test_integrand_adapter_shapes():
    """tests/test_vegas.py:840-969 (subset): every integrand kind is normalised to eval(x[n,D]) -> f[n,size]"""
    m = vegas.AdaptiveMap([[0, 1], [0, 2]])
    xs = np.array([0.3, 0.7])
    x = np.array([[0.1, 0.2], [0.3, 0.4], [0.5, 0.6]])

    def scalar(x):
        return x[0] + x[1]

    def scalar_arr(x):
        return [x[0], x[1], x[0] * x[1]]

    def scalar_dict(x):
        return dict(a=x[0], b=[x[1], x[0] * x[1]])

    @vegas.lbatchintegrand
    def lb(x):
        return x[:, 0] + x[:, 1]

    @vegas.rbatchintegrand
    def rb(x):
        return dict(a=x[0], b=[x[1], x[0] * x[1]])

    class C(vegas.LBatchIntegrand):
        def __call__(self, x):
            return np.stack([x[:, 0], x[:, 1]], axis=1).reshape(-1, 1, 2)

    v = vegas.VegasIntegrand(scalar, m, False, xs, False)
    assert v.shape == () and v.size == 1
    np.testing.assert_allclose(v.eval(x)[:, 0], x.sum(axis=1))
    v = vegas.VegasIntegrand(scalar_arr, m, False, xs, False)
    assert v.shape == (3,) and v.eval(x).shape == (3, 3)
    v = vegas.VegasIntegrand(scalar_dict, m, False, xs, False)
    assert v.shape is None and v.size == 3
    np.testing.assert_allclose(v.eval(x), np.stack([x[:, 0], x[:, 1], x[:, 0] * x[:, 1]], axis=1))
    v = vegas.VegasIntegrand(lb, m, False, xs, False)
    assert v.shape == () and v.eval(x).shape == (3, 1)
    v = vegas.VegasIntegrand(rb, m, False, xs, False)
    assert v.shape is None and v.size == 3
    np.testing.assert_allclose(v.eval(x), np.stack([x[:, 0], x[:, 1], x[:, 0] * x[:, 1]], axis=1))
    r = v.format_result(np.array([1., 2., 3.]), np.diag([1., 4., 9.]))
    assert str(r['a']) == '1.0(1.0)' and r['b'].shape == (2,)
    v = vegas.VegasIntegrand(C(), m, False, xs, False)
    assert v.shape == (1, 2) and v.eval(x).shape == (3, 2)
    with pytest.raises(ValueError):
        vegas.VegasIntegrand(C, m, False, xs, False)
    # dictionary-valued arguments
    I = vegas.Integrator(dict(u=[0., 1.], w=[[0., 1.], [0., 2.]]))

    def fd(xd):
        return xd['u'] * xd['w'][0] + xd['w'][1]
    v = I._make_std_integrand(fd)
    np.testing.assert_allclose(v.eval(np.array([[0.5, 0.2, 1.0]]))[0, 0], 0.5 * 0.2 + 1.0)


def test_device_integrand_twins_are_consistent():
    F = vegas.integrands
    rng = np.random.default_rng(1)
    x = rng.random((50, 8))
    np.testing.assert_allclose(F.Ridge(8, N=7)(x), F.Ridge(8, N=7, shifted=True)(x))
    p = F.PathIntegral(T=4., ndT=8, x0list=np.linspace(0, 2, 6))
    out = p(x - 0.5)
    assert list(out) == ['exp(-E0*T)', 'exp(-E0*T) * psi(x0)**2'] and out['exp(-E0*T) * psi(x0)**2'].shape == (50, 6)
    # Genz closed forms against brute-force quadrature in 2-D
    t = (np.arange(400) + 0.5) / 400
    X = np.stack(np.meshgrid(t, t, indexing='ij'), axis=-1).reshape(-1, 2)
    for kind in F.Genz.KINDS:
        g = F.Genz(kind, [1.3, 2.1], [0.4, 0.7])
        np.testing.assert_allclose(g(X).mean(), g.exact(), rtol=3e-3, err_msg=kind)


# ----------------------------------------------------------------------------- sharding arithmetic
@pytest.mark.parametrize('nh,slab,world', [(1080, 256, 2), (11239424, 16384, 8), (777, 256, 4), (256, 256, 3), (5, 256, 2)])
def test_block_cyclic_partition_covers_every_cube_once(nh, slab, world):
    seen = np.zeros(nh, np.int32)
    for r in range(world):
        idx = _local_cubes(nh, slab, r, world)
        seen[idx] += 1
        # matches the library's count
        L = _lib.load()
        out = ctypes.c_int64()
        h = ctypes.c_void_p()
    assert np.all(seen == 1)


def test_new_stratification_matches_reference_golden():
    """restratify's strata assignment (src/vegas/__init__.py:1349-1365) on the weights the unmodified
    reference computed (tests/golden/ref_restratify.npz): identical integer nstrat"""
    from vegas_b200._restratify import new_stratification
    from tests.golden.cases import RESTRATIFY
    G = np.load(os.path.join(ROOT, 'tests', 'golden', 'ref_restratify.npz'))
    for name, spec in RESTRATIFY.items():
        new = new_stratification(G[name + '_old_nstrat'], G[name + '_weight'], **spec['opt'])
        assert list(new) == list(G[name + '_new_nstrat']), (name, new)
    # axes rounded down to one stratum hand their share to the others: the product stays close
    new = new_stratification([4, 4, 4, 4], [1.0, 1e-6, 1e-6, 0.5])
    assert list(new[1:3]) == [1, 1] and 0.7 * 256 < np.prod(new) <= 256


def test_ravg_function():
    """vegas.ravg (src/vegas/__init__.py:1220-1311): running averages rebuilt from iteration results"""
    import vegas_b200 as vegas
    from vegas_b200._gv import gv
    rs = [gv.gvar(1.0, 0.1), gv.gvar(1.2, 0.2), gv.gvar(0.9, 0.1)]
    w, u = vegas.ravg(rs), vegas.ravg(rs, weighted=False)
    wts = np.array([100., 25., 100.])
    assert abs(w.mean - (wts * [1.0, 1.2, 0.9]).sum() / wts.sum()) < 1e-12 and abs(w.sdev - wts.sum() ** -0.5) < 1e-12
    assert abs(u.mean - np.mean([1.0, 1.2, 0.9])) < 1e-12 and w.nitn == u.nitn == 3
    assert abs(vegas.ravg(w, weighted=False).mean - u.mean) < 1e-12          # from a result object: its itn_results
    a = vegas.ravg([gv.gvar([1.0, 2.0], [0.1, 0.2]), gv.gvar([1.1, 2.1], [0.1, 0.2])])
    assert a.shape == (2,) and abs(a[0].mean - 1.05) < 1e-12
    d = vegas.ravg([dict(a=gv.gvar(1, 0.1)), dict(a=gv.gvar(1.1, 0.1))])
    assert abs(d['a'].mean - 1.05) < 1e-12
    with pytest.raises(ValueError):
        vegas.ravg([])


def test_ravg_and_ravgarray_weighted_unweighted():
    """reference tests:238-312: error of weighted / unweighted running averages of scalars and of
    correlated arrays, degrees of freedom, Q"""
    import vegas_b200 as vegas
    from vegas_b200._gv import gv
    rng = np.random.default_rng(123)
    N = 30
    mean = rng.uniform(-10., 10.)
    ravg = vegas.RAvg()
    for i in range(N):
        ravg.add(gv.gvar(rng.normal(mean, 1.), 1.))
        ravg.add(gv.gvar(rng.normal(mean, 0.1), 0.1))
    np.testing.assert_allclose(ravg.sdev, 1 / (N * (1. / 1. + 1. / 0.01)) ** 0.5)
    assert abs(ravg.mean - mean) < 5 * ravg.sdev and ravg.Q > 1e-3 and ravg.dof == 2 * N - 1
    ravg = vegas.RAvg(weighted=False)
    for i in range(N):
        ravg.add(gv.gvar(rng.normal(mean, 0.1), 0.1))
    np.testing.assert_allclose(ravg.sdev, 0.1 / N ** 0.5)
    assert abs(ravg.mean - mean) < 5 * ravg.sdev and ravg.Q > 1e-3 and ravg.dof == N - 1
    # arrays with correlations
    mean = rng.uniform(-10., 10., (2,))
    cov = np.array([[1., 0.5], [0.5, 2.]])
    ravg = vegas.RAvgArray((1, 2))
    for i in range(N):
        ravg.add([gv.gvar(rng.multivariate_normal(mean, cov), cov)])
        ravg.add([gv.gvar(rng.multivariate_normal(mean, cov / 10.), cov / 10.)])
    np.testing.assert_allclose(gv.evalcov(ravg.flat), cov / (10. + 1.) / N, rtol=1e-10)
    for i in range(2):
        assert abs(mean[i] - ravg[0, i].mean) < 5 * ravg[0, i].sdev
    assert ravg.dof == 4 * N - 2 and ravg.Q > 1e-3
    ravg = vegas.RAvgArray((1, 2), weighted=False)
    for i in range(N):
        ravg.add([gv.gvar(rng.multivariate_normal(mean, cov / 10.), cov / 10.)])
    np.testing.assert_allclose(gv.evalcov(ravg.flat), cov / 10. / N, rtol=1e-10)
    assert ravg.dof == 2 * N - 2 and ravg.Q > 1e-3


def test_max_mem():
    """reference tests:1401-1408: max_mem is enforced at construction and when neval is raised in the call"""
    import vegas_b200 as vegas
    with pytest.raises(MemoryError):
        vegas.Integrator(3 * [(0, 1)], max_mem=10)
    I = vegas.Integrator(3 * [(0, 1)], max_mem=12012, minimize_mem=True)
    with pytest.raises(MemoryError):
        I(lambda x: np.prod(x), neval=1e4)


def test_rescaling():
    """reference tests:535-552: weighted averages of components that differ by 50 orders of magnitude
    (rescaling by the last result keeps the covariance matrix invertible)"""
    import vegas_b200 as vegas
    from vegas_b200._gv import gv
    rng = np.random.default_rng(3)
    x = gv.gvar(1, 0.001)
    a = vegas.RAvgArray((2,))
    for i in range(3):
        xx = x - x.mean + rng.normal(1, 0.001)
        a.add([xx, 1e50 * xx])
    assert str(a[0] * 1e50) == str(a[1])
    assert str(gv.evalcorr(a).flat[:]) == str(np.ones(4, float))
    d = vegas.RAvgDict(dict(a=1., b=2.))
    for i in range(3):
        xx = x - x.mean + rng.normal(1, 0.001)
        d.add(dict(a=xx, b=1e50 * xx))
    assert str(d['a'] * 1e50) == str(d['b'])
    assert str(gv.evalcorr(d.buf).flat[:]) == str(np.ones(4, float))


def test_dimension_limit_is_a_clear_error():
    """the kernels carry per-axis parameters in fixed-size blocks (VB_MAXD = 32); the reference has no limit, so
    the difference must surface as a plain message at construction, not as a launch failure"""
    import pytest
    with pytest.raises(ValueError, match='up to 32 dimensions'):
        vegas.Integrator(33 * [[0., 1.]])
    assert vegas.Integrator(32 * [[0., 1.]]).dim == 32


def test_wide_reduce_plan_covers_every_accumulator_once():
    """integrands with more components than the reduce kernel's instantiations are reduced in column subsets
    (Integrator._wide_passes): every mean, every (lower-triangle or diagonal) covariance entry and sum_sigf is
    taken from exactly one pass; a pass has at most 8 columns; the first pass starts with component 0 (it is
    the one that trains the map and updates sigf)"""
    import torch
    for correlate in (True, False):
        for nf in (9, 13, 30):
            integ = vegas.Integrator(2 * [[0., 1.]], correlate_integrals=correlate)
            passes = integ._wide_passes(nf, 'cpu', torch)
            nv = nf * (nf + 1) // 2
            seen = np.zeros(nf + nv + 1, int)
            for k, (cols, m, src, dst) in enumerate(passes):
                cols = cols.tolist()
                assert m == len(cols) <= 8 and cols == sorted(cols)
                assert len(src) == len(dst) and int(src.max()) < m + m * (m + 1) // 2 + 1
                np.add.at(seen, dst.numpy(), 1)
                if k == 0:
                    assert cols[0] == 0 and int(dst[-1]) == nf + nv
            want = np.ones(nf + nv + 1, int)
            if not correlate:
                want[nf:nf + nv] = 0
                want[[nf + s * (s + 1) // 2 + s for s in range(nf)]] = 1
            assert np.array_equal(seen, want), (correlate, nf)


def test_save_only_on_the_writer_rank(tmp_path):
    """with the hypercube range sharded over ranks every rank holds the same results; only rank 0 writes the
    pickle (pyx:2923-2941), the others still serialise (pickling the integrator gathers sigf: a collective)"""
    import pickle
    from vegas_b200._results import VegasResult

    class Std(object):
        shape = ()

        def format_result(self, mean, var):
            from vegas_b200._gv import gv
            return gv.gvar(mean[0], var[0, 0] ** 0.5)

    res = VegasResult(Std(), weighted=True)
    res.update(np.array([1.0]), np.array([[0.01]]), 100)
    calls = []

    class Integ(object):
        def __reduce__(self):
            calls.append(1)
            return (dict, ())

    for writer in (False, True):
        res.is_writer = writer
        p1, p2 = tmp_path / ('r%d.pkl' % writer), tmp_path / ('a%d.pkl' % writer)
        res.save(str(p1))
        res.saveall(Integ(), str(p2))
        assert p1.exists() == writer and p2.exists() == writer
    assert len(calls) == 2                       # the integrator was pickled on the non-writer too
    r, i = pickle.load(open(str(tmp_path / 'a1.pkl'), 'rb'))
    assert abs(r.mean - 1.0) < 1e-12 and i == {}


def test_replacing_sigf_voids_a_waiting_prepass():
    """the allocation pre-pass launched behind the last iteration of a call waits for the next call
    (Integrator._plan_ahead); every replacement of the device copy of sigf -- set(sigf=...), a new stratification, a
    new context -- must void it (the caching allocator may hand the new tensor the old address, so the address in
    _plan_key proves nothing), while reading it back does not"""
    integ = vegas.Integrator(3 * [[0., 1.]], neval=1000)
    integ._plan_ahead = ('key', None)
    assert integ._sigf_dev is None and integ._plan_ahead == ('key', None)
    integ.set(sigf=np.ones(integ.nhcube))
    assert integ._plan_ahead is None
    integ._plan_ahead = ('key', None)
    integ.set(nstrat=[2, 2, 2])
    assert integ._plan_ahead is None
    integ._plan_ahead = ('key', None)
    integ._sigf_dev_stale()                      # nothing on the device: nothing replaced
    assert integ._plan_ahead == ('key', None)
    assert '_sigf_dev_t' in integ.__dict__ and '_sigf_dev' not in integ.__dict__


def test_integrand_adapters_every_argument_and_value_form():
    """VegasIntegrand.eval over the whole matrix the reference supports (pyx:2959-3383): x presented as flat rows,
    as an index array of xsample's shape or as a dictionary; one point at a time, lbatch, rbatch; values that are
    numbers, arrays or dictionaries; with and without jac.  Expected rows are computed directly from the flat
    x[n, D].  (jac reaches a one-point function by keyword for flat x, as second argument for a dictionary, and
    not at all for an index array: pyx:3200-3213.)"""
    import itertools
    from vegas_b200._integrand import VegasIntegrand
    D, n = 6, 5
    xsamples = dict(flat=np.linspace(0.1, 0.6, D), grid=np.arange(6.).reshape(2, 3) / 10.,
                    dict=gv.BufferDict([('a', 0.5), ('b', np.array([0.1, 0.2, 0.3])), ('c', np.array([[0.7, 0.8]]))]))

    def flatten(xa, kind):       # whatever the user function was handed -> c[D] (one point) or c[D, n]
        parts = [np.asarray(xa[k]) for k in xa] if hasattr(xa, 'keys') else [np.asarray(xa)]
        if kind == 'scalar':
            return np.concatenate([p.reshape(-1) for p in parts])
        if kind == 'lbatch':
            return np.concatenate([p.reshape(p.shape[0], -1) for p in parts], axis=1).T
        return np.concatenate([p.reshape(-1, p.shape[-1]) for p in parts], axis=0)

    def user(kind, out, with_jac):
        def body(x, jac=None):
            c = flatten(x, kind)
            if jac is not None:
                c = c * (1 + flatten(jac, kind))
            s, v = c.sum(axis=0), [c[0] * c[1], np.sin(c[2]), c[3] ** 2]
            m = [[c[i] * c[j] for j in (2, 3, 4)] for i in (0, 1)]
            if kind == 'lbatch':       # batch index first
                v, m = np.stack(v, axis=-1), np.moveaxis(np.array(m), -1, 0)
            else:
                v, m = np.array(v), np.array(m)
            return dict(number=s, vector=v, matrix=m, dictionary=dict(s=s, v=v, m=m))[out]
        f = (lambda x, jac=None: body(x, jac)) if with_jac else (lambda x: body(x))
        return dict(scalar=lambda g: g, lbatch=vegas.lbatchintegrand, rbatch=vegas.rbatchintegrand)[kind](f)

    rng = np.random.default_rng(2)
    for (xname, xs), kind, out, with_jac in itertools.product(xsamples.items(), ('scalar', 'lbatch', 'rbatch'),
                                                              ('number', 'vector', 'matrix', 'dictionary'), (False, True)):
        std = VegasIntegrand(user(kind, out, with_jac), None, with_jac, xs, False)
        x, jac = rng.uniform(size=(n, D)), (rng.uniform(size=(n, D)) if with_jac else None)
        c = x if (jac is None or (kind == 'scalar' and xname == 'grid')) else x * (1 + jac)
        s, v = c.sum(axis=1)[:, None], np.stack([c[:, 0] * c[:, 1], np.sin(c[:, 2]), c[:, 3] ** 2], axis=1)
        m = np.einsum('ni,nj->nij', c[:, :2], c[:, 2:5]).reshape(n, 6)
        want = dict(number=s, vector=v, matrix=m, dictionary=np.concatenate([s, v, m], axis=1))[out]
        label = (xname, kind, out, with_jac)
        assert std.size == want.shape[1] and std.fcntype == kind, label
        assert std.shape == dict(number=(), vector=(3,), matrix=(2, 3), dictionary=None)[out], label
        got = std.eval(x, jac=jac)
        assert got.shape == want.shape, label
        np.testing.assert_allclose(got, want, rtol=1e-15, atol=0, err_msg=str(label))
        np.testing.assert_array_equal(std.training(x, jac), got[:, 0])
        # the value's own structure back: means, GVars from a covariance matrix / from variances, evalx
        mean, cov = got.mean(axis=0), np.cov(got.T).reshape(std.size, std.size) + 1e-3 * np.eye(std.size)
        for formatted, sd in ((std.format_result(mean), None), (std.format_result(mean, cov), np.sqrt(np.diag(cov))),
                              (std.format_result(mean, np.diag(cov).copy()), np.sqrt(np.diag(cov)))):
            if out == 'dictionary':
                assert list(formatted.keys()) == ['s', 'v', 'm'] and np.shape(formatted['m']) == (2, 3)
                flat = formatted.buf
            else:
                assert np.shape(formatted) == std.shape
                flat = np.reshape(formatted, -1)
            np.testing.assert_allclose(gv.mean(flat) if sd is not None else np.asarray(flat, float), mean, rtol=1e-15)
            if sd is not None:
                np.testing.assert_allclose(gv.sdev(flat), sd, rtol=1e-12)
        ex = std.format_evalx(got)
        if out == 'dictionary':
            assert np.shape(ex['v']) == (n, 3) and np.shape(ex['s']) == (n,)
        else:
            assert ex.shape == (n,) + std.shape
        one = std(x[0], jac=None if jac is None else jac[:1])
        np.testing.assert_allclose(one.buf if out == 'dictionary' else np.reshape(one, -1), got[0], rtol=1e-15)
    # a standard-form integrand pickles whenever the user function does (PDFIntegrator pickles its pdf; saveall)
    std = VegasIntegrand(vegas.lbatchintegrand(_picklable_lbatch), None, False, xsamples['dict'], False)
    again = pickle.loads(pickle.dumps(std))
    x = rng.uniform(size=(n, D))
    np.testing.assert_array_equal(again.eval(x), std.eval(x))
    assert again.shape == std.shape == (2,) and again.size == 2


def _picklable_lbatch(p):
    return np.stack([p['a'] * p['b'][:, 0], p['c'][:, 0, 1]], axis=1)


def test_settings_strings_match_the_reference_fixture():
    """Integrator.settings() / AdaptiveMap.settings() against the strings recorded from the unmodified reference
    (tests/golden/ref_settings.json, made by make_golden_settings.py): one and two column tables (more than 20
    axes; an odd split), dictionary and index-array xsample, no adaptation, adapt_to_errors, beta=0, explicit
    nstrat, limits that need exponents, grid nodes"""
    import json
    from tests.golden.cases import SETTINGS, settings_limits
    ref = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'ref_settings.json')))
    assert sorted(ref) == sorted(SETTINGS)
    for name, spec in SETTINGS.items():
        integ = vegas.Integrator(settings_limits(spec['limits']), **spec['kw'])
        assert integ.settings(ngrid=spec.get('ngrid', 0)) == ref[name], name


class _ScriptedContext(object):
    """stands in for _lib.Context in the host loop of Integrator.__call__: records the calls and delivers scripted
    iteration heads ([mean, var, sum_sigf | NaN flag, 6 pre-pass words]); no device, no library"""
    device = 'cpu'

    def __init__(self, heads):
        self.heads, self.log, self.integrand_serial, self.in_flight = list(heads), [], 0, None

    def set_seed(self, seed):
        pass

    def set_map(self, grid, ninc):
        self.log.append('set_map')

    def set_strata(self, nstrat, slab, rank=0, world=1):
        self.log.append('set_strata')
        return int(np.prod(nstrat))

    def set_integrand(self, fid, params, keep=None):
        _lib.note_integrand(self, fid, params)
        return 1

    def plan(self, sigf, neval_sigf, min_nh, max_nh, uniform, neval_hcube=None):
        self.log.append('plan')
        return 1000, 2, 9, 1

    def plan_commit(self, neval_sigf, stats6):
        self.log.append(('plan_commit', [int(v) for v in stats6]))
        return 1000, 2, 9, 1

    def iteration_begin(self, itn, beta, flags, sigf, buf, nacc, nh, hstride, nf64, nwords, alpha_adapt, plan):
        assert self.in_flight is None
        self.in_flight = self.heads.pop(0)
        self.log.append(('begin', alpha_adapt > 0, plan is not None))

    def iteration_end(self, head):
        mean, var, sum_sigf, nan = self.in_flight
        self.in_flight = None
        head[:3] = mean, var, sum_sigf
        head[3:].view(np.int64)[:] = [nan, 11, 12, 13, 14, 15, 16]
        self.log.append('end')

    def get_map(self, shape):
        return np.tile(np.linspace(0., 1., shape[1]), (shape[0], 1))

    def launch_count(self):
        return 0


def _scripted_integrator(heads, **kw):
    integ = vegas.Integrator(2 * [[0., 1.]], neval=1000, seed=1, **kw)
    integ._ctx, integ._ctx_map_version, integ._ctx_strata = _ScriptedContext(heads), None, None
    return integ, integ._ctx


def test_host_loop_books_iteration_i_behind_the_launch_of_i_plus_1(monkeypatch):
    """the host side of the one-call iteration path (no GPU: a scripted context): results are booked in order, each
    behind the next launch, the last one after the loop; the pre-pass words of iteration i are committed before
    launch i + 1 and -- left by a call's last iteration -- at the start of the next call; with a tolerance every
    iteration is booked at once and the loop stops where the running average meets it; a NaN raises before anything
    of that iteration is booked and resets sigf; the context is brought up to date only once per call"""
    for k in ('VB200_NO_DEFER', 'VB200_NO_FAST_ITERATION', 'VB200_HOST_ADAPT', 'VB200_NO_PLAN_AHEAD'):
        monkeypatch.delenv(k, raising=False)
    f = vegas.integrands.Poly(1., [1.], [1])
    heads = [(1.5 + 0.01 * i, 1e-4, 400., 0) for i in range(4)]
    integ, ctx = _scripted_integrator(heads)
    order = []
    real_update = vegas._results.VegasResult.update
    monkeypatch.setattr(vegas._results.VegasResult, 'update',
                        lambda self, mean, var, neval: (order.append((float(mean[0]), len(ctx.log))), real_update(self, mean, var, neval))[1])
    r = integ(f, nitn=4)
    assert [x.mean for x in r.itn_results] == [1.5, 1.51, 1.52, 1.53] and r.sum_neval == 4000
    ends = [i for i, e in enumerate(ctx.log) if e == 'end']
    begins = [i for i, e in enumerate(ctx.log) if isinstance(e, tuple) and e[0] == 'begin']
    # iteration i is booked after launch i + 1 went out and before its wait returned; the last one after the loop
    assert [n for _, n in order] == [begins[1] + 1, begins[2] + 1, begins[3] + 1, ends[3] + 1]
    assert ctx.log.count('set_map') == 1 and ctx.log.count('set_strata') == 1 and ctx.log.count('plan') == 1
    assert [e for e in ctx.log if isinstance(e, tuple) and e[0] == 'plan_commit'] == 3 * [('plan_commit', [11, 12, 13, 14, 15, 16])]
    assert all(e == ('begin', True, True) for e in ctx.log if isinstance(e, tuple) and e[0] == 'begin')
    assert integ.sum_sigf == 400. and integ.map._device_owner is ctx
    # the next call starts from the pre-pass its predecessor left behind: no synchronous plan
    ctx.heads, ctx.log[:] = [(1.5, 1e-4, 400., 0)], []
    integ(f, nitn=1)
    assert 'plan' not in ctx.log and ctx.log[0] == ('plan_commit', [11, 12, 13, 14, 15, 16])
    # ... unless sigf was replaced in between
    ctx.heads, ctx.log[:] = [(1.5, 1e-4, 400., 0)], []
    integ.set(sigf=np.ones(integ.nhcube))
    integ(f, nitn=1)
    assert 'plan' in ctx.log and not any(isinstance(e, tuple) and e[0] == 'plan_commit' for e in ctx.log)

    # tolerances: booked at once, stops when met (sdev of the weighted average after k iterations: 1e-2 / sqrt(k))
    del order[:]
    integ, ctx = _scripted_integrator([(1.5, 1e-4, 400., 0) for _ in range(8)])
    r = integ(f, nitn=8, rtol=0.0045)
    assert len(r.itn_results) == 3 and len(ctx.heads) == 5
    assert [n for _, n in order] == [i + 1 for i, e in enumerate(ctx.log) if e == 'end']

    # NaN in the third iteration: two results booked, ValueError, sigf back to ones
    integ, ctx = _scripted_integrator([(1.5, 1e-4, 400., 0), (1.5, 1e-4, 400., 0), (float('nan'), 1e-4, 400., 1), (1.5, 1e-4, 400., 0)])
    del order[:]
    with pytest.raises(ValueError, match='nan'):
        integ(f, nitn=4)
    assert len(order) == 2 and len(ctx.heads) == 1
    assert integ.sum_sigf == integ.nhcube and float(integ._sigf_dev.min()) == float(integ._sigf_dev.max()) == 1.
