"""``restratify`` -- move the y-space strata to the axes where the integrand varies most
(reference: ``src/vegas/__init__.py:1313-1592``).

The reference measures, for every axis ``mu``, the profile ``dI[mu][i]`` = contribution to the
integral ``I`` from ``i/ndy <= y[mu] <= (i+1)/ndy`` by integrating an auxiliary integrand with
``1 + dim*ndy`` one-hot components (``_I_dI_integrand``, ``__init__.py:1390-1419``), which it builds on
the host from ``AdaptiveMap.invmap(x)``.  Here that integrand is never materialised: the sampler
writes ``x, wgt`` to HBM, the integrand's first component is evaluated there (device functor,
device callback, or a host integrand on copies), and two kernels consume the same buffers --
``vb200_reduce`` for ``I`` (and the ``sigf`` update, since the pass runs with ``adapt=True``,
``alpha=0``) and ``vb200_dy_profile`` for the ``dim*ndy`` profile components, which re-derives ``y``
from the Philox counter and performs the reference's per-hypercube two-pass per occupied bin.
"""
import numpy as np

from . import _lib
from ._gv import gv
from ._integrand import DeviceIntegrand
from ._integrator import Integrator, _dist
from ._results import VegasResult


class _ProfileIntegrand(object):
    """what ``VegasResult`` needs to know about the auxiliary integrand: a dictionary
    ``{'I': scalar, 'dI': [dim][ndy]}`` (``__init__.py:1414-1418``)"""

    def __init__(self, dim, ndy):
        self.shape = None
        self.bdict = gv.BufferDict()
        self.bdict['I'] = 0.0
        self.bdict['dI'] = np.zeros((dim, ndy), float)
        self.size = self.bdict.size

    def format_result(self, mean, var):
        return gv.BufferDict(self.bdict, buf=gv.gvar(mean, np.asarray(var) ** 0.5).reshape(-1))


def stratification_profile(integ, f, nitn=1, ndy=5):
    r""" ``nitn`` iterations of ``integ`` (``adapt=True``, ``alpha=0``: the map is left alone, ``sigf``
    is updated) that estimate ``I`` and the profile ``dI[mu][i]``; returns the ``RAvgDict`` with keys
    ``'I'`` and ``'dI'`` the reference obtains at ``__init__.py:1325-1330``. """
    if not 1 <= int(ndy) <= 32:
        raise ValueError('ndy = %s but require 1 <= ndy <= 32' % str(ndy))
    ndy = int(ndy)
    std = integ._make_std_integrand(f)
    device_fcn = f if (isinstance(f, DeviceIntegrand) and not integ.uses_jac and integ.dim <= _lib.MAX_FUSED_DIM) else None
    ctx, torch = integ._engine()
    dev = ctx.device
    rank, world = integ._rank_world()
    if device_fcn is not None:
        params, keep = device_fcn.params(integ.dim)
        nf_lib = ctx.set_integrand(device_fcn.fid, params, keep=(params, keep))
    dim = integ.dim
    yst = np.linspace(0, 1, ndy + 1)                      # __init__.py:1321
    result = VegasResult(_ProfileIntegrand(dim, ndy), weighted=True)
    for itn in range(int(nitn)):
        ctx, torch = integ._engine()
        acc = torch.zeros(3, dtype=torch.float64, device=dev)                 # I: mean, var; sum_sigf
        accd = torch.zeros(dim * ndy * 2, dtype=torch.float64, device=dev)    # dI: (mean, var) pairs
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        total, nmax, adaptive = integ._plan(ctx)
        flags = _lib.UPDATE_SIGF if (integ.beta > 0 and integ.nhcube > 1 and not integ.adapt_to_errors) else 0
        pitn = integ._next_itn()
        hs = int(integ.map.inc.shape[1])
        for c0, c1, rows in integ._batches(ctx, integ.max_batch):
            x = torch.empty((rows, dim), dtype=torch.float64, device=dev)
            wgt = torch.empty(rows, dtype=torch.float64, device=dev)
            jac1d = torch.empty_like(x) if integ.uses_jac else None
            ctx.sample(pitn, c0, c1, x, wgt, jac1d=jac1d)
            if device_fcn is not None:
                fx = torch.empty((rows, nf_lib), dtype=torch.float64, device=dev)
                ctx.eval_integrand(x, fx)
            elif std.on_device:
                fx = std.eval(x, jac=jac1d)
            else:
                jh = jac1d.cpu().numpy() if integ.uses_jac else None
                fh = np.asarray(std.eval(x.cpu().numpy(), jac=jh), dtype=float).reshape(rows, -1)
                fx = torch.from_numpy(np.ascontiguousarray(fh)).to(dev)
            f0 = fx.reshape(rows, -1)[:, 0].contiguous()                      # reference: eval(x)[:, 0], __init__.py:1406
            # the profile first: both kernels recompute the allocation from sigf, which reduce updates in place
            ctx.dy_profile(pitn, c0, c1, f0, 1, wgt, yst, accd)
            ctx.reduce(pitn, integ.beta, flags, c0, c1, f0, 1, wgt, integ._sigf_dev, acc, None, None, hs, status)
            integ._launches += 4
        if world > 1:
            dist = _dist()
            dist.all_reduce(acc)
            dist.all_reduce(accd)
            dist.all_reduce(status, op=dist.ReduceOp.MAX)
        integ._set_neval_stats(total, nmax, adaptive)
        if int(status.item()) != 0:
            raise ValueError('integrand evaluates to nan')
        a, ad = acc.cpu().numpy(), accd.cpu().numpy()
        mean = np.concatenate(([a[0]], ad[0::2]))
        var = np.concatenate(([a[1]], ad[1::2]))
        result.update(mean, var, integ.last_neval)
        if flags & _lib.UPDATE_SIGF:                       # pyx:2203-2211
            if a[2] > 0:
                integ.sum_sigf = float(a[2])
            else:
                if integ._sigf_dev is not None:
                    integ._sigf_dev.fill_(1.)
                integ.sum_sigf = integ._sigf_len
    return result.result


def new_stratification(old_nstrat, weight, gamma=1.0, below_avg_nstrat=None):
    r""" Strata per axis proportional to the axes' weights at (approximately) constant number of
    hypercubes, smallest first so that axes rounded down to 1 hand their share to the others
    (``__init__.py:1349-1365``). """
    old_nstrat = np.asarray(old_nstrat)
    dim = len(old_nstrat)
    w = np.asarray(gv.mean(weight), dtype=float)
    w_avg = np.average(w)
    w_gm = np.prod(w) ** (1. / dim)
    nstrat_gm = np.prod(old_nstrat) ** (1. / dim)
    nstrat = old_nstrat * ((w / w_gm) * nstrat_gm / old_nstrat) ** gamma
    order = np.array(np.argsort(nstrat))
    new_nstrat = np.array(old_nstrat)                     # same dtype as the integrator's array
    nleft = dim
    for mu in order:
        new_nstrat[mu] = nstrat[mu] if nstrat[mu] > 1 else 1
        if below_avg_nstrat and w[mu] < w_avg:
            new_nstrat[mu] = below_avg_nstrat
        nleft -= 1
        if nleft > 0:
            nstrat[order[-nleft:]] *= (nstrat[mu] / new_nstrat[mu]) ** (1. / nleft)
            nstrat[mu] = new_nstrat[mu]
    return new_nstrat


class restratifyIntegrator(Integrator):
    r""" Copy of ``integ`` whose strata have been redistributed for integrand ``f``
    (``__init__.py:1421-1434``).  Extra attributes: ``I``, ``dI[d][i]``, ``weight[d]``, ``Q``,
    ``old_nstrat``, ``ndy``, ``yst``. """

    def __init__(self, integ, f, nitn=1, ndy=5, below_avg_nstrat=None, verbose=False, gamma=1., **vargs):
        super(restratifyIntegrator, self).__init__(integ)
        for k in Integrator.engine_defaults:               # same device / seed stream / batch size
            if k not in vargs:
                setattr(self, k, getattr(integ, k))
        self._itn_counter = integ._itn_counter
        self.set(vargs)
        self.f = f
        self.ndy = int(ndy)
        self.yst = np.linspace(0, 1, self.ndy + 1)
        self.old_nstrat = np.array(self.nstrat)

        # I, dI with the old stratification
        result = stratification_profile(self, f, nitn=nitn, ndy=self.ndy)
        self.I, self.dI, self.Q = result['I'], result['dI'], result.Q
        if verbose:
            print('\n==================== restratify')
            print('BEFORE:')
            print(result.summary()[:-1])
            print('nstrat =', np.array2string(self.old_nstrat, max_line_width=60, prefix=9 * ' '))
            print('nhcube =', np.prod(self.old_nstrat), '\n')

        # weight[mu] = variance of the step function dI/dy[mu]   (__init__.py:1340-1344)
        dim = len(self.old_nstrat)
        self.weight = np.zeros(dim, object)
        dIavg = self.I / self.ndy
        for mu in range(dim):
            self.weight[mu] = np.sum((self.dI[mu] - dIavg) ** 2) * self.ndy
        if verbose:
            print('WEIGHTS:')
            for mu in range(dim):
                print('  weight %2d   %s' % (mu, self.weight[mu]))

        new_nstrat = new_stratification(self.old_nstrat, self.weight, gamma=gamma, below_avg_nstrat=below_avg_nstrat)
        self.set(nstrat=new_nstrat, neval=self.neval)
        # adaptive stratified sampling needs sigf on the new hypercubes: a few training iterations
        save = self.set(nitn=nitn, adapt=True)
        training = self(f)
        self.set(save)
        if verbose:
            print('\nAFTER:')
            print(training.summary()[:-1])
            print('nstrat =', np.array2string(np.asarray(self.nstrat), max_line_width=60, prefix=9 * ' '))
            print('nhcube =', np.prod(self.nstrat))
            print(20 * '=')


def restratify(integ, f=None, nitn=1, ndy=5, below_avg_nstrat=None, verbose=False, gamma=1., **vargs):
    r""" Return a |vegas| integrator with the y-space stratification optimised for integrand ``f``
    (same call as ``vegas.restratify``, ``__init__.py:1457-1592``)::

        integ = vegas.Integrator(20 * [[0, 1]])
        integ(f, neval=400_000, nitn=10)                            # adapt the map
        integ = vegas.restratify(integ, f, nitn=3, verbose=True)    # move the strata
        result = integ(f, nitn=10)

    ``nitn`` iterations of ``integ`` measure, for every axis ``d``, the variance ``weight[d]`` of the
    step function ``dI/dy[d]`` (``ndy`` steps; at most 32 here).  The new integrator has (approximately) as
    many hypercubes, with ``nstrat[d]`` proportional to ``weight[d]**gamma``; axes whose weight is below
    average get ``below_avg_nstrat`` strata if that is given.  ``vargs`` are further integrator settings.
    The result has the extra attributes ``I``, ``dI[d][i]`` and ``weight[d]``. """
    if f is None:
        raise ValueError('restratify needs the integrand f (PDFIntegrator is not part of this engine)')
    return restratifyIntegrator(integ, f=f, nitn=nitn, ndy=ndy, below_avg_nstrat=below_avg_nstrat, verbose=verbose,
                                gamma=gamma, **vargs)
