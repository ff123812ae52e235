"""``Integrator`` -- the reference's driver API (``_vegas.pyx:834-2230``) over the B200 engine.

What stays on the host (small, once per call / iteration): parameter handling and the integer
stratification set-up of ``set()`` (bit-exact with the reference), ``AdaptiveMap.adapt``, result
averaging.  What runs on the GPU: the whole per-iteration loop -- vegas+ allocation, Philox
sampling, stratified y, y->x map, integrand (device functor, or a batch callback fed from HBM),
per-hypercube two-pass mean/variance, ``sigf`` update and the map's training histogram.

With ``mpi=True`` and an initialised ``torch.distributed`` process group the hypercube index range
is sharded block-cyclically over the ranks (one GPU each) and the per-iteration sums are
all-reduced (NCCL over NVLink); every rank then performs the same host steps and returns the
same result.
"""
import os
import warnings

import numpy as np

from . import _lib
from ._gv import gv
from ._integrand import VegasIntegrand, DeviceIntegrand
from ._map import AdaptiveMap, HUGE
from ._results import VegasResult


def _dist():
    import torch.distributed as dist
    return dist


# reference settings that are accepted for API compatibility but change nothing here; each is
# reported once per process the first time it is given a non-default value
_NO_EFFECT = {
    'gpu_pad': 'batches are not padded: the engine sizes its own launches (pyx:2103-2106, 2130-2131)',
    'minimize_mem': 'sigf always lives in HBM; there is no h5py spill file (pyx:1430-1446)',
    'nproc': 'there is no multiprocessing pool: samples are generated and reduced on the GPU, Python '
             'integrands are evaluated in this process (pyx:1291-1307, 2108-2123)',
    'sync_ran': 'the Philox stream is a pure function of (seed, iteration, hypercube, sample); the seed is '
                'broadcast from rank 0',
}
_warned = set()


def _warn_no_effect(name):
    if name not in _warned:
        _warned.add(name)
        warnings.warn('vegas_b200: %s has no effect -- %s' % (name, _NO_EFFECT[name]), stacklevel=3)


class Integrator(object):
    r""" Adaptive multidimensional Monte Carlo integration (vegas / vegas+).

    Same constructor, ``set()``, ``__call__``, ``random_batch()`` ... and the same parameters as the
    reference class (docstring at ``_vegas.pyx:834-1112``).  Additional, engine-specific settings
    (not in ``defaults``): ``device`` (CUDA device index, default the current device), ``seed``
    (Philox key; default drawn from ``gvar.RNG`` so ``gvar.ranseed`` makes runs reproducible),
    ``fused`` (use the compiled device functor when the integrand has one; default True),
    ``max_batch`` (rows per batch handed to device callbacks), ``train_bins`` (callback path: pass the
    samples' training bins from the sampler to the reduce kernel through HBM, 2 bytes per axis,
    instead of replaying the Philox stream there; default True).
    """

    # Settings accessible via the constructor and Integrator.set (same keys as _vegas.pyx:1115-1140)
    defaults = dict(
        map=None,               # integration region, AdaptiveMap, or Integrator
        neval=1000,             # number of evaluations per iteration
        maxinc_axis=1000,       # number of adaptive-map increments per axis
        min_neval_batch=100000,  # min. number of evaluations per batch
        max_neval_hcube=50000,  # max number of evaluations per h-cube
        gpu_pad=False,          # pad batches for use by GPUs
        neval_frac=0.75,        # fraction of evaluations used for adaptive stratified sampling
        max_mem=1e9,            # memory cutoff (# of floats)
        nitn=10,                # number of iterations
        alpha=0.5,              # damping parameter for importance sampling
        beta=0.75,              # damping parameter for stratified sampliing
        adapt=True,             # flag to turn adaptation on or off
        correlate_integrals=True,  # calculate correlations between different integrals
        minimize_mem=False,     # minimize work memory (when neval very large)?
        adapt_to_errors=False,  # alternative approach to stratified sampling (low dim)?
        uniform_nstrat=False,   # require same nstrat[d] for all directions d?
        rtol=0,                 # relative error tolerance
        atol=0,                 # absolute error tolerance
        analyzer=None,          # analyzes results from each iteration
        ran_array_generator=None,  # alternative random number generator
        sync_ran=True,          # synchronize random generators across MPI processes?
        mpi=False,              # allow multi-GPU (torch.distributed)?
        uses_jac=False,         # return Jacobian to integrand?
        nproc=1,                # number of processors to use
    )
    # engine-specific settings (kept out of ``defaults`` so that dict mirrors the reference)
    engine_defaults = dict(device=None, seed=None, fused=True, max_batch=1 << 22, slab=None, train_bins=True)

    def __init__(self, map, **kargs):
        self.neval_hcube_range = None
        self.last_neval = 0
        self.pool = None
        self._ctx = None
        self._ctx_map_version = None
        self._ctx_strata = None
        self._sigf_dev = None
        self._sigf_layout = None     # (nhcube, slab, rank, world) of the device copy of sigf
        self._sigf_host = None
        self._sigf_len = 0
        self._itn_counter = 0
        self._launches = 0
        self._plan_ahead = None  # (key, statistics) of an allocation pre-pass launched ahead of the next iteration
        self._trace = None      # test hook: called with the raw per-iteration sums before adapt
        self._timing = None     # bench hook: list receiving per-iteration CUDA-event pairs
        self._unfused_events = []   # bench hook: (events around sample / callback / reduce, rows) per batch
        for k, v in self.engine_defaults.items():
            setattr(self, k, v)
        for k in list(kargs):
            if k in self.engine_defaults:
                setattr(self, k, kargs.pop(k))
        if isinstance(map, Integrator):
            self._set_map(map)
            args = {}
            for k in Integrator.defaults:
                if k != 'map':
                    args[k] = getattr(map, k)
            self._sigf_host = np.array(map.sigf)
            self._sigf_len = len(self._sigf_host)
            self.sum_sigf = np.sum(self._sigf_host)
            self.nstrat = np.array(map.nstrat)
            if 'nstrat' not in kargs:
                kargs['nstrat'] = map.nstrat
                if 'neval' not in kargs:
                    kargs['neval'] = map.neval
        else:
            self._sigf_host = np.array([], float)
            self._sigf_len = 0
            self.sum_sigf = HUGE
            args = dict(Integrator.defaults)
            del args['map']
            self._set_map(map)
            self.nstrat = np.full(self.map.dim, 0, dtype=np.intp)
        for k in Integrator.defaults:
            if k != 'map' and not hasattr(self, k):
                setattr(self, k, Integrator.defaults[k])
        args.update(kargs)
        if 'nstrat' in kargs and 'neval' not in kargs and 'neval' in args:
            del args['neval']
        if 'neval' in kargs and 'nstrat' not in kargs and 'nstrat' in args:
            del args['nstrat']
        self.set(args)

    # ------------------------------------------------------------------ pickling (pyx:1194-1207)
    def __reduce__(self):
        odict = dict()
        for k in Integrator.defaults:
            if k in ['map']:
                continue
            odict[k] = getattr(self, k)
        odict['nstrat'] = np.asarray(self.nstrat)
        odict['sigf'] = np.asarray(self.sigf)
        # engine settings and the Philox iteration counter: a restored integrator continues the stream
        # instead of replaying the uniforms of the iterations already done
        for k in Integrator.engine_defaults:
            if k != 'device':
                odict[k] = getattr(self, k)
        odict['_itn_counter'] = self._itn_counter
        return (Integrator, (self.map,), odict)

    def __setstate__(self, odict):
        odict = dict(odict)
        self._itn_counter = int(odict.pop('_itn_counter', 0))
        self.set(odict)

    # ------------------------------------------------------------------ sigf: host view of device state
    # (a new device tensor -- never an in-place update by the kernels -- voids a pre-pass launched ahead: it may have
    # been left waiting by the last iteration of an earlier call)
    @property
    def _sigf_dev(self):
        return self.__dict__.get('_sigf_dev_t')

    @_sigf_dev.setter
    def _sigf_dev(self, t):
        self.__dict__['_sigf_dev_t'] = t
        self._plan_ahead = None

    def _get_sigf(self):
        if self._sigf_dev is not None:
            local = self._sigf_dev.cpu().numpy()
            # the layout the device copy was created with (set() may have changed slab / mpi since)
            nh, slab, rank, world = self._sigf_layout
            if world == 1:
                return local
            return self._gather_sigf(local, nh, slab, rank, world)
        if self._sigf_host is not None:
            return self._sigf_host
        return np.ones(self._sigf_len, float)

    sigf = property(_get_sigf, doc="sigf[h] = |variance of hypercube h|**(beta/2) (host copy)")

    def _gather_sigf(self, local, nh, slab, rank, world):
        import torch
        dist = _dist()
        counts = [len(_local_cubes(nh, slab, r, world)) for r in range(world)]
        bufs = [torch.empty(c, dtype=torch.float64, device=self._sigf_dev.device) for c in counts]
        dist.all_gather(bufs, self._sigf_dev)
        out = np.empty(nh, float)
        for r in range(world):
            out[_local_cubes(nh, slab, r, world)] = bufs[r].cpu().numpy()
        return out

    def _set_map(self, map):
        r""" install new map, create xsample (pyx:1209-1253) """
        self._set_map_checked(map)
        if self.map.dim > _lib.MAXDIM:
            raise ValueError('vegas_b200 integrates up to %d dimensions (the kernels carry per-axis parameters in '
                             'fixed-size blocks, VB_MAXD); this map has %d' % (_lib.MAXDIM, self.map.dim))

    def _set_map_checked(self, map):
        if isinstance(map, AdaptiveMap):
            self.map = AdaptiveMap(map)
            self.xsample = np.empty(self.map.dim, dtype=float)
            for d in range(self.map.dim):
                self.xsample[d] = gv.RNG.uniform(*self.map.region(d))
        elif isinstance(map, Integrator):
            self.map = AdaptiveMap(map.map)
            self.xsample = gv.BufferDict(map.xsample) if map.xsample.shape is None else np.array(map.xsample)
        elif hasattr(map, 'keys'):
            map = gv.asbufferdict(map)
            self.xsample = gv.BufferDict()
            limits = []
            for k in map:
                shape = np.shape(map[k])[:-1]
                if shape == ():
                    self.xsample[k] = gv.RNG.uniform(*map[k])
                    limits.append(map[k])
                else:
                    tmp = np.empty(shape, dtype=float)
                    for idx in np.ndindex(shape):
                        tmp[idx] = gv.RNG.uniform(*map[k][idx])
                    self.xsample[k] = tmp
                    limits += np.array(map[k]).reshape(-1, 2).tolist()
            self.map = AdaptiveMap(limits)
        else:
            map = np.array(map, dtype=object)
            if np.shape(map.flat[0]) == ():
                self.xsample = np.empty(map.shape[:-1], dtype=float)
                grid = map.reshape(-1, 2)
            else:
                self.xsample = np.empty(map.shape, dtype=float)
                grid = map.reshape(-1)
            self.map = AdaptiveMap(grid)
            for i, idx in enumerate(np.ndindex(self.xsample.shape)):
                self.xsample[idx] = gv.RNG.uniform(*self.map.region(i))

    # ------------------------------------------------------------------ set (pyx:1256-1446)
    def set(self, ka={}, **kargs):
        r""" Reset default parameters in integrator; returns the dictionary of the old values so
        that ``integ.set(old_defaults)`` restores them. """
        if kargs:
            kargs.update(ka)
        else:
            kargs = ka
        old_val = dict()
        nstrat = None
        for k in kargs:
            if k == 'map':
                old_val['map'] = self.map
                self._set_map(kargs['map'])
            elif k == 'nstrat':
                if kargs['nstrat'] is None:
                    continue
                old_val['nstrat'] = self.nstrat
                nstrat = np.array(kargs['nstrat'], dtype=np.intp)
            elif k == 'sigf':
                old_val['sigf'] = self.sigf
                self._sigf_host = np.fabs(np.asarray(kargs['sigf'], dtype=float))
                self._sigf_len = len(self._sigf_host)
                self._sigf_dev = None
                self.sum_sigf = np.sum(self._sigf_host)
            elif k == 'nproc':
                old_val['nproc'] = self.nproc
                self.nproc = kargs['nproc'] if kargs['nproc'] is not None else os.cpu_count()
                if self.nproc is None:
                    self.nproc = 1
                if self.nproc != 1:
                    _warn_no_effect('nproc')
            elif k in Integrator.defaults:
                old_val[k] = getattr(self, k)
                try:
                    setattr(self, k, kargs[k])
                except Exception:
                    setattr(self, k, type(old_val[k])(kargs[k]))
                if k in _NO_EFFECT and kargs[k] != Integrator.defaults[k]:
                    _warn_no_effect(k)
            elif k in Integrator.engine_defaults:
                old_val[k] = getattr(self, k)
                setattr(self, k, kargs[k])
                if k in ('device', 'slab'):
                    self._ctx = None
                    self._ctx_strata = None
            elif k not in ['nhcube_batch', 'max_nhcube']:
                raise AttributeError('no parameter named "%s"' % str(k))

        # 2) sanity checks
        if nstrat is not None:
            if len(nstrat) != self.map.dim:
                raise ValueError('nstrat[d] has wrong length: %d not %d' % (len(nstrat), self.map.dim))
            if np.any(nstrat < 1):
                raise ValueError('bad nstrat: ' + str(np.asarray(self.nstrat)))
        if self.neval_frac < 0 or self.neval_frac >= 1:
            raise ValueError('neval_frac = {} but require 0 <= neval_frac < 1'.format(self.neval_frac))
        if 'neval' in old_val and self.neval < 2:
            raise ValueError('neval>2 required, not ' + str(self.neval))
        neval_frac = 0 if (self.beta == 0 or self.adapt_to_errors) else self.neval_frac

        self.dim = self.map.dim

        # 3) determine # strata in each direction
        if nstrat is not None:
            if len(nstrat) != self.dim or min(nstrat) < 1:
                raise ValueError('bad nstrat = %s' % str(np.asarray(nstrat)))
            nhcube = np.prod(nstrat)
            if 'neval' not in old_val:
                old_val['neval'] = self.neval
                self.neval = type(self.neval)(2. * nhcube / (1. - neval_frac))
            elif self.neval < 2. * nhcube / (1. - neval_frac):
                raise ValueError('neval too small: {} < {}'.format(self.neval, 2. * nhcube / (1. - neval_frac)))
        elif 'neval' in old_val or 'neval_frac' in old_val:
            ns = int(abs((1 - neval_frac) * self.neval / 2.) ** (1. / self.dim))     # strata / axis
            if ns < 1:
                ns = 1
            d = int((np.log((1 - neval_frac) * self.neval / 2.) - self.dim * np.log(ns)) / np.log(1 + 1. / ns))
            if ((ns + 1) ** d * ns ** (self.dim - d)) > self.max_mem and not self.minimize_mem:
                raise MemoryError("work arrays larger than max_mem; set minimize_mem=True or increase max_mem")
            if self.uniform_nstrat:
                d = 0
            nstrat = np.empty(self.dim, np.intp)
            nstrat[:d] = ns + 1
            nstrat[d:] = ns
        else:
            nstrat = self.nstrat

        # 4) reconfigure vegas map, if necessary
        if self.adapt_to_errors:
            self.map.adapt(ninc=np.asarray(nstrat))
        else:
            ni = min(int(self.neval / 10.), self.maxinc_axis)     # increments/axis
            ninc = np.empty(self.dim, np.intp)
            for d in range(self.dim):
                if ni >= nstrat[d]:
                    ninc[d] = int(ni / nstrat[d]) * nstrat[d]
                elif nstrat[d] <= self.maxinc_axis:
                    ninc[d] = nstrat[d]
                else:
                    nstrat[d] = int(nstrat[d] / ni) * ni
                    ninc[d] = ni
            if not np.all(np.equal(self.map.ninc, ninc)):
                self.map.adapt(ninc=ninc)

        if not np.all(np.equal(self.nstrat, nstrat)):
            if 'sigf' not in old_val:
                old_val['sigf'] = self.sigf
                self._sigf_host = np.array([], float)
                self._sigf_dev = None
                self._sigf_len = 0
                self.sum_sigf = HUGE
            self.nstrat = nstrat

        # 5) set min_neval_hcube
        self.nhcube = int(np.prod(self.nstrat, dtype=np.int64))
        if self.nhcube == 1:
            self.min_neval_hcube = int(self.neval)
        else:
            self.min_neval_hcube = int((1 - neval_frac) * self.neval / self.nhcube)
        if self.min_neval_hcube < 2:
            self.min_neval_hcube = 2

        # 6) sigf (lives in HBM once sampling starts; "all ones" is represented lazily)
        nsigf = self.nhcube
        if self.beta >= 0 and self.nhcube > 1 and not self.adapt_to_errors and self._sigf_len != nsigf:
            self._sigf_host = None
            self._sigf_dev = None
            self._sigf_len = nsigf
            self.sum_sigf = nsigf
        workspace = self.min_neval_batch + self.max_neval_hcube
        if workspace > self.neval:
            workspace = self.neval + 1
        if (3 * self.dim + 3) * workspace + (0 if self.minimize_mem else self.nhcube) > self.max_mem:
            raise MemoryError('work arrays larger than max_mem; reduce min_neval_batch or max_neval_hcube (or increase max_mem)')
        return old_val

    # ------------------------------------------------------------------ settings (the text of pyx:1448-1590)
    def _axis_labels(self):
        """key/index label of every axis in xsample's order (None for a plain list of limits): ``key`` for a
        number, ``key i,j`` -- the key on the first element only -- for an array, ``i,j`` for an index array"""
        def index_text(idx):
            return ','.join(str(i) for i in idx)
        xs = self.xsample
        if xs.shape is None:
            labels = []
            for k in xs:
                shape = np.shape(xs[k])
                if shape == ():
                    labels.append(str(k))
                else:
                    labels.extend(('%s %s' % (k, index_text(idx))) if n == 0 else index_text(idx)
                                  for n, idx in enumerate(np.ndindex(shape)))
            return labels
        if len(xs.shape) > 1:
            return [index_text(idx) for idx in np.ndindex(xs.shape)]
        return None

    def settings(self, ngrid=0):
        r""" Assemble summary of integrator settings into string. """
        vegas_plus = self.beta > 0 and not self.adapt_to_errors
        nhcube = np.prod(self.nstrat)
        per_iteration = self.neval if self.beta > 0 else nhcube * self.min_neval_hcube
        alpha, beta = (0., 0.) if not self.adapt else (self.alpha, 0. if self.adapt_to_errors else self.beta)

        def ints(a, label):         # (continuation lines of a long array line up behind the label)
            return np.array2string(np.asarray(a), max_line_width=80, prefix=len(label) * ' ')
        strata, incs = '    number of: strata/axis = ', '               increments/axis = '
        lines = [
            'Integrator Settings:',
            '    %.6g%s integrand evaluations in each of %d iterations' % (per_iteration, ' (approx)' if vegas_plus else '', self.nitn),
            strata + ints(self.nstrat, strata),
            incs + ints(self.map.ninc, incs),
            '               h-cubes = %.6g  processors = %d' % (nhcube, self.nproc),
            '               evaluations/batch >= %.2g' % float(self.min_neval_batch),
            '               %d <= evaluations/h-cube <= %.2g' % (self.min_neval_hcube, float(max(self.max_neval_hcube, self.min_neval_hcube))),
            '    minimize_mem = %s  adapt_to_errors = %s  adapt = %s' % (self.minimize_mem, self.adapt_to_errors, self.adapt),
            '    accuracy: relative = %g  absolute = %g' % (self.rtol, self.atol),
            '    damping: alpha = %g  beta= %g' % (alpha, beta),
            '']
        # the table of integration limits: right-aligned columns [key/index,] axis, limits under a ruled header;
        # beyond 20 axes two such blocks side by side, the first one taking the extra row of an odd count
        labels = self._axis_labels()
        rows = [(str(d), '({:.5}, {:.5})'.format(float(lo), float(hi))) for d, (lo, hi) in enumerate(self.map.region())]
        header = ('axis', 'integration limits')
        if labels is not None:
            rows, header = [(lab,) + r for lab, r in zip(labels, rows)], ('key/index',) + header
        widths = [max(len(r[c]) for r in rows + [header]) for c in range(len(header))]

        def text(row):
            return '    '.join(cell.rjust(w) for cell, w in zip(row, widths))
        nblock = 1 if self.map.dim <= 20 else 2
        per_block = -(-len(rows) // nblock)
        blocks = []
        for b in range(nblock):
            body = [text(r) for r in rows[b * per_block:(b + 1) * per_block]]
            blocks.append([text(header), len(text(header)) * '-'] + body)
        for i in range(per_block + 2):
            lines.append('    ' + '  '.join(blk[i] for blk in blocks if i < len(blk)))
        ans = '\n'.join(lines) + '\n'
        if ngrid > 0:
            ans += '\n' + self.map.settings(ngrid=ngrid)
        return ans

    def _get_mpi_rank(self):
        return self._rank_world()[0]

    mpi_rank = property(_get_mpi_rank, doc="rank (>=0) in the torch.distributed group")

    @staticmethod
    def synchronize_random():
        """kept for API compatibility: Philox streams are a pure function of (seed, itn, hypercube,
        sample), and the seed is broadcast from rank 0 -- nothing else to synchronise"""

    # ------------------------------------------------------------------ device plumbing
    def _rank_world(self):
        if self.mpi:
            try:
                dist = _dist()
                if dist.is_available() and dist.is_initialized():
                    return dist.get_rank(), dist.get_world_size()
            except ImportError:
                pass
        return 0, 1

    def _slab(self, world):
        if self.slab is not None:
            return int(self.slab)
        if world == 1:
            return _lib.CHUNK
        # At least 512 slabs per rank, of 256 ... 16384 hypercubes: the vegas+ allocation varies smoothly over the
        # hypercube index, so many small slabs balance the SAMPLES of the ranks, not just their cubes.  Measured on
        # 8 GPUs (tools/skew_probe.py, 8-D ridge at neval = 1e8 fixed, 1.1e7 hypercubes): slabs of 16384 cubes (86 per
        # rank) left the ranks' sample counts 5.2 % apart (max / mean) and the step at 26.0 ms; 2048 (686 per rank):
        # 0.8 %, 25.2 ms; 512: 25.2 ms.  Large problems keep large slabs (config 5, 7.8e8 hypercubes: 346 ms per step
        # with 16384-cube slabs, 376 ms with 2048 -- windows move at every slab boundary).
        per = -(-int(self.nhcube) // (world * 512))
        # whole chunks of the light geometry (512 cubes) unless the problem is too small to give every rank two of them
        unit = 2 * _lib.CHUNK if int(self.nhcube) >= 4 * _lib.CHUNK * world else _lib.CHUNK
        per = -(-per // unit) * unit
        return int(max(_lib.CHUNK, min(64 * _lib.CHUNK, per)))

    def _engine(self):
        """bring the device context up to date with map / strata / sigf; returns (ctx, torch)"""
        import torch
        if self._ctx is None:
            dev = self.device
            if dev is None and torch.cuda.is_available():
                dev = torch.cuda.current_device()
            self._ctx = _lib.Context(dev)
            self._ctx_map_version = None
            self._ctx_strata = None
            self._sigf_dev_stale()
        ctx = self._ctx
        rank, world = self._rank_world()
        if self.seed is None:
            seed = int(gv.RNG.integers(1, 2 ** 62))
            if world > 1:
                t = torch.tensor([seed], dtype=torch.int64, device=ctx.device)
                _dist().broadcast(t, 0)
                seed = int(t.item())
            self.seed = seed
        ctx.set_seed(self.seed)
        mv = (id(self.map), self.map._version)
        if self._ctx_map_version != mv:
            ctx.set_map(self.map.grid, self.map.ninc)
            self._ctx_map_version = mv
        key = (tuple(int(v) for v in self.nstrat), rank, world, self._slab(world))
        if self._ctx_strata != key:
            self._nlocal = ctx.set_strata(self.nstrat, key[3], rank, world)
            self._ctx_strata = key
            self._sigf_dev_stale()
        # sigf
        need_sigf = self.beta >= 0 and self.nhcube > 1 and not self.adapt_to_errors
        if need_sigf and self._sigf_dev is None:
            if self._sigf_host is not None and len(self._sigf_host) == self.nhcube:
                full = np.ascontiguousarray(self._sigf_host, dtype=float)
                local = full[_local_cubes(self.nhcube, key[3], rank, world)] if world > 1 else full
                self._sigf_dev = torch.from_numpy(np.ascontiguousarray(local)).to(ctx.device)
            else:
                self._sigf_dev = torch.ones(self._nlocal, dtype=torch.float64, device=ctx.device)
            self._sigf_host = None
            self._sigf_layout = (int(self.nhcube), key[3], rank, world)
        return ctx, torch

    def _sigf_dev_stale(self):
        if self._sigf_dev is not None:
            self._sigf_host = self._get_sigf()
            self._sigf_dev = None

    def _plan_args(self):
        adaptive = self.beta > 0 and self.nhcube > 1 and not self.adapt_to_errors
        neval_sigf = (self.neval_frac * self.neval / self.sum_sigf
                      if self.beta > 0 and self.sum_sigf > 0 and not self.adapt_to_errors else 0.0)
        max_nh = max(self.max_neval_hcube, self.min_neval_hcube)
        return adaptive, neval_sigf, max_nh, int(self.neval / self.nhcube)

    def _plan_key(self, ctx, neval_sigf, max_nh, uniform):
        """everything the allocation pre-pass depends on: a pre-pass launched ahead of time is used only
        when none of it has changed since"""
        return (id(ctx), getattr(ctx, 'integrand_serial', 0), self._ctx_strata,
                None if self._sigf_dev is None else self._sigf_dev.data_ptr(),
                float(neval_sigf), int(self.min_neval_hcube), int(max_nh), int(uniform))

    def _plan(self, ctx, neval_hcube_out=None):
        """vegas+ allocation for the coming iteration (pyx:1657-1706) -> (local total, max).  The
        pre-pass over sigf is normally already done: it was launched behind the previous iteration's
        kernels (``_plan_next``) and its statistics came back with that iteration's results."""
        adaptive, neval_sigf, max_nh, uniform = self._plan_args()
        ahead, self._plan_ahead = self._plan_ahead, None
        if (ahead is not None and adaptive and neval_hcube_out is None
                and ahead[0] == self._plan_key(ctx, neval_sigf, max_nh, uniform)):
            total, nmin, nmax, nchunks = ctx.plan_commit(neval_sigf, ahead[1])
        else:
            total, nmin, nmax, nchunks = ctx.plan(self._sigf_dev if adaptive else None, neval_sigf,
                                                 self.min_neval_hcube, max_nh, uniform, neval_hcube_out)
        self._nchunks = nchunks
        return total, nmax, adaptive

    def _plan_next(self, ctx, sum_sigf_dev, stats_dev):
        """launch the pre-pass of the NEXT iteration behind this one's kernels (``vb200_plan_ahead``):
        the new sigf and sum_sigf are final on the device, so the host round trip ``vb200_plan`` needs
        (launch, copy back, synchronise) disappears from the iteration"""
        adaptive, _, max_nh, uniform = self._plan_args()
        ctx.plan_ahead(self._sigf_dev, sum_sigf_dev, self.neval_frac * self.neval, self.min_neval_hcube, max_nh,
                       uniform, stats_dev)
        self._launches += 1

    def _flags(self, nf):
        f = 0
        if self.beta > 0 and self.nhcube > 1 and self.adapt and not self.adapt_to_errors:
            f |= _lib.UPDATE_SIGF
        if self.adapt_to_errors and self.adapt:
            f |= _lib.TRAIN_ERRORS
        elif self.adapt and self.alpha > 0:
            f |= _lib.TRAIN
        if self.correlate_integrals and nf > 1:
            f |= _lib.CORRELATE
        return f

    def _batches(self, ctx, target_rows):
        """ranges of local chunks [c0, c1) holding >= target_rows samples each (last one smaller)"""
        off = ctx.chunk_offsets(self._nchunks + 1)
        out, c0 = [], 0
        while c0 < self._nchunks:
            c1 = int(np.searchsorted(off, off[c0] + target_rows, side='left'))
            c1 = min(max(c1, c0 + 1), self._nchunks)
            out.append((c0, c1, int(off[c1] - off[c0])))
            c0 = c1
        return out

    # ------------------------------------------------------------------ sampling API (pyx:1601-1863)
    def random_batch(self, yield_hcube=False, yield_y=False):
        r""" Low-level batch iterator over integration points and weights: yields ``x, wgt`` (plus
        ``y`` and/or ``hcube`` on request) as numpy arrays, one iteration's worth in total, in
        hypercube order (``_vegas.pyx:1601-1634``).  This rank's share when sharded. """
        for t in self._random_batch(yield_hcube, yield_y, device=False):
            yield t

    def random_batch_device(self, yield_hcube=False, yield_y=False):
        r""" Same as :meth:`random_batch` but yields CUDA ``torch`` tensors (no host copies). """
        for t in self._random_batch(yield_hcube, yield_y, device=True):
            yield t

    def _random_batch(self, yield_hcube, yield_y, device):
        ctx, torch = self._engine()
        total, nmax, adaptive = self._plan(ctx)
        self._set_neval_stats(total, nmax, adaptive)
        itn = self._next_itn()
        target = self.max_batch if device else self.min_neval_batch
        for c0, c1, rows in self._batches(ctx, target):
            x = torch.empty((rows, self.dim), dtype=torch.float64, device=ctx.device)
            wgt = torch.empty(rows, dtype=torch.float64, device=ctx.device)
            y = torch.empty_like(x) if yield_y else None
            hc = torch.empty(rows, dtype=torch.int64, device=ctx.device) if yield_hcube else None
            ctx.sample(itn, c0, c1, x, wgt, y=y, hcube=hc, u=self._injected_uniforms(torch, ctx, rows))
            self._launches += 1
            ans = (x,)
            if yield_y:
                ans += (y,)
            ans += (wgt,)
            if yield_hcube:
                ans += (hc,)
            yield ans if device else tuple(t.cpu().numpy() for t in ans)

    random_vec = random_batch

    def random(self, yield_hcube=False, yield_y=False):
        r""" Low-level iterator over single integration points and weights (``_vegas.pyx:1770-1814``). """
        for t in self.random_batch(yield_hcube=yield_hcube, yield_y=yield_y):
            for i in range(t[0].shape[0]):
                yield tuple(ti[i] for ti in t)

    def sample(self, nbatch=None, mode='rbatch'):
        r""" Generate random sample of integration weights and points: ``wgt, x = integ.sample()`` with
        ``sum(wgt * f(x))`` an estimate of the integral (``_vegas.pyx:1816-1863``). """
        neval = self.last_neval if self.last_neval > 0 else self.neval
        nbatch = neval if nbatch is None else int(nbatch)
        nit = nbatch // neval
        if nit * neval < nbatch:
            nit += 1
        samples, wgts = [], []
        for _ in range(nit):
            for x, w in self.random_batch():
                samples.append(np.array(x))
                wgts.append(np.array(w))
        samples = np.concatenate(samples, axis=0)
        wgts = np.concatenate(wgts) / nit
        if self.xsample.shape is None:
            if mode == 'rbatch':
                samples = gv.BufferDict(self.xsample, rbatch_buf=samples.T)
            else:
                samples = gv.BufferDict(self.xsample, lbatch_buf=samples)
        elif self.xsample.shape != ():
            if mode == 'rbatch':
                samples = samples.T
                samples.shape = self.xsample.shape + (-1,)
            else:
                samples.shape = (-1,) + self.xsample.shape
        return wgts, samples

    def _injected_uniforms(self, torch, ctx, rows):
        """uniforms of one batch from the user's ``ran_array_generator`` (pyx:1081-1086, 1676-1680, 1732:
        called once per batch with the shape ``(rows, dim)``; rows are consumed in hypercube order), as a
        device tensor -- or None when the engine's Philox stream is used"""
        if self.ran_array_generator is None:
            return None
        if self._rank_world()[1] > 1:
            raise NotImplementedError('ran_array_generator is not supported with the hypercube range sharded over GPUs')
        u = np.ascontiguousarray(self.ran_array_generator((rows, self.dim)), dtype=float)
        if u.shape != (rows, self.dim):
            raise ValueError('ran_array_generator returned shape %s, expected %s' % (u.shape, (rows, self.dim)))
        return torch.from_numpy(u).to(ctx.device)

    def _next_itn(self):
        self._itn_counter += 1
        return self._itn_counter & 0xFFFFFF

    def _set_neval_stats(self, total, nmax, adaptive, reduced=False):
        rank, world = self._rank_world()
        if world > 1 and not reduced:
            total, nmax = allreduce_neval_stats(total, nmax, self._ctx.device)
        self.last_neval = int(total)
        # pyx:1682,1699-1702: both ends start at min_neval_hcube
        hi = max(self.min_neval_hcube, nmax) if adaptive else self.min_neval_hcube
        self.neval_hcube_range = np.array([self.min_neval_hcube, hi], dtype=np.intp)

    def _make_std_integrand(self, fcn, xsample=None):
        r""" Convert integrand ``fcn`` into an lbatch integrand (:class:`VegasIntegrand`). """
        if isinstance(fcn, VegasIntegrand):
            return fcn
        return VegasIntegrand(fcn=fcn, map=self.map, uses_jac=self.uses_jac,
                              xsample=self.xsample if xsample is None else xsample, mpi=False)

    # ------------------------------------------------------------------ the iteration loop (pyx:1912-2230)
    def __call__(self, fcn, save=None, saveall=None, **kargs):
        r""" Integrate integrand ``fcn`` (reference docstring at ``_vegas.pyx:1913-2035``).

        ``fcn`` may be: a plain Python function of one point; an ``@lbatchintegrand`` /
        ``@rbatchintegrand`` (numpy, evaluated on the host from samples generated on the GPU); a
        ``@devicebatchintegrand`` (torch CUDA tensors in HBM); or a ``DeviceIntegrand`` whose functor
        is compiled into the library (fused kernel, nothing leaves the SMs).  Returns ``RAvg`` /
        ``RAvgArray`` / ``RAvgDict``. """
        if kargs:
            self.set(kargs)
        # built-in functors are compiled for up to 20 dimensions; above that their numpy twins run
        # through the callback path like any other lbatch integrand
        device_fcn = fcn if (isinstance(fcn, DeviceIntegrand) and self.fused and not self.uses_jac
                             and self.dim <= _lib.MAX_FUSED_DIM and self.ran_array_generator is None) else None
        std = self._make_std_integrand(fcn)
        nf = std.size
        ctx, torch = self._engine()
        dev = ctx.device
        rank, world = self._rank_world()
        if device_fcn is not None:
            params, keep = device_fcn.params(self.dim)
            nf_lib = ctx.set_integrand(device_fcn.fid, params, keep=(params, keep))
            if nf_lib != nf:
                raise ValueError('device integrand has %d components, its numpy twin %d' % (nf_lib, nf))
        nv = nf * (nf + 1) // 2
        result = VegasResult(std, weighted=self.adapt)
        result.is_writer = world == 1 or rank == 0

        def book(acc_h, neval, trace=None):
            """one iteration's [mean, cov] into the running average"""
            mean = acc_h[:nf].copy()
            if self.correlate_integrals:
                var = np.zeros((nf, nf), float)
                var[_tril(nf)] = acc_h[nf:nf + nv]
                var = var + np.tril(var, -1).T
            else:
                var = acc_h[nf:nf + nv][[s * (s + 1) // 2 + s for s in range(nf)]].copy()
            if trace is not None:
                trace(mean, var)
            result.update(mean, var, neval)

        # One-call iterations with nothing looking at the running average between them (no tolerances, analyzer or
        # save files): iteration i's results are booked while the kernels of iteration i + 1 run
        # (vb200_iteration_begin / _end) -- at everyday sizes the Python bookkeeping is a third of an iteration.
        defer = (self.rtol == 0 and self.atol == 0 and self.analyzer is None and save is None and saveall is None
                 and not os.environ.get('VB200_NO_DEFER'))
        pending = None
        env_host_adapt, env_no_ahead, env_no_fast = (os.environ.get(k) for k in           # (developer switches)
                                                     ('VB200_HOST_ADAPT', 'VB200_NO_PLAN_AHEAD', 'VB200_NO_FAST_ITERATION'))
        fast = False
        for itn in range(self.nitn):
            if self.analyzer is not None:
                self.analyzer.begin(itn, self)
            if not (fast and self.analyzer is None):
                ctx, torch = self._engine()      # map / sigf may have changed (not behind a one-call iteration: the
                #                                  device adapted its own grid and nothing else ran in between)
            hs = int(self.map._inc.shape[1])       # (the shape only: must not pull a device-adapted grid back every iteration)
            # One allocation and one device-to-host copy per iteration:
            #   buf_f (fp64):  [mean, cov, sum_sigf | sum_f | n_f as fp64 | samples, NaN count, max samples per
            #                   hypercube of rank 0 .. world-1]   -- the part a sharded run all-reduces (SUM), once
            #   buf_i (int64): [n_f | NaN flag | statistics of the next iteration's allocation pre-pass]
            nacc, nh = nf + nv + 1, self.dim * hs
            n_bf, n_bi = nacc + 2 * nh + 2 + world, nh + 1 + 6
            raw = getattr(self, '_raw', None)          # one device buffer for the integrator's lifetime, zeroed per iteration
            if raw is None or raw.numel() != 8 * (n_bf + n_bi) or raw.device != torch.device(dev):
                raw = self._raw = torch.empty(8 * (n_bf + n_bi), dtype=torch.uint8, device=dev)

            def views():
                raw.zero_()
                bf = raw[:8 * n_bf].view(torch.float64)
                bi = raw[8 * n_bf:].view(torch.int64)
                return (bf, bf[:nacc], bf[nacc:nacc + nh].view(self.dim, hs), bi[:nh].view(self.dim, hs),
                        bi[nh:nh + 1].view(torch.int32), bi[nh + 1:])       # (status: the kernels set its low word)
            if self._timing is not None:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                ev[0].record()
            total, nmax, adaptive = self._plan(ctx)
            flags = self._flags(nf)
            pitn = self._next_itn()
            if self._timing is not None:
                ev[1].record()
            # AdaptiveMap.adapt on the device, behind the kernels (the grid stays in HBM; the host copy is refreshed
            # when somebody looks at integ.map.grid).  The host route remains for everything the kernel does not
            # cover: analyzers / trace hooks that want the histogram, adapt_to_errors, alpha <= 0, single-increment
            # axes, training data added by hand.
            dev_adapt = (bool(flags & _lib.TRAIN) and self.alpha > 0 and self.adapt and self.analyzer is None
                         and self._trace is None and self.map.sum_f is None and int(np.min(self.map.ninc)) > 1
                         and not env_host_adapt)
            # the next iteration's allocation pre-pass rides behind this one (sum_sigf is final on the device) -- behind
            # the last one of a call as well: the next call usually continues with the same settings (_plan_key decides),
            # and a 4 us kernel here saves it the synchronous pre-pass, a third of a one-iteration call at everyday sizes
            plan_next = bool((flags & _lib.UPDATE_SIGF) and self._sigf_dev is not None and not env_no_ahead)
            # Everyday sizes: the whole iteration in one library call (vb200_iteration: zero, engine, adapt, pre-pass
            # of the next iteration, one small copy back) -- a handful of binding calls cost more than the kernels.
            fast = (device_fcn is not None and world == 1 and self._timing is None and self._trace is None
                    and (dev_adapt or not (flags & (_lib.TRAIN | _lib.TRAIN_ERRORS)))
                    and not env_no_fast)
            head = None
            if fast:
                head = np.empty(nacc + 7, dtype=np.float64)
                _, _, max_nh, uniform = self._plan_args()
                try:
                    ctx.iteration_begin(pitn, self.beta, flags, self._sigf_dev, raw, nacc, nh, hs, n_bf, n_bf + n_bi,
                                        self.alpha if dev_adapt else 0.,
                                        (self.neval_frac * self.neval, self.min_neval_hcube, max_nh, uniform) if plan_next else None)
                    if pending is not None:          # (the previous iteration's bookkeeping, behind this one's kernels)
                        book(*pending)
                        pending = None
                    ctx.iteration_end(head)
                except _lib.VegasB200Error as err:
                    if getattr(err, 'code', 0) != -4:
                        raise
                    device_fcn, fast, head = None, False, None      # no fused instantiation: callback path from here on
            if not fast:
                buf_f, acc, sum_f, n_f, status, stats_next = views()
            if fast:
                pass
            elif device_fcn is not None:
                try:
                    ctx.iterate_fused(pitn, self.beta, flags, self._sigf_dev, acc, sum_f, n_f, hs, status)
                    self._launches += 2
                except _lib.VegasB200Error as err:
                    if getattr(err, 'code', 0) != -4:
                        raise
                    device_fcn = None          # no fused instantiation for this dimension: callback path from here on
            if device_fcn is None:
                self._iterate_unfused(ctx, torch, std, pitn, flags, acc, sum_f, n_f, hs, status)
            if self._timing is not None:
                ev[2].record()
            if world > 1:
                pack_iteration(buf_f, nacc, nh, n_f, status, total, nmax, rank)
                exchange_iteration(buf_f)
            if dev_adapt and not fast:
                if world > 1:       # all-reduced counts (fp64) and NaN count of all ranks: every rank decides alike
                    counts = buf_f[nacc + nh:nacc + 2 * nh].view(self.dim, hs)
                    nan_any = (buf_f[nacc + 2 * nh + 1:nacc + 2 * nh + 2] != 0).to(torch.int32)
                else:
                    counts, nan_any = n_f, status
                ctx.map_adapt_device(sum_f, counts, hs, self.alpha, nan_any)
                self._launches += 1
            if plan_next and not fast:
                self._plan_next(ctx, acc[nf + nv:], stats_next)
            if self._timing is not None:
                ev[3].record()
                self._timing.append((ev, total))
            if fast:
                hf, tail_i = head, head[nacc:].view(np.int64)
                nan_seen, stats6 = (int(tail_i[0]) & 0xffffffff) != 0, tail_i[1:7].copy()
                n_f_h = None
            else:
                hraw = raw.cpu().numpy()
                hf, hi = hraw[:8 * n_bf].view(np.float64), hraw[8 * n_bf:].view(np.int64)
                nan_seen, stats6 = int(hi[nh]) != 0, hi[nh + 1:nh + 7].copy()
                n_f_h = hi[:nh]
            if world > 1:
                tail = hf[nacc + 2 * nh:]
                total, nan_seen, nmax = int(tail[0]), tail[1] != 0, int(np.max(tail[2:]))
                n_f_h = hf[nacc + nh:nacc + 2 * nh].astype(np.int64)      # counts < 2^53: exact in fp64
            self._set_neval_stats(total, nmax, adaptive, reduced=True)
            if nan_seen:
                # the reference raises before touching sigf (pyx:2133-2134); the kernels have already
                # overwritten it, so put the stratification back into a consistent state first
                if self._sigf_dev is not None:
                    self._sigf_dev.fill_(1.)
                    self.sum_sigf = self._sigf_len
                raise ValueError('integrand evaluates to nan')
            acc_h = hf[:nacc]
            if not fast:
                sum_f_h, n_f_h = hf[nacc:nacc + nh].reshape(self.dim, hs), n_f_h.reshape(self.dim, hs)
            sum_sigf = float(acc_h[nf + nv])
            if pending is not None:
                book(*pending)
                pending = None
            if fast and defer:
                pending = (acc_h, self.last_neval)
            else:
                book(acc_h, self.last_neval, None if self._trace is None else (lambda mean, var: self._trace(dict(
                    itn=pitn, mean=mean, var=var, sum_sigf=sum_sigf, last_neval=self.last_neval,
                    sum_f=sum_f_h.copy(), n_f=n_f_h.copy(), flags=flags))))

            if self.beta > 0 and not self.adapt_to_errors and self.adapt:
                if sum_sigf > 0:
                    self.sum_sigf = sum_sigf
                    if plan_next:
                        _, neval_sigf, max_nh, uniform = self._plan_args()
                        self._plan_ahead = (self._plan_key(ctx, neval_sigf, max_nh, uniform), stats6)
                else:
                    # integrand appears to be a constant => even distribution of points
                    if self._sigf_dev is not None:
                        self._sigf_dev.fill_(1.)
                    self.sum_sigf = self._sigf_len
            if dev_adapt:
                self.map._adapted_on_device(ctx)
                self._ctx_map_version = (id(self.map), self.map._version)     # the context already holds this grid
            else:
                if flags & (_lib.TRAIN | _lib.TRAIN_ERRORS):
                    self.map._accumulate_training(sum_f_h, n_f_h)
                if self.alpha > 0 and self.adapt:
                    self.map.adapt(alpha=self.alpha)
            if self.analyzer is not None:
                result.update_analyzer(self.analyzer)
            if save is not None:
                result.save(save)
            if saveall is not None:
                result.saveall(self, saveall)
            if pending is None and result.converged(self.rtol, self.atol):
                break
        if pending is not None:
            book(*pending)
        return result.result

    def _iterate_unfused(self, ctx, torch, std, pitn, flags, acc, sum_f, n_f, hs, status):
        """sample -> user batch integrand -> reduce, batch by batch (pyx:2096-2197)"""
        nf = std.size
        on_device = std.on_device
        target = self.max_batch if on_device else self.min_neval_batch
        for c0, c1, rows in self._batches(ctx, target):
            x = torch.empty((rows, self.dim), dtype=torch.float64, device=ctx.device)
            wgt = torch.empty(rows, dtype=torch.float64, device=ctx.device)
            jac1d = torch.empty_like(x) if self.uses_jac else None
            # training bins travel from the sampler to the reduce kernel (2 bytes per axis) instead of
            # being re-derived there from the Philox counter
            u = self._injected_uniforms(torch, ctx, rows)
            train = flags & (_lib.TRAIN | (_lib.TRAIN_ERRORS if u is not None else 0))
            want_bins = (self.train_bins and not self.uses_jac) or u is not None    # injected uniforms cannot be replayed
            if u is not None and train and int(np.max(self.map.ninc)) > 0xffff:
                raise ValueError('ran_array_generator needs maxinc_axis <= 65535')
            bins = (torch.empty((rows, self.dim), dtype=torch.int16, device=ctx.device)
                    if train and want_bins and int(np.max(self.map.ninc)) <= 0xffff else None)
            if self._timing is not None:
                tev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                tev[0].record()
            ctx.sample(pitn, c0, c1, x, wgt, jac1d=jac1d, bins=bins, u=u)
            if self._timing is not None:
                tev[1].record()
            if on_device:
                fx = std.eval(x, jac=jac1d)
            else:
                xh = x.cpu().numpy()
                jh = jac1d.cpu().numpy() if self.uses_jac else None
                fh = np.ascontiguousarray(np.asarray(std.eval(xh, jac=jh), dtype=float).reshape(rows, nf))
                fx = torch.from_numpy(fh).to(ctx.device)
            if fx.shape != (rows, nf):
                raise ValueError('integrand returned shape %s for %d points, expected %s'
                                 % (tuple(fx.shape), rows, (rows, nf)))
            if self._timing is not None:
                tev[2].record()
            if nf <= _lib.MAX_REDUCE_NF:
                ctx.reduce(pitn, self.beta, flags, c0, c1, fx, nf, wgt, self._sigf_dev, acc, sum_f, n_f, hs, status, bins=bins)
            else:
                self._reduce_wide(ctx, torch, pitn, flags, c0, c1, fx, nf, wgt, acc, sum_f, n_f, hs, status, bins)
            if self._timing is not None:
                tev[3].record()
                self._unfused_events.append((tev, rows))
            self._launches += 3

    def _wide_passes(self, nf, device, torch):
        """plan of the reduce passes for an integrand with more components than the reduce kernel's
        instantiations (``_lib.MAX_REDUCE_NF``): [(columns, source indices, destination indices in acc)].
        Correlated integrands: components in groups of 4, one pass per pair of groups (every covariance
        block is inside some pair); uncorrelated: groups of 8, one pass each.  Means, diagonal blocks and
        sum_sigf are taken from the first pass that has them."""
        key = (nf, bool(self.correlate_integrals), str(device))
        if getattr(self, '_wide_plan', (None,))[0] == key:
            return self._wide_plan[1]
        tri = lambda s_, t_: s_ * (s_ + 1) // 2 + t_
        nv = nf * (nf + 1) // 2
        if self.correlate_integrals:
            gs = 4
            groups = [list(range(g0, min(g0 + gs, nf))) for g0 in range(0, nf, gs)]
            sets = [groups[i] + groups[j] for i in range(len(groups)) for j in range(i + 1, len(groups))]
        else:
            sets = [list(range(g0, min(g0 + _lib.MAX_REDUCE_NF, nf))) for g0 in range(0, nf, _lib.MAX_REDUCE_NF)]
        done_mean, done_var, passes = set(), set(), []
        for k, cols in enumerate(sets):
            m = len(cols)
            src, dst = [], []
            for a, s_ in enumerate(cols):
                if s_ not in done_mean:
                    done_mean.add(s_)
                    src.append(a); dst.append(s_)
                for b in range(a + 1):
                    t_ = cols[b]
                    if (s_, t_) not in done_var and (self.correlate_integrals or s_ == t_):
                        done_var.add((s_, t_))
                        src.append(m + tri(a, b)); dst.append(nf + tri(s_, t_))
            if k == 0:
                src.append(m + m * (m + 1) // 2); dst.append(nf + nv)        # sum_sigf
            passes.append((torch.tensor(cols, dtype=torch.int64, device=device), m,
                           torch.tensor(src, dtype=torch.int64, device=device),
                           torch.tensor(dst, dtype=torch.int64, device=device)))
        self._wide_plan = (key, passes)
        return passes

    def _reduce_wide(self, ctx, torch, pitn, flags, c0, c1, fx, nf, wgt, acc, sum_f, n_f, hs, status, bins):
        """per-hypercube reduce of an integrand with more than ``_lib.MAX_REDUCE_NF`` components: the reduce
        kernel runs on column subsets gathered in HBM (the reference handles any ``fcn.size`` in one loop,
        pyx:2136-2197); training and the ``sigf`` update happen in the first pass only, which holds component 0"""
        quiet = flags & ~(_lib.UPDATE_SIGF | _lib.TRAIN | _lib.TRAIN_ERRORS)
        # the pass that updates sigf runs last: the others must still see the allocation the rows were sampled with
        plan = list(enumerate(self._wide_passes(nf, ctx.device, torch)))
        for k, (cols, m, src, dst) in plan[1:] + plan[:1]:
            sub = fx.index_select(1, cols).contiguous()
            acc_p = torch.zeros(m + m * (m + 1) // 2 + 1, dtype=torch.float64, device=ctx.device)
            ctx.reduce(pitn, self.beta, flags if k == 0 else quiet, c0, c1, sub, m, wgt, self._sigf_dev, acc_p, sum_f, n_f,
                       hs, status, bins=bins)
            acc.index_add_(0, dst, acc_p.index_select(0, src))
            self._launches += 1

    @property
    def gpu_launches(self):
        """kernels launched by this integrator's context so far"""
        return self._ctx.launch_count() if self._ctx is not None else 0


_TRIL = {}


def _tril(nf):
    """indices of the lower triangle in the order the kernels accumulate it (cached: numpy builds them in ~35 us)"""
    if nf not in _TRIL:
        _TRIL[nf] = np.tril_indices(nf)
    return _TRIL[nf]


def allreduce_iteration(acc, sum_f, n_f, status):
    """The one exchange step of an iteration when the hypercube range is sharded: sum the
    per-rank partial sums [mean, var, sum_sigf], the training histogram (fp64 sums, integer
    counts) and the NaN flag over the process group (NCCL over NVLink on GPUs; gloo in the CPU
    tests).  In place; every rank ends with identical values."""
    dist = _dist()
    dist.all_reduce(acc)
    dist.all_reduce(sum_f)
    dist.all_reduce(n_f)
    dist.all_reduce(status, op=dist.ReduceOp.MAX)


def pack_iteration(buf_f, nacc, nh, n_f, status, total, nmax, rank):
    """Bring everything a sharded iteration exchanges into the one fp64 buffer that is all-reduced:
    the training counts (integers below 2^53 are exact in fp64), this rank's samples, its NaN flag, and
    its largest hypercube in slot ``rank`` of a per-rank tail (a SUM over one-hot slots is a gather,
    from which the host takes the maximum)."""
    buf_f[nacc + nh:nacc + 2 * nh].copy_(n_f.reshape(-1))
    tail = buf_f[nacc + 2 * nh:]
    tail[0] = float(total)
    tail[1:2].copy_(status[:1])
    tail[2 + rank] = float(nmax)


def exchange_iteration(buf_f):
    """The one exchange step of a sharded iteration: a single SUM all-reduce of the packed fp64
    buffer ``[mean, cov, sum_sigf | sum_f | n_f | samples, NaN count, max samples per hypercube by rank]``
    (NCCL over NVLink on GPUs; gloo in the CPU tests).  In place; every rank ends with identical values."""
    _dist().all_reduce(buf_f)


def allreduce_neval_stats(total, nmax, device):
    """(sum of samples, max samples per hypercube) over the ranks"""
    import torch
    dist = _dist()
    tot = torch.tensor([total], dtype=torch.int64, device=device)
    mx = torch.tensor([nmax], dtype=torch.int64, device=device)
    dist.all_reduce(tot)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    return int(tot.item()), int(mx.item())


def _slab_rot(ls, world):
    """rotation of the rank order in round ``ls`` of the slab deal (csrc/common.cuh: slab_rot)"""
    return (((ls * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF) >> 40) % world


def _local_cubes(nhcube, slab, rank, world):
    """global hypercube indices owned by ``rank`` in local order (host mirror of the device's
    block-cyclic ``local_to_global``)"""
    nslab = -(-nhcube // slab)
    if world == 1:
        return np.arange(nhcube)
    idx = []
    for ls in range(-(-nslab // world)):
        s = ls * world + (rank + _slab_rot(ls, world)) % world
        if s < nslab:
            idx.append(np.arange(s * slab, min((s + 1) * slab, nhcube)))
    return np.concatenate(idx) if idx else np.zeros(0, np.int64)
