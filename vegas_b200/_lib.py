"""ctypes binding of ``libvegas_b200.so`` (C ABI in ``include/vegas_b200.h``).

There is no CPU fallback: if the shared library is missing or no B200 is present, calls raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('VB200_LIB') or os.path.join(_HERE, 'libvegas_b200.so')   # VB200_LIB: developer override for A/B builds

MAXDIM = 32
CHUNK = 256
MAX_FUSED_DIM = 20      # largest padded dimension the fused kernels are instantiated for (csrc/fused_*.cu)
UPDATE_SIGF, TRAIN, TRAIN_ERRORS, CORRELATE = 1, 2, 4, 8
F_POLY, F_GAUSS_MIX, F_RIDGE, F_GENZ_OSC, F_GENZ_PRODPEAK, F_GENZ_CORNER, F_GENZ_GAUSS, F_GENZ_C0, \
    F_GENZ_DISC, F_PATHINT = range(10)

MAX_REDUCE_NF = 8            # components vb200_reduce is instantiated for (wider integrands: Integrator._reduce_wide)

# every symbol include/vegas_b200.h declares
SYMBOLS = (
    'vb200_abi_version', 'vb200_last_error', 'vb200_create', 'vb200_destroy', 'vb200_set_seed',
    'vb200_set_map', 'vb200_set_strata', 'vb200_set_integrand', 'vb200_plan', 'vb200_chunk_offsets',
    'vb200_iterate_fused', 'vb200_sample', 'vb200_reduce', 'vb200_map', 'vb200_invmap', 'vb200_jac1d',
    'vb200_add_training_data', 'vb200_map_adapt', 'vb200_uniforms', 'vb200_fp64_peak', 'vb200_launch_count',
    'vb200_last_launch', 'vb200_eval_integrand', 'vb200_dy_profile', 'vb200_sample_from_uniforms',
    'vb200_plan_ahead', 'vb200_plan_commit', 'vb200_pdf_map', 'vb200_pdf_weight',
    'vb200_map_adapt_device', 'vb200_get_map', 'vb200_iteration', 'vb200_iteration_begin', 'vb200_iteration_end',
)


class PolyParams(ctypes.Structure):
    _fields_ = [('c0', ctypes.c_double), ('c', ctypes.c_double * MAXDIM), ('p', ctypes.c_int32 * MAXDIM)]


class GaussMixParams(ctypes.Structure):
    _fields_ = [('npeak', ctypes.c_int32), ('pad', ctypes.c_int32), ('a', ctypes.c_double),
                ('norm', ctypes.c_double), ('centers_host', ctypes.c_void_p)]


class RidgeParams(ctypes.Structure):
    _fields_ = [('n', ctypes.c_int32), ('mode', ctypes.c_int32), ('a', ctypes.c_double),
                ('norm', ctypes.c_double), ('x0_host', ctypes.c_void_p)]


class GenzParams(ctypes.Structure):
    _fields_ = [('a', ctypes.c_double * MAXDIM), ('u', ctypes.c_double * MAXDIM)]


class PathIntParams(ctypes.Structure):
    _fields_ = [('T', ctypes.c_double), ('m', ctypes.c_double), ('xscale', ctypes.c_double),
                ('c2', ctypes.c_double), ('c4', ctypes.c_double), ('nx0', ctypes.c_int32),
                ('pad', ctypes.c_int32), ('x0list', ctypes.c_double * 7)]


class VegasB200Error(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VegasB200Error(
            'vegas_b200: %s is missing -- build it with `python -c "import __graft_entry__ as g; g.build()"` '
            'or `make -C vegas_b200/csrc`.  There is no CPU fallback.' % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64, f64 = (ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint32,
                                   ctypes.c_uint64, ctypes.c_double)
    pi64 = ctypes.POINTER(ctypes.c_int64)
    L.vb200_abi_version.restype = i32
    L.vb200_last_error.restype = ctypes.c_char_p
    L.vb200_create.argtypes = [ctypes.POINTER(vp), i32]
    L.vb200_destroy.argtypes = [vp]
    L.vb200_destroy.restype = None
    L.vb200_set_seed.argtypes = [vp, u64]
    L.vb200_set_map.argtypes = [vp, vp, vp, i32, i64]
    L.vb200_set_strata.argtypes = [vp, vp, i32, i64, i32, i32, pi64]
    L.vb200_set_integrand.argtypes = [vp, i32, vp, ctypes.c_size_t, ctypes.POINTER(i32)]
    L.vb200_plan.argtypes = [vp, vp, f64, i64, i64, i64, vp, pi64, vp]
    L.vb200_chunk_offsets.argtypes = [vp, vp, i64]
    L.vb200_plan_ahead.argtypes = [vp, vp, vp, f64, i64, i64, i64, vp, vp]
    L.vb200_plan_commit.argtypes = [vp, f64, pi64, pi64]
    L.vb200_iterate_fused.argtypes = [vp, u32, f64, i32, vp, vp, vp, vp, i64, vp, vp]
    L.vb200_sample.argtypes = [vp, u32, i64, i64, vp, vp, vp, vp, vp, vp, i32, vp]
    L.vb200_sample_from_uniforms.argtypes = [vp, u32, i64, i64, vp, vp, vp, vp, vp, vp, vp, i32, vp]
    L.vb200_reduce.argtypes = [vp, u32, f64, i32, i64, i64, vp, i32, vp, vp, vp, vp, vp, i64, vp, vp, vp]
    L.vb200_map.argtypes = [vp, vp, vp, vp, i64, vp]
    L.vb200_invmap.argtypes = [vp, vp, vp, vp, i64, vp]
    L.vb200_jac1d.argtypes = [vp, vp, vp, i64, vp]
    L.vb200_add_training_data.argtypes = [vp, vp, vp, i64, vp, vp, i64, vp]
    L.vb200_map_adapt.argtypes = [vp, vp, i32, i64, vp, vp, i64, f64, vp, vp, i64]
    L.vb200_uniforms.argtypes = [vp, u32, i64, i64, vp, vp]
    L.vb200_eval_integrand.argtypes = [vp, vp, i64, vp, vp]
    L.vb200_dy_profile.argtypes = [vp, u32, i64, i64, vp, i32, vp, i32, vp, vp, vp]
    L.vb200_pdf_map.argtypes = [vp, vp, i64, i32, f64, f64, i32, vp, vp, vp, vp, vp]
    L.vb200_pdf_weight.argtypes = [vp, vp, i32, vp, i64, i32, vp, vp]
    L.vb200_map_adapt_device.argtypes = [vp, vp, vp, vp, i64, f64, vp, vp]
    L.vb200_get_map.argtypes = [vp, vp, i64, vp]
    L.vb200_iteration.argtypes = [vp, u32, f64, i32, vp, vp, i64, i64, i64, i64, i64, f64, f64, i64, i64, i64, vp, vp]
    L.vb200_iteration_begin.argtypes = [vp, u32, f64, i32, vp, vp, i64, i64, i64, i64, i64, f64, f64, i64, i64, i64, vp]
    L.vb200_iteration_end.argtypes = [vp, vp, i64, vp]
    L.vb200_fp64_peak.argtypes = [i32, i32, ctypes.POINTER(f64), ctypes.POINTER(f64)]
    L.vb200_launch_count.argtypes = [vp]
    L.vb200_launch_count.restype = i64
    L.vb200_last_launch.argtypes = [vp, pi64]
    for name in SYMBOLS:
        if name not in ('vb200_last_error', 'vb200_destroy', 'vb200_launch_count'):
            getattr(L, name).restype = i32
    if L.vb200_abi_version() != 3:
        raise VegasB200Error('vegas_b200: ABI version mismatch')
    _lib = L
    return L


def check(rc):
    if rc != 0:
        err = VegasB200Error(load().vb200_last_error().decode())
        err.code = rc            # -4: no kernel compiled for this dimension / number of components
        raise err


def _ptr(t):
    """device pointer of a torch tensor (or None)"""
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    """raw handle of torch's current CUDA stream (the private fast accessor when torch has it: the public
    ``torch.cuda.current_stream()`` builds a Stream object, ~20 us a call -- more than a kernel launch)"""
    import torch
    try:
        return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch.cuda.current_device()))
    except AttributeError:
        return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def note_integrand(ctx, fid, params):
    """``ctx.integrand_serial`` counts the integrands a context has seen -- an allocation pre-pass launched ahead of
    time plans the light geometry's chunks for the integrand of its day (``Integrator._plan_key``).  Setting the
    same functor again (same id, same parameter bytes: every ``Integrator.__call__`` does) is not a new integrand."""
    sig = (int(fid), bytes(params))
    if sig != getattr(ctx, '_integrand_sig', None):
        ctx._integrand_sig = sig
        ctx.integrand_serial = getattr(ctx, 'integrand_serial', 0) + 1


class Context(object):
    """Owns one ``vb200_ctx`` on one GPU."""

    def __init__(self, device=None):
        import torch
        if not torch.cuda.is_available():
            raise VegasB200Error('vegas_b200: no CUDA device is visible; the engine runs only on a B200 '
                                 '(sm_100a) and has no CPU fallback')
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device('cuda', device if isinstance(device, int) else torch.device(device).index or 0)
        self.L = load()
        h = ctypes.c_void_p()
        check(self.L.vb200_create(ctypes.byref(h), self.device.index))
        self.h = h
        self._keep = None

    def __del__(self):
        try:
            if getattr(self, 'h', None):
                self.L.vb200_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ------------------------------------------------------------------ state
    def set_seed(self, seed):
        check(self.L.vb200_set_seed(self.h, ctypes.c_uint64(int(seed) & (2 ** 64 - 1))))

    def set_map(self, grid, ninc):
        grid = np.ascontiguousarray(grid, dtype=np.float64)
        ninc = np.ascontiguousarray(ninc, dtype=np.int64)
        check(self.L.vb200_set_map(self.h, grid.ctypes.data, ninc.ctypes.data, grid.shape[0], grid.shape[1]))

    def map_adapt_device(self, sum_f, n_f, hstride, alpha, status):
        """AdaptiveMap.adapt of the context's grid on the device from the iteration's histogram (device tensors;
        ``n_f`` int64 counts or -- after an all-reduce -- float64)"""
        import torch
        as_int = n_f.dtype == torch.int64
        check(self.L.vb200_map_adapt_device(self.h, _ptr(sum_f), _ptr(n_f) if as_int else None, None if as_int else _ptr(n_f),
                                            int(hstride), float(alpha), _ptr(status), _stream()))

    def get_map(self, shape):
        """the context's grid as a host array [dim, gstride]"""
        out = np.empty(shape, dtype=np.float64)
        check(self.L.vb200_get_map(self.h, out.ctypes.data, int(shape[1]), _stream()))
        return out

    def set_strata(self, nstrat, slab, rank=0, world=1):
        nstrat = np.ascontiguousarray(nstrat, dtype=np.int64)
        out = ctypes.c_int64()
        check(self.L.vb200_set_strata(self.h, nstrat.ctypes.data, len(nstrat), int(slab), rank, world,
                                      ctypes.byref(out)))
        return out.value

    def set_integrand(self, fid, params, keep=None):
        nf = ctypes.c_int()
        self._keep = keep
        note_integrand(self, fid, params)
        check(self.L.vb200_set_integrand(self.h, fid, ctypes.byref(params), ctypes.sizeof(params),
                                         ctypes.byref(nf)))
        return nf.value

    # ------------------------------------------------------------------ per iteration
    def plan(self, sigf, neval_sigf, min_nh, max_nh, uniform_neval, neval_hcube=None):
        stats = (ctypes.c_int64 * 4)()
        check(self.L.vb200_plan(self.h, _ptr(sigf), float(neval_sigf), int(min_nh), int(max_nh),
                                int(uniform_neval), _ptr(neval_hcube), stats, _stream()))
        return stats[0], stats[1], stats[2], stats[3]

    def plan_ahead(self, sigf, sum_sigf_dev, neval_scaled, min_nh, max_nh, uniform_neval, stats_dev):
        """launch the allocation pre-pass of the NEXT iteration (no host round trip); statistics -> stats_dev[6]"""
        check(self.L.vb200_plan_ahead(self.h, _ptr(sigf), _ptr(sum_sigf_dev), float(neval_scaled), int(min_nh), int(max_nh),
                                      int(uniform_neval), _ptr(stats_dev), _stream()))

    def plan_commit(self, neval_sigf, stats6):
        """install the statistics of the pre-pass ``plan_ahead`` launched (host copies)"""
        a = (ctypes.c_int64 * 6)(*[int(v) for v in stats6])
        out = (ctypes.c_int64 * 4)()
        check(self.L.vb200_plan_commit(self.h, float(neval_sigf), a, out))
        return out[0], out[1], out[2], out[3]

    def chunk_offsets(self, count):
        out = np.empty(count, np.int64)
        check(self.L.vb200_chunk_offsets(self.h, out.ctypes.data, count))
        return out

    def iterate_fused(self, itn, beta, flags, sigf, acc, sum_f, n_f, hstride, status):
        check(self.L.vb200_iterate_fused(self.h, itn, float(beta), flags, _ptr(sigf), _ptr(acc), _ptr(sum_f),
                                         _ptr(n_f), hstride, _ptr(status), _stream()))

    def iteration(self, itn, beta, flags, sigf, buf, nacc, nh, hstride, nf64, nwords, alpha_adapt, plan, head):
        """one fused iteration in one call (``vb200_iteration``); ``plan`` = (neval_scaled, min, max, uniform) or None;
        ``head``: host float64 array of nacc + 7 words"""
        pn, pmin, pmax, puni = plan if plan is not None else (0., 0, 0, 0)
        check(self.L.vb200_iteration(self.h, itn, float(beta), flags, _ptr(sigf), _ptr(buf), nacc, nh, hstride, nf64, nwords,
                                     float(alpha_adapt), float(pn), int(pmin), int(pmax), int(puni), head.ctypes.data, _stream()))

    def iteration_begin(self, itn, beta, flags, sigf, buf, nacc, nh, hstride, nf64, nwords, alpha_adapt, plan):
        """launch half of :meth:`iteration` (``vb200_iteration_begin``): returns with the kernels in flight"""
        pn, pmin, pmax, puni = plan if plan is not None else (0., 0, 0, 0)
        check(self.L.vb200_iteration_begin(self.h, itn, float(beta), flags, _ptr(sigf), _ptr(buf), nacc, nh, hstride, nf64,
                                           nwords, float(alpha_adapt), float(pn), int(pmin), int(pmax), int(puni), _stream()))

    def iteration_end(self, head):
        """wait for the iteration :meth:`iteration_begin` launched and fetch its head (``vb200_iteration_end``)"""
        check(self.L.vb200_iteration_end(self.h, head.ctypes.data, head.size, _stream()))

    def sample(self, itn, c0, c1, x, wgt, y=None, jac1d=None, hcube=None, transposed=False, bins=None, u=None):
        """``u``: uniforms [rows, dim] supplied by the caller (``ran_array_generator``) instead of Philox"""
        if u is not None:
            check(self.L.vb200_sample_from_uniforms(self.h, itn, c0, c1, _ptr(u), _ptr(x), _ptr(wgt), _ptr(y),
                                                    _ptr(jac1d), _ptr(hcube), _ptr(bins), int(transposed), _stream()))
            return
        check(self.L.vb200_sample(self.h, itn, c0, c1, _ptr(x), _ptr(wgt), _ptr(y), _ptr(jac1d), _ptr(hcube),
                                  _ptr(bins), int(transposed), _stream()))

    def reduce(self, itn, beta, flags, c0, c1, f, nf, wgt, sigf, acc, sum_f, n_f, hstride, status, bins=None):
        check(self.L.vb200_reduce(self.h, itn, float(beta), flags, c0, c1, _ptr(f), nf, _ptr(wgt), _ptr(sigf),
                                  _ptr(acc), _ptr(sum_f), _ptr(n_f), hstride, _ptr(bins), _ptr(status), _stream()))

    def eval_integrand(self, x, f):
        """f[rows, nf] = the context's built-in functor at x[rows, dim] (device tensors)"""
        check(self.L.vb200_eval_integrand(self.h, _ptr(x), x.shape[0], _ptr(f), _stream()))

    def pdf_map(self, theta, scale, dp_dchiv, gaussian, mean, vec_sig, p, w):
        """PDFIntegrator's change of variables: p[rows, dim], w[rows] from theta[rows, dim] (device tensors;
        mean[dim], vec_sig[dim, dim] device tensors)"""
        check(self.L.vb200_pdf_map(self.h, _ptr(theta), theta.shape[0], theta.shape[1], float(scale), float(dp_dchiv),
                                   int(bool(gaussian)), _ptr(mean), _ptr(vec_sig), _ptr(p), _ptr(w), _stream()))

    def pdf_weight(self, fp, w, out, pdf_first):
        """out[rows, nfp + 1] = [w | fp * w] (pdf_first) or [fp * w | w]; fp[rows, nfp] or None"""
        nfp = 0 if fp is None else fp.shape[1]
        check(self.L.vb200_pdf_weight(self.h, _ptr(fp), nfp, _ptr(w), w.shape[0], int(bool(pdf_first)), _ptr(out), _stream()))

    def dy_profile(self, itn, c0, c1, f, fstride, wgt, yst, acc):
        yst = np.ascontiguousarray(yst, dtype=np.float64)
        check(self.L.vb200_dy_profile(self.h, itn, c0, c1, _ptr(f), int(fstride), _ptr(wgt), len(yst) - 1,
                                      yst.ctypes.data, _ptr(acc), _stream()))

    def uniforms(self, itn, c0, c1, u):
        check(self.L.vb200_uniforms(self.h, itn, c0, c1, _ptr(u), _stream()))

    # ------------------------------------------------------------------ AdaptiveMap methods
    def map(self, y, x, jac):
        check(self.L.vb200_map(self.h, _ptr(y), _ptr(x), _ptr(jac), y.shape[0], _stream()))

    def invmap(self, x, y, jac):
        check(self.L.vb200_invmap(self.h, _ptr(x), _ptr(y), _ptr(jac), x.shape[0], _stream()))

    def jac1d(self, y, out):
        check(self.L.vb200_jac1d(self.h, _ptr(y), _ptr(out), y.shape[0], _stream()))

    def add_training_data(self, y, f, sum_f, n_f, hstride):
        check(self.L.vb200_add_training_data(self.h, _ptr(y), _ptr(f), y.shape[0], _ptr(sum_f), _ptr(n_f),
                                             hstride, _stream()))

    def launch_count(self):
        return self.L.vb200_launch_count(self.h)

    def last_launch(self):
        """geometry of the most recent engine launch"""
        out = (ctypes.c_int64 * 6)()
        check(self.L.vb200_last_launch(self.h, out))
        return dict(grid=out[0], ctas_per_sm=out[1], smem_bytes=out[2], window_bins=out[3], threads=out[4],
                    chunk_cubes=out[5])


def map_adapt(grid, ninc, sum_f, n_f, alpha, new_ninc):
    """Host step ``AdaptiveMap.adapt``: returns the new grid [dim, max(new_ninc)+1]."""
    grid = np.ascontiguousarray(grid, dtype=np.float64)
    ninc = np.ascontiguousarray(ninc, dtype=np.int64)
    new_ninc = np.ascontiguousarray(new_ninc, dtype=np.int64)
    out = np.empty((grid.shape[0], int(new_ninc.max()) + 1), np.float64)
    if sum_f is not None:
        sum_f = np.ascontiguousarray(sum_f, dtype=np.float64)
        n_f = np.ascontiguousarray(n_f, dtype=np.float64)
        sp, npp, hs = sum_f.ctypes.data, n_f.ctypes.data, sum_f.shape[1]
    else:
        sp, npp, hs = None, None, 0
    check(load().vb200_map_adapt(grid.ctypes.data, ninc.ctypes.data, grid.shape[0], grid.shape[1], sp, npp, hs,
                                 float(alpha), new_ninc.ctypes.data, out.ctypes.data, out.shape[1]))
    return out


def fp64_peak(device=0, iters=20000):
    """Measured FP64 FMA throughput (TFLOP/s) of the device -- the fused kernel's roofline."""
    t, ms = ctypes.c_double(), ctypes.c_double()
    check(load().vb200_fp64_peak(device, iters, ctypes.byref(t), ctypes.byref(ms)))
    return t.value, ms.value
