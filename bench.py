#!/usr/bin/env python
"""bench.py -- throughput of the vegas+ iteration hot path: fp64 integrand samples/sec, 8-D ridge.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one vegas+ iteration (allocate -> sample -> map -> evaluate -> per-hypercube reduce ->
train -> adapt) of the 8-D Gaussian ridge of examples/ridge.py (N=1000 Gaussians along the
diagonal), beta=0.75, alpha=0.5, neval=1e8 per GPU (weak scaling: the hypercube range is sharded
over the ranks).  W untimed adaptation iterations put the grid and sigf in steady state first.

Keys of the JSON line (rank 0):
  value        samples/s from CUDA events around each step's DEVICE work (allocation pre-pass, fused
               kernel, finalize, all-reduce), state resident in HBM; max over ranks
  e2e          samples/s through the public API ``Integrator.__call__`` for K steps, bracketed by
               barrier + synchronize: adds every step's D2H of the sums/histogram, the host
               ``AdaptiveMap.adapt`` and the H2D upload of the new grid
  roofline     fused kernel vs the FP64-FMA peak measured live by the library's DFMA probe
  cpu_baseline the reference's CPU path on this box's cores, bounded sample (rank 0, N=1)
``--impl reference`` times the unmodified reference (oracle/_ref, numpy integrand, nproc = all cores).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIM = 8
RIDGE_N = 1000
NEVAL_PER_GPU = int(1e8)
NCU_DRAM_BYTES_PER_LAUNCH = 145.98e6   # 95.15 MB read + 50.83 MB written: ncu capture of one launch at the bench configuration (profiles/prof_ridge1000_r01.summary.txt)
C_EXP = 18            # fp64 flops charged per exp(): the table-driven vb_exp_n executes 8 DFMA + 1 DADD + 1 DMUL
METRIC = 'fp64 integrand samples/sec, 8-D ridge (N=%d), vegas+ beta=0.75' % RIDGE_N


# ------------------------------------------------------------------------------------------------
# the reference CPU arm: examples/ridge.py's integrand as a numpy lbatch function (picklable so that
# the reference's multiprocessing nproc mode can ship it to its workers)
class RidgeNumpy(object):
    fcntype = 'lbatch'

    def __init__(self, dim=DIM, N=RIDGE_N):
        self.dim, self.N = dim, N
        self.x0 = np.linspace(0.4, 0.6, N)
        self.norm = (100. / np.pi) ** (dim / 2.)

    def __call__(self, x):
        out = np.empty(x.shape[0])
        for i in range(0, x.shape[0], 2048):
            xb = x[i:i + 2048]
            dx2 = np.zeros((xb.shape[0], self.N))
            for d in range(x.shape[1]):
                dx2 += (xb[:, d, None] - self.x0[None, :]) ** 2
            out[i:i + 2048] = np.average(np.exp(-100. * dx2), axis=1) * self.norm
        return out


def reference_arm(steps, warmup, neval=None, nproc=None):
    """the reference's own implementation (oracle/_ref = unmodified _vegas.pyx compiled here) with
    its multiprocessing nproc mode on all host cores; falls back to the C restatement
    (oracle port, 1 core) if the compiled reference did not travel"""
    cores = nproc or os.cpu_count() or 1
    neval = int(neval or 1e6)
    f = RidgeNumpy()
    kind = 'reference'
    try:
        sys.path.insert(0, os.path.join(ROOT, 'oracle', 'gvar_shim'))
        sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
        import vegas as ref
        integ = ref.Integrator(DIM * [[0., 1.]], nproc=cores, sync_ran=False)
        integ(f, nitn=max(warmup, 1), neval=neval)
        t0 = time.perf_counter()
        r = integ(f, nitn=steps, neval=neval)
        dt = time.perf_counter() - t0
        nsamp = float(r.sum_neval)
        used = integ.nproc
    except ImportError:
        kind, used = 'port', 1
        from oracle import oracle as O
        v = O.Vegas(DIM * [[0., 1.]], neval=neval)
        rng = np.random.default_rng(1)
        gen = lambda h0, nh: rng.random((int(nh.sum()), DIM))
        for _ in range(max(warmup, 1)):
            v.iterate(lambda x: f(x), gen)
            v.adapt_map()
        nsamp, t0 = 0.0, time.perf_counter()
        for _ in range(steps):
            v.iterate(lambda x: f(x), gen)
            v.adapt_map()
            nsamp += v.last_neval
        dt = time.perf_counter() - t0
    return dict(value=nsamp / dt, unit='samples/s', cores=int(used), kind=kind,
                sample='%d iterations of neval=%d (8-D ridge N=%d, numpy lbatch integrand)' % (steps, neval, RIDGE_N),
                seconds=dt, steps=steps)


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([v.strip() for v in out.strip().split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith('active')})
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


class Env(object):
    """rank / device / collectives of this bench process"""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get('RANK', 0))
        self.world = int(os.environ.get('WORLD_SIZE', 1))
        self.local = int(os.environ.get('LOCAL_RANK', 0))
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group('nccl', device_id=self.dev)
            self.dist = dist

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, v):
        if self.dist is None:
            return float(v)
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t)
        return float(t.item())


def timed_iterations(env, integ, f, steps):
    """K iterations of an adapted integrator, bracketed by barrier + synchronize.  Returns the result,
    the whole job's device time of the steps (max over ranks of the summed per-step event times),
    the same for the engine kernel alone, and this rank's samples."""
    integ._timing = []
    env.barrier()
    r = integ(f, nitn=steps)
    env.barrier()
    step_ms = float(np.sum([ev[0].elapsed_time(ev[3]) for ev, _ in integ._timing]))
    kern_ms = float(np.sum([ev[1].elapsed_time(ev[2]) for ev, _ in integ._timing]))
    local = float(np.sum([tot for _, tot in integ._timing]))
    integ._timing = None
    return r, env.max_over_ranks(step_ms), kern_ms, local


def kernel_name(launch, functor):
    light = launch.get('threads') == 256
    return 'k_engine<FusedSrc<%s, D, %s>> (%d-thread CTAs, %d per SM, %d-cube chunks)' % (
        functor + ('Light' if light and functor == 'FRidge' else ''), 'light geometry' if light else 'heavy geometry',
        launch.get('threads', 0), launch.get('ctas_per_sm', 0), launch.get('chunk_cubes', 0))


def ours(args):
    import vegas_b200 as vegas
    from vegas_b200 import _lib
    env = Env()
    torch, rank, world = env.torch, env.rank, env.world

    neval = NEVAL_PER_GPU * world
    f = vegas.integrands.Ridge(DIM, N=args.ridge_n)
    integ = vegas.Integrator(DIM * [[0., 1.]], neval=neval, mpi=world > 1, seed=0x5eed + 1, max_mem=1e11)
    fp64_peak, _ = _lib.fp64_peak(env.local, 20000)
    integ(f, nitn=max(args.warmup, 3))                      # untimed adaptation (>= 3 warm-up steps)

    # ---- value: device time of each step's GPU work
    sampler = ClockSampler(env.local)
    if rank == 0:
        sampler.start()
    l0 = integ.gpu_launches
    t0 = time.perf_counter()
    res_v, dev_ms, kern_ms, local_samples = timed_iterations(env, integ, f, args.steps)
    wall_v = time.perf_counter() - t0
    launches = integ.gpu_launches - l0
    samples = float(res_v.sum_neval)
    value = samples / (dev_ms * 1e-3)

    # ---- e2e: the public API call, K steps, host epilogue and copies included
    env.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res_e = integ(f, nitn=args.steps)
    e1.record()
    env.barrier()
    e2e_ms = env.max_over_ranks(e0.elapsed_time(e1))
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    grid_bytes = integ.map.grid.size * 8 + integ.map.ninc.size * 8
    hs = integ.map.inc.shape[1]
    d2h = (3 + 2 * DIM * hs + 2 + world) * 8 + (DIM * hs + 1 + 6) * 8     # the iteration's packed fp64 + int64 buffers
    e2e = dict(value=float(res_e.sum_neval) / (e2e_ms * 1e-3), unit='samples/s', ms_per_step=e2e_ms / args.steps,
               h2d_bytes_per_step=int(grid_bytes), d2h_bytes_per_step=int(d2h),
               result='%s Q=%.2f' % (res_e, res_e.Q))

    # ---- roofline of the fused kernel (this rank's launches)
    flops_per_sample = (9 * DIM + 10) + f.flops_per_sample(C_EXP)
    ach = local_samples * flops_per_sample / (kern_ms * 1e-3) / 1e12
    geom = integ._ctx.last_launch()
    nlocal = int(integ._nlocal)
    roofline = dict(bound='fp64', achieved=ach, peak=fp64_peak, unit='TFLOP/s', frac=ach / fp64_peak,
                    traffic=NCU_DRAM_BYTES_PER_LAUNCH if (args.ridge_n == RIDGE_N and world == 1) else None,
                    traffic_model_bytes=16 * nlocal,
                    traffic_source='ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch at the 1-GPU '
                                   'configuration (profiles/prof_ridge1000_r02.summary.txt); the kernel is FP64-bound, its HBM '
                                   'traffic is the sigf stream, 8 B read + 8 B written per hypercube of this rank '
                                   '(traffic_model_bytes; part of the write-back is still in L2 when the kernel ends)',
                    kernel=kernel_name(geom, 'FRidge'), kernel_ms=kern_ms / args.steps,
                    launch=geom, flops_per_sample=flops_per_sample,
                    flops_note='9*D+10 engine + N*(3*D+2+C_exp) integrand, C_exp=%d (8 DFMA + 1 DADD + 1 DMUL executed by '
                               'the table-driven exp); a ridge term is 28 FP64 instructions for %d flops, so 100%% FP64-pipe '
                               'occupancy corresponds to frac=%.3f' % (C_EXP, 3 * DIM + 2 + C_EXP, (3 * DIM + 2 + C_EXP) / 56.),
                    peak_source='measured live: vb200_fp64_peak DFMA probe (MEASURED_PEAKS.json has no FP64 entry); '
                                'nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2')
    out = dict(metric=METRIC.replace('N=%d' % RIDGE_N, 'N=%d' % args.ridge_n), value=value, unit='samples/s', n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
               ms_per_step=dev_ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f64',
               data='synthetic',
               config=dict(workload='8-D Gaussian ridge N=%d (examples/ridge.py), vegas+ beta=0.75 alpha=0.5, '
                                    'neval=%.0e per GPU per iteration' % (args.ridge_n, NEVAL_PER_GPU),
                           neval=neval, nstrat=[int(v) for v in integ.nstrat], nhcube=int(integ.nhcube),
                           parallelism='hypercube range sharded block-cyclically over %d GPU(s), one all-reduce per iteration' % world,
                           cache='no inputs are re-read: samples are generated in registers; sigf (%.0f MB) > L2 is streamed once'
                                 % (integ.nhcube * 8 / 1e6)),
               e2e=e2e, roofline=roofline, gpu_launches=int(launches), clocks=None,
               result='%s Q=%.2f' % (res_v, res_v.Q), wall_s=wall_v)
    if rank == 0:
        out['clocks'] = sampler.summary()
        if world == 1 and not args.no_cpu:
            # separate process: the reference forks a multiprocessing pool, which must not inherit CUDA
            # bounded sample: 1 + 3 iterations of neval=1e6 (~10-15 s on 16 cores; the GPU arm runs 1e8)
            cp = subprocess.run([sys.executable, os.path.abspath(__file__), '--impl', 'reference', '--steps', '3',
                                 '--warmup', '1', '--cpu-neval', '1000000'], capture_output=True, text=True)
            try:
                out['cpu_baseline'] = json.loads(cp.stdout.strip().splitlines()[-1])['cpu_baseline']
            except Exception:
                out['cpu_baseline'] = dict(error=(cp.stderr or cp.stdout)[-300:])
    if args.variants:
        # companion numbers at every world size (all ranks take part: the integrators are sharded);
        # never allowed to take the headline line down with them
        for key, fn in (('variants', variants), ('other_configs', other_configs), ('shard_parity', shard_parity)):
            if key == 'shard_parity' and world == 1:
                continue
            try:
                out[key] = fn(env, vegas, _lib, fp64_peak)
            except Exception as e:        # noqa: BLE001
                out[key] = dict(error='%s: %s' % (type(e).__name__, str(e)[:300]))
    if rank == 0:
        print(json.dumps(out))
    if env.dist is not None:
        env.dist.destroy_process_group()


def variants(env, vegas, _lib, fp64_peak):
    """The north-star line: the same 8-D vegas+ workload with short ridges, where the sampler around
    the integrand matters -- N=1 (a single Gaussian, ~130 flops per sample) and N=30 (the longest
    ridge for which 1e11 samples/s on 8 GPUs is compatible with 50 % of the FP64 peak) --
    and the N=1000 ridge through the shifted-mean identity.  Weak scaling like the headline
    (neval = 1e8 per GPU); value = whole-job samples/s from the device time of the steps, max over
    ranks; roofline_frac = this rank's algorithmic flops / its kernel time / measured FP64 peak."""
    world = env.world
    out = {}
    for n, shifted in ((1, False), (30, False), (72, False), (RIDGE_N, True)):
        f = vegas.integrands.Ridge(DIM, N=n, lo=0.5 if n == 1 else 0.4, hi=0.5 if n == 1 else 0.6, shifted=shifted)
        integ = vegas.Integrator(DIM * [[0., 1.]], neval=NEVAL_PER_GPU * world, seed=77, mpi=world > 1, max_mem=1e11)
        integ(f, nitn=5)
        r, ms, kms, local = timed_iterations(env, integ, f, 5)
        env.barrier()
        e0, e1 = env.torch.cuda.Event(enable_timing=True), env.torch.cuda.Event(enable_timing=True)
        e0.record()
        re = integ(f, nitn=5)
        e1.record()
        env.barrier()
        e2e_ms = env.max_over_ranks(e0.elapsed_time(e1))
        fl = (9 * DIM + 10) + f.flops_per_sample(C_EXP)
        geom = integ._ctx.last_launch()
        out['ridge_N%d%s' % (n, '_shifted' if shifted else '')] = dict(
            value=float(r.sum_neval) / (ms * 1e-3), unit='samples/s', n_gpus=world, ms_per_step=ms / 5,
            e2e=float(re.sum_neval) / (e2e_ms * 1e-3), kernel_ms=kms / 5,
            roofline_frac=local * fl / (kms * 1e-3) / 1e12 / fp64_peak, flops_per_sample=fl, result=str(r),
            kernel=kernel_name(geom, 'FRidge'), launch=geom)
    return out


def hbm_peak_gbs():
    """HBM copy bandwidth the driver measured on this pool (MEASURED_PEAKS.json), else the figure it
    recorded when the survey was written"""
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['hbm_gbs']), 'MEASURED_PEAKS.json'
    except Exception:
        return 6549.8, 'BASELINE.md section 4 (MEASURED_PEAKS.json absent)'


def other_configs(env, vegas, _lib, fp64_peak):
    """The other BASELINE.json configurations.  On N > 1 GPUs these are STRONG scaling: neval is the
    configuration's own (fixed), the hypercube range is sharded over the ranks; value = samples of the
    whole job / device time of the steps (max over ranks).  Config 2 at fixed neval=1e8 is included so
    that the headline workload has a strong-scaling curve as well.  One GPU adds the HBM-bound callback
    path of config 4 (kernel-level and end to end through ``Integrator.__call__``)."""
    torch, world = env.torch, env.world
    F = vegas.integrands
    rng = np.random.default_rng(0x5eed + 3)
    out = {}
    cfgs = {
        'cfg1_gauss4_neval1e4': (F.GaussMix([4 * [0.5]], 100., 1013.2118364296088), [[-1., 1.]] + 3 * [[0., 1.]], dict(neval=1e4), 10),
        'cfg2_ridge8_N1000_neval1e8_fixed': (F.Ridge(DIM, N=RIDGE_N), DIM * [[0., 1.]], dict(neval=1e8), 3),
        'cfg3_genz10_product_peak_neval1e9': (F.Genz('product_peak', 2 + 3 * rng.random(10), rng.random(10)), 10 * [[0., 1.]],
                                              dict(neval=1e9, max_mem=1e10), 3),
        'cfg3_genz10_oscillatory_neval1e9': (F.Genz('oscillatory', rng.random(10), rng.random(10)), 10 * [[0., 1.]],
                                             dict(neval=1e9, max_mem=1e10), 3),
        'cfg4_pathint10_fused_neval1e8': (F.PathIntegral(T=4., ndT=10, x0list=np.linspace(0, 2., 6)), 10 * [[-np.pi / 2, np.pi / 2]],
                                          dict(neval=1e8, alpha=0.1), 3),
        'cfg5_peaks20_nstrat60x5_neval1e10': (F.GaussMix([5 * [c] + 15 * [0.45] for c in (.23, .39, .74)], 100., 356047712484621.56),
                                              20 * [[0., 1.]], dict(neval=1e10, nstrat=5 * [60] + 15 * [1], max_mem=1e11), 3),
    }
    integ = None
    for name, (f, limits, kw, steps) in cfgs.items():
        if name.startswith('cfg1') and world > 1:
            continue                                    # 1080 hypercubes: nothing to shard
        del integ
        torch.cuda.empty_cache()
        integ = vegas.Integrator(limits, seed=5, mpi=world > 1, **kw)
        integ(f, nitn=5)
        r, ms, kms, local = timed_iterations(env, integ, f, steps)
        r0 = r if not hasattr(r, 'keys') else r['exp(-E0*T)']
        out[name] = dict(value=float(r.sum_neval) / (ms * 1e-3), unit='samples/s', scaling='strong' if world > 1 else None,
                         n_gpus=world, neval=kw['neval'], ms_per_step=ms / steps, kernel_ms=kms / steps,
                         nhcube=int(integ.nhcube), neval_hcube_range=[int(v) for v in integ.neval_hcube_range],
                         result='%s Q=%.2f' % (r0, r.Q), launch=integ._ctx.last_launch())
        if name.startswith('cfg1'):
            # the everyday size: wall time of the public call, host epilogue (D2H, adapt, result update) included
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            integ(f, nitn=20)
            torch.cuda.synchronize()
            out[name]['e2e_ms_per_step'] = (time.perf_counter() - t0) / 20 * 1e3
        if name.startswith('cfg4'):
            integ4, f4 = integ, f
            integ = None
    if world == 1:
        try:
            out['cfg4_pathint10_callback_path'] = callback_path(env, vegas, integ4, f4)
        except Exception as e:        # noqa: BLE001
            out['cfg4_pathint10_callback_path'] = dict(error='%s: %s' % (type(e).__name__, str(e)[:300]))
        try:
            out['pdf6_expectation_values'] = pdf_path(env, vegas)
        except Exception as e:        # noqa: BLE001
            out['pdf6_expectation_values'] = dict(error='%s: %s' % (type(e).__name__, str(e)[:300]))
        try:
            out['cfg1_reference_cpu'] = cfg1_reference()
        except Exception as e:        # noqa: BLE001
            out['cfg1_reference_cpu'] = dict(error='%s: %s' % (type(e).__name__, str(e)[:300]))
    return out


def pdf_path(env, vegas):
    """PDFIntegrator (reference src/vegas/__init__.py:373-1188) end to end: expectation values of three functions of
    six correlated Gaussian parameters, f(p) a @devicebatchintegrand, neval=1e7: sampler -> k_pdf_map -> f(p) in HBM ->
    k_pdf_weight -> reduce (4 components with covariances), through PDFIntegrator.__call__"""
    torch = env.torch
    from vegas_b200._gv import gv
    rng = np.random.default_rng(11)
    a = rng.normal(size=(6, 6))
    cov = a @ a.T + 0.5 * np.eye(6)
    mean = rng.normal(size=6)
    integ = vegas.PDFIntegrator(gv.gvar(mean, cov), neval=1e7, seed=12)
    integ(nitn=5)
    f = vegas.devicebatchintegrand(lambda p: torch.stack([p[:, 0], p[:, 0] * p[:, 1], p[:, 2] ** 2], dim=1))
    integ(f, nitn=1, adapt=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = integ(f, nitn=5, adapt=False)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    exact = [mean[0], cov[0, 1] + mean[0] * mean[1], cov[2, 2] + mean[2] ** 2]
    pulls = [float((r[i].mean - exact[i]) / r[i].sdev) for i in range(3)]
    return dict(value=float(r.sum_neval) / sec, unit='samples/s', ms_per_step=sec / 5 * 1e3, neval=1e7, dim=6, components=4,
                result=str(np.asarray(r)), pdfnorm=str(r.pdfnorm), pulls_vs_exact=pulls, Q=float(r.Q))


def cfg1_reference():
    """the reference's own CPU time for config 1 (examples/simple.py-style: 4-D Gaussian, nitn=10, neval=1e4)"""
    code = ("import sys,time,numpy as np\n"
            "sys.path.insert(0,%r); sys.path.insert(0,%r)\n"
            "import vegas\n"
            "f=vegas.lbatchintegrand(lambda x: np.exp(-100.*np.sum((x-0.5)**2,axis=1))*1013.2118364296088)\n"
            "integ=vegas.Integrator([[-1.,1.]]+3*[[0.,1.]])\n"
            "integ(f,nitn=10,neval=1e4)\n"
            "t0=time.perf_counter(); r=integ(f,nitn=10,neval=1e4); dt=time.perf_counter()-t0\n"
            "print(dt/10*1e3, r.sum_neval/dt)\n") % (os.path.join(ROOT, 'oracle', 'gvar_shim'), os.path.join(ROOT, 'oracle', '_ref'))
    cp = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=120)
    ms, sps = [float(v) for v in cp.stdout.split()[-2:]]
    return dict(ms_per_step=ms, value=sps, unit='samples/s', cores=1, kind='reference',
                note='unmodified reference (oracle/_ref), numpy lbatch integrand, 10 iterations of neval=1e4 after 10 of adaptation')


def callback_path(env, vegas, integ, f):
    """config 4 through the HBM-bound callback path on one GPU: kernel-level rows/s of the sampler, the
    functor-on-buffers kernel and the reduce kernel on one 8M-row batch of the adapted state, against the
    HBM copy bandwidth; and end to end through ``Integrator.__call__`` with a ``@devicebatchintegrand``."""
    torch = env.torch
    peak, src = hbm_peak_gbs()
    ctx, _ = integ._engine()
    integ._plan(ctx)
    batches = integ._batches(ctx, 1 << 23)
    c0, c1, rows = batches[len(batches) // 2]
    dim, nf, dev = integ.dim, 7, ctx.device
    x = torch.empty((rows, dim), dtype=torch.float64, device=dev)
    wgt = torch.empty(rows, dtype=torch.float64, device=dev)
    bins = torch.empty((rows, dim), dtype=torch.int16, device=dev)
    fx = torch.empty((rows, nf), dtype=torch.float64, device=dev)
    hs = int(integ.map.inc.shape[1])
    acc = torch.zeros(nf + nf * (nf + 1) // 2 + 1, dtype=torch.float64, device=dev)
    sum_f = torch.zeros((dim, hs), dtype=torch.float64, device=dev)
    n_f = torch.zeros((dim, hs), dtype=torch.int64, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    sig = integ._sigf_dev.clone()
    flags = integ._flags(nf)

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e-3

    ts = timeit(lambda: ctx.sample(7, c0, c1, x, wgt, bins=bins))
    te = timeit(lambda: ctx.eval_integrand(x, fx))
    tr = timeit(lambda: ctx.reduce(7, integ.beta, flags, c0, c1, fx, nf, wgt, sig, acc, sum_f, n_f, hs, status, bins=bins))
    b_s, b_e, b_r = 8 * dim + 8 + 2 * dim, 8 * dim + 8 * nf, 8 * nf + 8 + 2 * dim
    out = dict(
        rows=int(rows), note='kernel-level, one batch of the adapted state (vegas+ allocation %s samples per hypercube); '
                             'inputs > L2 (x alone is %.0f MB); bytes per row: sampler writes x, wgt, training bins; the '
                             'integrand kernel (library functor on buffers) reads x, writes f; reduce reads f, wgt, bins'
                             % (list(integ.neval_hcube_range), rows * 8 * dim / 1e6),
        sampler=dict(rows_per_s=rows / ts, bytes_per_row=b_s, gbs=rows * b_s / ts / 1e9, frac=rows * b_s / ts / 1e9 / peak),
        integrand=dict(rows_per_s=rows / te, bytes_per_row=b_e, gbs=rows * b_e / te / 1e9, frac=rows * b_e / te / 1e9 / peak),
        reduce=dict(rows_per_s=rows / tr, bytes_per_row=b_r, gbs=rows * b_r / tr / 1e9, frac=rows * b_r / tr / 1e9 / peak),
        path=dict(value=rows / (ts + te + tr), unit='samples/s', bytes_per_sample=b_s + b_e + b_r,
                  gbs=rows * (b_s + b_e + b_r) / (ts + te + tr) / 1e9, frac=rows * (b_s + b_e + b_r) / (ts + te + tr) / 1e9 / peak),
        hbm_peak_gbs=peak, hbm_peak_source=src)
    del x, wgt, bins, fx
    # ---- end to end: Integrator.__call__ with the functor as a device batch callback (DLPack fp64 buffers in HBM)
    fdev = f.device_twin(dim)
    i2 = vegas.Integrator(integ, seed=6)
    i2(fdev, nitn=1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = i2(fdev, nitn=3)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3
    bps = b_s + b_e + b_r
    out['e2e'] = dict(value=float(r.sum_neval) / sec, unit='samples/s', ms_per_step=sec / 3 * 1e3, bytes_per_sample=bps,
                      gbs=float(r.sum_neval) * bps / sec / 1e9, frac=float(r.sum_neval) * bps / sec / 1e9 / peak,
                      result='%s Q=%.2f' % (r[0], r.Q),
                      note='Integrator.__call__(devicebatchintegrand), 3 iterations of neval=1e8 in batches of max_batch rows: '
                           'sample -> callback (library functor on the HBM buffers) -> reduce, plus the host epilogue')
    return out


def shard_parity(env, vegas, _lib, fp64_peak):
    """N GPUs == 1 GPU, checked inside the bench run (the driver's test box has one GPU): the same small
    vegas+ problem integrated with the hypercube range sharded over the ranks and, on every rank,
    unsharded.  The Philox stream is a function of (seed, iteration, hypercube, sample), so the first
    iteration sees identical samples: its training counts n_f and last_neval must be EQUAL, its
    sums agree to rounding; later iterations start from grids that differ in the last bits."""
    f = vegas.integrands.Ridge(4, N=20)
    rows = {}
    for mode in ('sharded', 'single'):
        tr = []
        integ = vegas.Integrator(4 * [[0., 1.]], neval=4e5, seed=4242, mpi=(mode == 'sharded'), slab=256)
        integ._trace = tr.append
        r = integ(f, nitn=3)
        rows[mode] = (tr, r)
    a, b = rows['sharded'][0], rows['single'][0]
    rel = lambda u, v: float(np.max(np.abs(np.asarray(u) - np.asarray(v)) / np.maximum(np.abs(np.asarray(v)), 1e-300)))
    ok = dict(n_gpus=env.world,
              n_f_equal=bool(np.array_equal(a[0]['n_f'], b[0]['n_f'])),
              last_neval_equal=bool(a[0]['last_neval'] == b[0]['last_neval']),
              mean_rel=rel(a[0]['mean'], b[0]['mean']), var_rel=rel(a[0]['var'], b[0]['var']),
              sum_sigf_rel=rel(a[0]['sum_sigf'], b[0]['sum_sigf']), sum_f_rel=rel(a[0]['sum_f'].sum(), b[0]['sum_f'].sum()),
              mean_rel_all_iterations=max(rel(x['mean'], y['mean']) for x, y in zip(a, b)),
              last_neval_all=[[int(x['last_neval']), int(y['last_neval'])] for x, y in zip(a, b)],
              results=[str(rows['sharded'][1]), str(rows['single'][1])])
    # every rank must have seen the same comparison
    ok['all_ranks_agree'] = bool(env.sum_over_ranks(1.0 if (ok['n_f_equal'] and ok['last_neval_equal']) else 0.0) == env.world)
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours')
    ap.add_argument('--ridge-n', type=int, default=RIDGE_N)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cpu-neval', type=int, default=1000000)
    ap.add_argument('--variants', action='store_true', default=True)
    ap.add_argument('--no-variants', dest='variants', action='store_false')
    args = ap.parse_args()
    if args.impl == 'reference':
        if int(os.environ.get('RANK', 0)) != 0:
            return
        r = reference_arm(steps=args.steps, warmup=args.warmup, neval=args.cpu_neval)
        print(json.dumps(dict(
            impl='reference', metric=METRIC, value=r['value'], unit='samples/s', n_gpus=args.gpus, steps=args.steps,
            warmup=args.warmup, ms_per_step=r['seconds'] / args.steps * 1e3, higher_is_better=True, scaling='weak',
            vs_baseline=None, dtype='f64', data='synthetic',
            config=dict(workload='8-D Gaussian ridge N=%d (examples/ridge.py), vegas+ beta=0.75 alpha=0.5; CPU '
                                 'sample: neval=%d per iteration' % (RIDGE_N, args.cpu_neval)),
            cpu_baseline=r, e2e=dict(value=r['value'], unit='samples/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0),
            gpu_launches=0)))
        return
    ours(args)


if __name__ == '__main__':
    main()
