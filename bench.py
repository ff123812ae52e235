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


def ours(args):
    import torch
    import vegas_b200 as vegas
    from vegas_b200 import _lib
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    neval = NEVAL_PER_GPU * world
    f = vegas.integrands.Ridge(DIM, N=args.ridge_n)
    integ = vegas.Integrator(DIM * [[0., 1.]], neval=neval, mpi=world > 1, seed=0x5eed + 1, max_mem=1e11)
    fp64_peak, _ = _lib.fp64_peak(local, 20000)
    integ(f, nitn=max(args.warmup, 3))                      # untimed adaptation (>= 3 warm-up steps)

    # ---- value: device time of each step's GPU work
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    integ._timing = []
    l0 = integ.gpu_launches
    barrier()
    t0 = time.perf_counter()
    res_v = integ(f, nitn=args.steps)
    barrier()
    wall_v = time.perf_counter() - t0
    launches = integ.gpu_launches - l0
    step_ms = [ev[0].elapsed_time(ev[3]) for ev, _ in integ._timing]
    kern_ms = [ev[1].elapsed_time(ev[2]) for ev, _ in integ._timing]
    local_samples = [tot for _, tot in integ._timing]
    integ._timing = None
    dev_ms = max_over_ranks(float(np.sum(step_ms)))
    samples = float(res_v.sum_neval)
    value = samples / (dev_ms * 1e-3)

    # ---- e2e: the public API call, K steps, host epilogue and copies included
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res_e = integ(f, nitn=args.steps)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    if rank == 0:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    grid_bytes = integ.map.grid.size * 8 + integ.map.ninc.size * 8
    hs = integ.map.inc.shape[1]
    d2h = (1 + 1 + 1) * 8 + 2 * DIM * hs * 8 + 4 + 4 * 8      # acc, sum_f + n_f, status, plan stats
    e2e = dict(value=float(res_e.sum_neval) / (e2e_ms * 1e-3), unit='samples/s', ms_per_step=e2e_ms / args.steps,
               h2d_bytes_per_step=int(grid_bytes + 4 * 8), d2h_bytes_per_step=int(d2h),
               result='%s Q=%.2f' % (res_e, res_e.Q))

    # ---- roofline of the fused kernel (this rank's launches)
    flops_per_sample = (9 * DIM + 10) + f.flops_per_sample(C_EXP)
    ach = float(np.sum(local_samples)) * flops_per_sample / (float(np.sum(kern_ms)) * 1e-3) / 1e12
    geom = integ._ctx.last_launch()
    roofline = dict(bound='fp64', achieved=ach, peak=fp64_peak, unit='TFLOP/s', frac=ach / fp64_peak,
                    traffic=NCU_DRAM_BYTES_PER_LAUNCH if args.ridge_n == RIDGE_N else None,
                    traffic_source='ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch at this '
                                   'configuration (profiles/prof_ridge1000_r01.summary.txt); the kernel is FP64-bound, HBM '
                                   'traffic is the sigf stream (8 B read + 8 B written per hypercube)',
                    kernel='k_engine<FusedSrc<FRidge,8,false>>', kernel_ms=float(np.mean(kern_ms)),
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
                           parallelism='hypercube range sharded block-cyclically over %d GPU(s)' % world,
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
        if world == 1 and args.variants:
            # companion numbers: never allowed to take the headline line down with them
            for key, fn in (('variants', variants), ('other_configs', other_configs)):
                try:
                    out[key] = fn(vegas, _lib, fp64_peak)
                except Exception as e:        # noqa: BLE001
                    out[key] = dict(error='%s: %s' % (type(e).__name__, str(e)[:300]))
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def variants(vegas, _lib, fp64_peak):
    """engine-bound companion numbers on one GPU: the same 8-D workload with a single Gaussian
    (N=1) -- the integrand is then ~100 flops and the sampler itself is what is timed"""
    import torch
    out = {}
    for n, shifted in ((1, False), (30, False), (RIDGE_N, True)):
        f = vegas.integrands.Ridge(DIM, N=n, lo=0.5 if n == 1 else 0.4, hi=0.5 if n == 1 else 0.6, shifted=shifted)
        integ = vegas.Integrator(DIM * [[0., 1.]], neval=NEVAL_PER_GPU, seed=77)
        integ(f, nitn=5)
        integ._timing = []
        r = integ(f, nitn=5)
        torch.cuda.synchronize()
        ms = float(np.sum([ev[0].elapsed_time(ev[3]) for ev, _ in integ._timing]))
        kms = float(np.sum([ev[1].elapsed_time(ev[2]) for ev, _ in integ._timing]))
        fl = (9 * DIM + 10) + f.flops_per_sample(C_EXP)
        out['ridge_N%d%s' % (n, '_shifted' if shifted else '')] = dict(value=float(r.sum_neval) / (ms * 1e-3), unit='samples/s',
                                    roofline_frac=float(r.sum_neval) * fl / (kms * 1e-3) / 1e12 / fp64_peak,
                                    flops_per_sample=fl, result=str(r), launch=integ._ctx.last_launch())
    return out


def hbm_peak_gbs():
    """HBM copy bandwidth the driver measured on this pool (MEASURED_PEAKS.json), else the figure it
    recorded when the survey was written"""
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as fh:
            return float(json.load(fh)['hbm_gbs']), 'MEASURED_PEAKS.json'
    except Exception:
        return 6549.8, 'BASELINE.md section 4 (MEASURED_PEAKS.json absent)'


def other_configs(vegas, _lib, fp64_peak):
    """one-GPU numbers of the other BASELINE.json configurations (fused kernels) and of the HBM-bound
    callback path (config 4: 10-D path integral, 7 outputs): kernel-level rows/s of the sampler and
    the reduce kernel on one 8M-row batch, against the HBM copy bandwidth"""
    import torch
    F = vegas.integrands
    rng = np.random.default_rng(0x5eed + 3)
    out = {}
    cfgs = {
        'cfg3_genz10_product_peak': (F.Genz('product_peak', 2 + 3 * rng.random(10), rng.random(10)), 10 * [[0., 1.]],
                                     dict(neval=1e9, max_mem=1e10)),
        'cfg5_peaks20_nstrat30x5': (F.GaussMix([5 * [c] + 15 * [0.45] for c in (.23, .39, .74)], 100., 356047712484621.56),
                                    20 * [[0., 1.]], dict(neval=5e8, nstrat=5 * [30] + 15 * [1], max_mem=1e10)),
        'cfg4_pathint10_fused': (F.PathIntegral(T=4., ndT=10, x0list=np.linspace(0, 2., 6)), 10 * [[-np.pi / 2, np.pi / 2]],
                                 dict(neval=1e8, alpha=0.1)),
    }
    for name, (f, limits, kw) in cfgs.items():
        integ = vegas.Integrator(limits, seed=5, **kw)
        integ(f, nitn=5)
        integ._timing = []
        r = integ(f, nitn=3)
        torch.cuda.synchronize()
        kms = float(np.sum([ev[1].elapsed_time(ev[2]) for ev, _ in integ._timing]))
        r0 = r if not hasattr(r, 'keys') else r['exp(-E0*T)']
        out[name] = dict(value=float(r.sum_neval) / (kms * 1e-3), unit='samples/s', neval=kw['neval'], kernel_ms=kms / 3,
                         nhcube=int(integ.nhcube), neval_hcube_range=[int(v) for v in integ.neval_hcube_range],
                         result='%s Q=%.2f' % (r0, r.Q), launch=integ._ctx.last_launch())
    # ---- callback path kernels on the adapted path-integral state (integ, f from the last loop turn)
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
    out['cfg4_pathint10_callback_path'] = dict(
        rows=int(rows), note='kernel-level, one batch of the adapted state (vegas+ allocation 2..50000 samples per hypercube); '
                             'inputs > L2 (x alone is %.0f MB); bytes per row: sampler writes x, wgt, training bins; the '
                             'integrand kernel (library functor on buffers) reads x, writes f; reduce reads f, wgt, bins' % (rows * 8 * dim / 1e6),
        sampler=dict(rows_per_s=rows / ts, bytes_per_row=b_s, gbs=rows * b_s / ts / 1e9, frac=rows * b_s / ts / 1e9 / peak),
        integrand=dict(rows_per_s=rows / te, bytes_per_row=b_e, gbs=rows * b_e / te / 1e9, frac=rows * b_e / te / 1e9 / peak),
        reduce=dict(rows_per_s=rows / tr, bytes_per_row=b_r, gbs=rows * b_r / tr / 1e9, frac=rows * b_r / tr / 1e9 / peak),
        path=dict(value=rows / (ts + te + tr), unit='samples/s', bytes_per_sample=b_s + b_e + b_r,
                  gbs=rows * (b_s + b_e + b_r) / (ts + te + tr) / 1e9, frac=rows * (b_s + b_e + b_r) / (ts + te + tr) / 1e9 / peak),
        hbm_peak_gbs=peak, hbm_peak_source=src)
    return out


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
