"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the hypercube range sharded over two
ranks (NCCL all-reduce of the per-iteration sums) must reproduce the single-GPU iteration --
integers exactly, fp64 sums to 1e-12 (only the order of the partial sums differs), because the
Philox stream is a pure function of (seed, iteration, hypercube, sample)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

LIMITS = 4 * [[0., 1.]]
KW = dict(neval=400000, seed=4242)
NITN = 3


def _run(mpi, light):
    import vegas_b200 as vegas
    if light == 'callback':
        os.environ['VB200_ITEM'] = '512'                 # callback path (sample -> host integrand -> reduce), split chunks
    elif light:
        os.environ['VB200_LIGHT'] = '1'
    f = vegas.integrands.Ridge(dim=4, N=3)
    integ = vegas.Integrator(LIMITS, mpi=mpi, fused=(light != 'callback'), **KW)
    recs = []
    integ._trace = lambda rec: recs.append(dict(rec))
    r = integ(f, nitn=NITN)
    return recs, float(r.mean), float(r.sdev), np.array(integ.sigf), np.array(integ.map.grid), integ._ctx.last_launch()


def _worker(rank, world, port, light, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    out = _run(True, light)
    q.put((rank,) + out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('light', [False, True, 'callback'])
def test_two_gpus_match_one(light):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000) + [False, True, 'callback'].index(light)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, light, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    one = _run(False, light)
    os.environ.pop('VB200_LIGHT', None)
    os.environ.pop('VB200_ITEM', None)
    for o in outs:
        recs, mean, sdev, sigf, grid, geom = o[1:]
        assert len(recs) == NITN
        for a, b in zip(recs, one[0]):
            assert a['last_neval'] == b['last_neval']
            np.testing.assert_allclose(a['mean'], b['mean'], rtol=1e-12)
            np.testing.assert_allclose(a['var'], b['var'], rtol=1e-11)
            np.testing.assert_allclose(a['sum_sigf'], b['sum_sigf'], rtol=1e-12)
            assert np.array_equal(a['n_f'], b['n_f'])
            # from the second iteration on the two runs start from grids that differ in the last bit
            # (atomic summation order), and inc = g[i+1] - g[i] amplifies that ~1e3x in J^2: two
            # single-GPU runs differ by the same ~1e-12 in the tail bins
            np.testing.assert_allclose(a['sum_f'], b['sum_f'], rtol=1e-10, atol=1e-300)
        np.testing.assert_allclose(mean, one[1], rtol=1e-12)
        np.testing.assert_allclose(sigf, one[3], rtol=1e-8, atol=1e-300)
        np.testing.assert_allclose(grid, one[4], rtol=1e-10, atol=1e-14)
        if light is True:
            assert geom['threads'] > 128, geom                 # the light geometry really ran on the shards
    assert np.array_equal(outs[0][5], outs[1][5])              # identical grids on both ranks


def _restrat_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    q.put((rank,) + _restrat(True))
    dist.barrier()
    dist.destroy_process_group()


def _restrat(mpi):
    import vegas_b200 as vegas
    f = vegas.integrands.Genz('gaussian', [12., 12., .2, .2], [.5, .3, .5, .5])
    integ = vegas.Integrator(LIMITS, mpi=mpi, neval=200000, seed=99)
    integ(f, nitn=3)
    new = vegas.restratify(integ, f, nitn=1, ndy=5)
    return ([int(v) for v in new.nstrat], float(new.I.mean), [[float(g.mean) for g in row] for row in new.dI])


def test_restratify_two_gpus_match_one():
    """the stratification profile is all-reduced like the iteration sums: same dI, same new nstrat"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_restrat_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    one = _restrat(False)
    for o in outs:
        assert o[1] == one[0]
        np.testing.assert_allclose(o[2], one[1], rtol=1e-9)
        np.testing.assert_allclose(o[3], one[2], rtol=1e-8, atol=1e-300)
