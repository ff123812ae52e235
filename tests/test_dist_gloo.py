"""world_size-2 test of the N>1 host logic on CPU (gloo): block-cyclic sharding of the hypercube
range + the per-iteration all-reduce reproduce the single-rank sums, and every rank derives the
same adapted grid afterwards.  The per-cube partial results are produced here by the CPU oracle
(the checker), sharded exactly as the device shards them."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NH, SLAB, DIM, NINC = 1080, 256, 4, 50


def _per_cube_data():
    rng = np.random.default_rng(7)
    mean_c = rng.standard_normal(NH)                 # per-cube contributions to mean / var / sigf
    var_c = rng.random(NH)
    sigf_c = rng.random(NH) ** 0.75
    n_c = rng.integers(2, 40, NH)
    bins = rng.integers(0, NINC, (NH, DIM))          # one training point per cube (enough to test sums)
    fdv2 = rng.random(NH)
    return mean_c, var_c, sigf_c, n_c, bins, fdv2


def _partials(idx):
    mean_c, var_c, sigf_c, n_c, bins, fdv2 = _per_cube_data()
    acc = torch.tensor([mean_c[idx].sum(), var_c[idx].sum(), sigf_c[idx].sum()], dtype=torch.float64)
    sum_f = torch.zeros((DIM, NINC), dtype=torch.float64)
    n_f = torch.zeros((DIM, NINC), dtype=torch.int64)
    for h in idx:
        for d in range(DIM):
            sum_f[d, bins[h, d]] += fdv2[h]
            n_f[d, bins[h, d]] += 1
    return acc, sum_f, n_f, int(n_c[idx].sum()), int(n_c[idx].max())


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import vegas_b200 as vegas
    from vegas_b200._integrator import _local_cubes, allreduce_iteration, allreduce_neval_stats
    idx = _local_cubes(NH, SLAB, rank, world)
    acc, sum_f, n_f, tot, mx = _partials(idx)
    status = torch.zeros(1, dtype=torch.int32)
    # the packed form Integrator.__call__ uses -- ONE fp64 SUM all-reduce per iteration -- must give the same answers
    from vegas_b200._integrator import exchange_iteration, pack_iteration
    nacc, nh = 3, DIM * NINC
    buf_f = torch.zeros(nacc + 2 * nh + 2 + world, dtype=torch.float64)
    buf_f[:nacc] = acc
    buf_f[nacc:nacc + nh] = sum_f.reshape(-1)
    nan_flag = torch.tensor([1 if rank == 1 else 0], dtype=torch.int32)          # rank 1 saw a NaN
    pack_iteration(buf_f, nacc, nh, n_f, nan_flag, tot, mx, rank)
    exchange_iteration(buf_f)
    allreduce_iteration(acc, sum_f, n_f, status)
    tot, mx = allreduce_neval_stats(tot, mx, 'cpu')
    assert torch.equal(buf_f[:3], acc) and torch.equal(buf_f[3:3 + nh].reshape(DIM, NINC), sum_f)
    assert torch.equal(buf_f[3 + nh:3 + 2 * nh].reshape(DIM, NINC).to(torch.int64), n_f)
    tail = buf_f[3 + 2 * nh:]
    assert int(tail[0]) == tot and int(tail[1]) == 1 and int(tail[2:].max()) == mx
    # every rank runs the same deterministic host adapt on the reduced histogram
    m = vegas.AdaptiveMap(DIM * [[0., 1.]], ninc=NINC)
    m._accumulate_training(sum_f.numpy(), n_f.numpy())
    m.adapt(alpha=0.5)
    integ = vegas.Integrator(DIM * [[0., 1.]], neval=1e4, mpi=True)
    q.put((rank, acc.numpy(), sum_f.numpy(), n_f.numpy(), tot, mx, m.grid.copy(), len(idx),
           integ._rank_world(), integ._slab(world)))
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_single_rank():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    acc1, sum_f1, n_f1, tot1, mx1 = _partials(np.arange(NH))
    assert sum(o[7] for o in out) == NH
    for o in out:
        np.testing.assert_allclose(o[1], acc1.numpy(), rtol=1e-13)
        np.testing.assert_allclose(o[2], sum_f1.numpy(), rtol=1e-13)
        assert np.array_equal(o[3], n_f1.numpy())
        assert o[4] == tot1 and o[5] == mx1
        assert o[8] == (o[0], world) and o[9] % 256 == 0
    assert np.array_equal(out[0][6], out[1][6])          # identical grids on both ranks
    assert np.array_equal(out[0][1], out[1][1])          # identical reduced sums on both ranks
