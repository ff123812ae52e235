"""per-rank kernel time and sample count of a sharded run (where does a multi-GPU step's time go?):
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/skew_probe.py [ridge|pathint] [neval]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import vegas_b200 as vegas

local = int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
rank, world = dist.get_rank(), dist.get_world_size()
what = sys.argv[1] if len(sys.argv) > 1 else 'ridge'
neval = float(sys.argv[2]) if len(sys.argv) > 2 else 1e8
F = vegas.integrands
if what == 'ridge':
    f, limits, kw = F.Ridge(8, N=1000), 8 * [[0., 1.]], {}
else:
    f, limits, kw = F.PathIntegral(T=4., ndT=10, x0list=np.linspace(0, 2., 6)), 10 * [[-np.pi / 2, np.pi / 2]], dict(alpha=0.1)
if os.environ.get('SLAB'):
    kw['slab'] = int(os.environ['SLAB'])
integ = vegas.Integrator(limits, neval=neval, seed=5, mpi=True, **kw)
integ(f, nitn=5)
integ._timing = []
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
r = integ(f, nitn=5)
torch.cuda.synchronize(); dist.barrier()
wall = (time.perf_counter() - t0) / 5 * 1e3
rows = []
for ev, tot in integ._timing:
    rows.append([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), float(tot)])
mine = torch.tensor(rows, dtype=torch.float64, device='cuda')
allr = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(allr, mine)
# the all-reduce alone
buf = torch.zeros(3 + 2 * 8 * 1001 + 2 + world, dtype=torch.float64, device='cuda')
for _ in range(5):
    dist.all_reduce(buf)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    dist.all_reduce(buf)
e1.record(); torch.cuda.synchronize()
if rank == 0:
    a = torch.stack(allr).cpu().numpy()          # [rank][itn][plan, kernel, reduce+plan_next, samples]
    print('slab', integ._slab(world), end='  ')
    print('%s neval=%.0e world=%d: wall %.3f ms/iteration, all-reduce of the packed buffer alone %.3f ms' % (what, neval, world, wall, e0.elapsed_time(e1) / 20))
    print('kernel ms by rank (mean over iterations):', np.round(a[:, :, 1].mean(axis=1), 3))
    print('samples  by rank (last iteration, 1e6):  ', np.round(a[:, -1, 3] / 1e6, 3))
    print('plan ms by rank:', np.round(a[:, :, 0].mean(axis=1), 3), ' exchange+plan-ahead ms by rank:', np.round(a[:, :, 2].mean(axis=1), 3))
    print('kernel: max/mean over ranks = %.3f   samples: max/mean = %.3f' % (a[:, :, 1].mean(axis=1).max() / a[:, :, 1].mean(), a[:, -1, 3].max() / a[:, -1, 3].mean()))
dist.destroy_process_group()
