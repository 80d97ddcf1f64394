"""Timing breakdown of the multi-GPU plumbing (run under torchrun): interface exchange alone vs a CG iteration."""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import torch, torch.distributed as dist
from bench import slab_interfaces
pkg = g.load_package()
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
eng = pkg.lib.Engine(lr)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
X = np.linspace(0, 1, n + 1)
grid = pkg.simplexgrid(X, X, np.linspace(float(rank), float(rank + 1), n + 1))
FES = pkg.FESpace(pkg.H1P2(1, 3), grid)
mesh = eng.mesh_set(grid.coords, grid.cellnodes, grid.cellregions, grid.cellvolumes)
sp = eng.space_set(mesh, 2, 1, FES.celldofs, FES.ndofs)
pat = eng.pattern_build([sp])
uid = [pkg.lib.Engine.dist_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
eng.dist_init(rank, world, uid[0])
plan = slab_interfaces(pkg, FES, rank, world)
eng.dist_set_interfaces(pat, plan)
eng.assemble_bilinear(pat, eng.make_opdesc([(0, 1)], [(0, 1)]))
eng.assemble_linear(pat, eng.make_opdesc([(0, 0)], kernel_id=pkg.lib.kernel_id("sincos301"), params=[1.0]))
xyz = FES.dof_coordinates()
onb = (xyz[:, 0] == 0) | (xyz[:, 0] == 1) | (xyz[:, 1] == 0) | (xyz[:, 1] == 1) | (xyz[:, 2] == 0) | (xyz[:, 2] == float(world))
eng.apply_penalties(pat, np.nonzero(onb & (plan.owned == 1))[0] + 1, None, 1e30)
out = {}
for _ in range(5):
    eng.dist_sum_rhs(pat)
eng.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(200):
    eng.dist_sum_rhs(pat)
eng.synchronize()
out["exchange_us"] = (time.perf_counter() - t0) / 200 * 1e6
eng.dist_cg(pat, rtol=1e-30, maxit=5)
for k in (20, 60):
    eng.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    eng.dist_cg(pat, rtol=1e-30, maxit=k)
    eng.synchronize()
    out[f"cg_{k}_ms"] = (time.perf_counter() - t0) * 1e3
out["cg_iteration_ms"] = (out["cg_60_ms"] - out["cg_20_ms"]) / 40
if rank == 0:
    print(json.dumps(out))
dist.barrier(); dist.destroy_process_group(); eng.close()
