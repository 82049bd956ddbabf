"""Kernel-tuning script (not a test): the per-step workload of ONE rank of a `world`-way z-slab decomposition of an
n^3 domain on a single GPU, without its neighbours (the exchanges are skipped, ghost planes go stale): kernel
times of a slab rank without needing `world` GPUs.  usage: dbg_slab_rank.py n world rank [steps]"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from fluidnet_b200 import slab as slab_mod, tfluids
n, world, rank = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
batch_np, mconf, mnp = bench.make_problem(n)
tb = {k: torch.from_numpy(v) for k, v in batch_np.items()}


class Lone(slab_mod.NativeSlabSimulator):
    pass


import torch.distributed as dist
orig = dist.broadcast_object_list
dist.broadcast_object_list = lambda *a, **k: None            # no process group here
ctx = tfluids.context(0)
real_init = ctx.lib.tfl_comm_init
sim = None
try:
    # nil id -> the library keeps the rank / world but makes no communicator
    ctx.lib.tfl_comm_init.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]
    class L(slab_mod.NativeSlabSimulator):
        def __init__(self, *a, **k):
            pass
    sim = L.__new__(L)
    from fluidnet_b200 import model as fmodel, simulate
    sim.group, sim.rank, sim.world, sim.device, sim.ctx = None, rank, world, torch.device("cuda", 0), ctx
    ctx.check(ctx.lib.tfl_comm_init(ctx.h, None, rank, world))
    sim.mconf = dict(mconf); sim.mc = simulate.make_mconf(sim.mconf)
    sim.model = fmodel.ProjectionModel(mnp["layers"], True)
    host = lambda k: np.ascontiguousarray(batch_np[k], np.float32) if batch_np.get(k) is not None else None
    sim._shape = tuple(batch_np["flags"].shape)
    arrs = [host(k) for k in ("flags", "UBC", "UBCInvMask", "densityBC", "densityBCInvMask")]
    h = C.c_void_p()
    ctx.use_current_stream()
    ctx.check(ctx.lib.tfl_slab_sim_create(ctx.h, n, n, n, 2, *[a.ctypes.data if a is not None else None for a in arrs], C.byref(h)))
    sim.h = h
    ctx.check(ctx.lib.tfl_slab_sim_upload(ctx.h, h, host("pDiv").ctypes.data, host("UDiv").ctypes.data, host("density").ctypes.data))
finally:
    dist.broadcast_object_list = orig
for _ in range(3):
    sim.step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(steps):
    sim.step()
b.record()
torch.cuda.synchronize()
ctx.trace_faults()
print("rank %d of %d, %d^3: %.3f ms / step (no exchanges)" % (rank, world, n, a.elapsed_time(b) / steps))
