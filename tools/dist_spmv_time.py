"""torchrun program: times the row-partitioned SpMV (halo exchange included) of this rank's slab.
Usage: torchrun --nproc-per-node N tools/dist_spmv_time.py NX NY NZ"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from opm_simulators_legacy_b200.distributed import DistributedSolver  # noqa: E402
from opm_simulators_legacy_b200.jacobian import synth_blackoil_jacobian  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
nx, ny, nz = (int(a) for a in sys.argv[1:4])
s = synth_blackoil_jacobian(nx, ny, nz, perm="lognormal")
g = DistributedSolver(s, local)
g.set_values_dev(g.vals)
x = g.rhs.clone(); y = torch.zeros_like(x)
for _ in range(10):
    g.spmv_dev(x, y)
dist.barrier(); torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        g.spmv_dev(x, y)
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 50 * 1e3)
t = torch.tensor([min(ts)], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if dist.get_rank() == 0:
    print(f"{nx}x{ny}x{nz} on {dist.get_world_size()} GPUs, axis {g.axis}, OPMGPU_HALO_OVERLAP={os.environ.get('OPMGPU_HALO_OVERLAP', '1')}: "
          f"SpMV incl. halo + staging copy {float(t):.1f} us (max over ranks); rows {g.N}")
dist.barrier()
dist.destroy_process_group()
