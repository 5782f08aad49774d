"""torchrun worker: one traced multi-GPU solve of the bench workload (OSFM_BA_TRACE=1)."""
import os, sys, time
sys.path.insert(0, ".")
import torch
from opensfm_b200 import bundle, dist as odist
import bench
rank, world, local = odist.init_from_env("nccl")
torch.cuda.set_device(local)
pb, feats, pairs, w = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else "c4")
for rep in range(3):
    t0 = time.perf_counter()
    r = bundle.solve(pb, device=local, rank=rank, world=world, allreduce="nccl")
    if rank == 0:
        s = r["summary"]
        print("rep %d solve %.1f ms run %.1f device %.1f  lin %.2f schur %.2f pcg %.2f back %.2f" % (
            rep, 1e3 * (time.perf_counter() - t0), 1e3 * s["time_run_s"], s["time_device_ms"], s["time_linearize_ms"],
            s["time_schur_ms"], s["time_pcg_ms"], s["time_backsub_ms"]), flush=True)
torch.distributed.barrier()
torch.distributed.destroy_process_group()
