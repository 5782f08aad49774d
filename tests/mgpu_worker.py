"""torchrun worker: N-GPU bundle adjustment (points sharded, one all-reduce of the reduced camera
system per LM iteration) must reproduce the 1-GPU solve; pair-list sharding must reproduce the
unsharded match lists."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from opensfm_b200 import bundle, dist as odist, matching, synthetic as syn  # noqa: E402

rank, world, local = odist.init_from_env("nccl")
torch.cuda.set_device(local)
allreduce = odist.make_allreduce(device=local)

sc = syn.cube_scene(16, 3000, 1.0, max_obs_per_point=8)
pb = syn.scene_to_problem(sc)
multi_cb = bundle.solve(pb, device=local, rank=rank, world=world, allreduce=allreduce)   # torch.distributed callback
multi = bundle.solve(pb, device=local, rank=rank, world=world, allreduce="nccl")         # the library's own communicator
single = bundle.solve(pb, device=local)
assert multi_cb["summary"]["iterations"] == multi["summary"]["iterations"]
assert abs(multi_cb["summary"]["final_cost"] - multi["summary"]["final_cost"]) <= 1e-9 * multi["summary"]["final_cost"]
assert np.abs(multi_cb["points"] - multi["points"]).max() < 1e-8
sm, ss = multi["summary"], single["summary"]
assert sm["termination"] == "CONVERGENCE", sm
assert sm["iterations"] == ss["iterations"], (sm, ss)
# the sharded run sums the reduced camera system in a different order and the PCG stops at a relative
# residual of 1e-8: agreement is to solver tolerance, not bitwise
print("rank", rank, "cost", sm["final_cost"], ss["final_cost"], "dpts", np.abs(multi["points"] - single["points"]).max(),
      "dinst", np.abs(multi["inst"] - single["inst"]).max(), flush=True)
assert abs(sm["final_cost"] - ss["final_cost"]) <= 1e-7 * ss["final_cost"], (sm["final_cost"], ss["final_cost"])
assert np.abs(multi["points"] - single["points"]).max() < 1e-6
assert np.abs(multi["inst"] - single["inst"]).max() < 1e-6
assert np.abs(multi["cam_params"] - single["cam_params"]).max() < 1e-6
assert np.abs(multi["reprojection_errors"] - single["reprojection_errors"]).max() < 1e-6
assert sm["num_observations_local"] < pb.num_observations  # really sharded

feats = {s: sc.features_of_shot(s)[0] for s in range(sc.num_shots)}
pairs = [(i, j) for i in range(8) for j in range(i + 1, 8)]
sizes = {s: len(f) for s, f in feats.items()}
mine = matching.shard_pairs(pairs, sizes, world)[rank]
pm = matching.PairMatcher(device=local)
for s in sorted({i for p in mine for i in p}):
    pm.add(s, feats[s])
local_res = pm.match_pairs(mine, {"lowes_ratio": 0.8}) if mine else {}
allres = odist.gather_pair_results(local_res, world)
if rank == 0:
    pm_all = matching.PairMatcher(device=local)
    for s in range(8):
        pm_all.add(s, feats[s])
    ref = pm_all.match_pairs(pairs, {"lowes_ratio": 0.8})
    assert sorted(allres) == sorted(ref)
    for p in pairs:
        assert np.array_equal(allres[p], ref[p]), p
    print("MGPU_OK world=%d iterations=%d cost=%.9e" % (world, sm["iterations"], sm["final_cost"]))
torch.distributed.barrier()
torch.distributed.destroy_process_group()
