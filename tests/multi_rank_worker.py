"""Worker for the multi-rank tests (launched once per rank).

gpu mode  : one process per GPU; shards landmarks with (rank, world), exchanges the
            NCCL id through a file, runs the sharded LM solve (NCCL all-reduce of the
            partial RCS) and writes its summary + state.
gloo mode : CPU only; checks the host-side sharding logic with torch.distributed
            (gloo): per-shard oracle costs / gradients all-reduced == whole problem.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import pba_b200 as pb  # noqa: E402


def main():
    mode, rank, world, outdir = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    kind = int(sys.argv[5]) if len(sys.argv) > 5 else pb.MODE_PHOTOMETRIC
    prob, _ = pb.make_scene(kind, 14, 900, "pinhole")
    hub = 9.0 if kind == pb.MODE_PHOTOMETRIC else 1.0
    if mode == "gpu":
        from pba_b200.engine import nccl_unique_id
        opts = pb.BundleAdjustmentOptions(verbosity_level=0, huber_parameter=hub, device=rank)
        eng = pb.Engine(prob, opts, rank=rank, world_size=world)
        idfile = os.path.join(outdir, "nccl_id.bin")
        if rank == 0:
            with open(idfile + ".tmp", "wb") as f:
                f.write(nccl_unique_id())
            os.rename(idfile + ".tmp", idfile)
        t0 = time.time()
        while not os.path.exists(idfile):
            time.sleep(0.05)
            assert time.time() - t0 < 60
        eng.comm_init(open(idfile, "rb").read())
        cost0 = eng.evaluate(True)
        s = eng.minimize()
        poses, rho, aff = eng.get_state()
        np.savez(os.path.join(outdir, "rank%d.npz" % rank), poses=poses, rho=rho, first=eng.first_landmark,
                 cost0=cost0, initial_cost=s.initial_cost, final_cost=s.final_cost, iterations=s.num_iterations,
                 costs=np.array([i["cost"] for i in s.iterations]), collective=eng.collective)
        eng.close()
    else:
        import oracle_ffi as of
        import torch
        import torch.distributed as dist
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[6], rank=rank, world_size=world)
        b = pb.partition_landmarks(prob.lm_obs_ptr, world)
        shard = prob.subset_landmarks(b[rank], b[rank + 1])
        cost, r, J = of.evaluate("oracle", shard, True, hub, threads=1)
        # gradient w.r.t. the (replicated) poses from this shard's blocks
        g = np.zeros((prob.n_poses, 6))
        o = 0
        for l in range(shard.n_landmarks):
            for k in range(int(shard.lm_obs_ptr[l]), int(shard.lm_obs_ptr[l + 1])):
                g[shard.lm_host[l]] += J[k][:, 0:6].T @ r[k]
                g[shard.obs_target[k]] += J[k][:, 6:12].T @ r[k]
        t = torch.tensor(np.r_[cost, shard.n_obs, shard.n_landmarks, g.ravel()])
        dist.all_reduce(t)
        if rank == 0:
            json.dump({"cost": t[0].item(), "n_obs": t[1].item(), "n_lm": t[2].item(), "bounds": b},
                      open(os.path.join(outdir, "gloo.json"), "w"))
            np.save(os.path.join(outdir, "gloo_grad.npy"), t[3:].numpy())
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
