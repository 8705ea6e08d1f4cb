"""Worker of tests/test_multigpu.py::test_two_processes_two_gpus (launched with torch.distributed.run, one rank per GPU):
the REAL multi-process slab path — CUDA IPC mapping of the peers' publications, flags and pulls across GPUs over NVLink,
migration through the outboxes — checked against the oracle and against a single-GPU run of the same system."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import importlib  # noqa: E402

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import __graft_entry__ as graft  # noqa: E402
from bench import make_workload  # noqa: E402


def main():
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = graft.load_package()
    mg = importlib.import_module(pkg.__name__ + ".multigpu")
    w = dict(make_workload("c4", 40 ** 3))
    w["vel"] = (w["vel"] * 3.0).astype(np.float32)  # hot: atoms cross the slab borders within the run
    out = {}
    for migrate_every in (0, 5):
        sim = mg.SlabSimulation(pkg, w, rank, world, lr, dist, migrate_every=migrate_every, headroom=2.5)
        sim.step_async(30)
        sim.sync()
        par = mg.parity_digest(sim, w, dist, graft)
        ke, pe = sim.h.mg_get_energies()
        e = torch.tensor([ke, pe, float(sim.h.mg_owned_count())], dtype=torch.float64, device=sim.dev)
        dist.all_reduce(e)
        ids = [None] * world
        dist.all_gather_object(ids, sim.owned_original_ids())
        if rank == 0:
            allids = np.concatenate(ids)
            out[f"migrate_every_{migrate_every}"] = {
                "parity": {k: par[k] for k in ("count", "count_match", "xor_match", "sum_match")},
                "ke": float(e[0]), "pe": float(e[1]), "atoms": int(e[2]),
                "every_atom_owned_once": bool(len(allids) == w["n"] and len(np.unique(allids)) == w["n"]),
                "moved_rank": int(sum(len(np.setdiff1d(ids[g], sim.order[sim.bounds[g]:sim.bounds[g + 1]])) for g in range(world)))}
        sim.close()
        dist.barrier()
    if rank == 0:
        h = pkg.Handle(w["n"], device=0)
        h.set_forcefield(w["eps"], w["sigma"], w["kcoul"], w["cutoff"], True)
        h.set_system(w["pos"], w["vel"], w["mass"], w["charge"])
        h.step(30, w["dt"])
        ke1, pe1 = h.get_energies()
        h.close()
        out["single_gpu"] = {"ke": ke1, "pe": pe1}
        print("TWO_RANK_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
