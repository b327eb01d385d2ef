"""Time loop of main.cc:908-990 on N GPUs against the same loop on one GPU (launched by hand or by
tests/test_gpu_multi.py):

  python tests/multi_gpu_step_check.py --record /tmp/step_n1.json                       # one GPU: writes the record
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
      --master-port 29533 tests/multi_gpu_step_check.py --record /tmp/step_n1.json      # N GPUs: compares with it

3-D Q2 channel (input_channel.json at dim 3), mg_min_level 1 so that every rank holds cells on every level:
partitioned level operators, transfers, smoothers, coarse solve and Krylov reductions.  Criterion: identical
Newton and GMRES iteration counts step by step, residual norms and the solution norm to 1e-6."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from dealii_ns_gls_b200.driver import ChannelParameters, Driver  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--record", required=True)
    ap.add_argument("--refinements", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--mg-number", default="double")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank, lr = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    p = ChannelParameters(dim=3, fe_degree=2, n_global_refinements=a.refinements, mg_min_level=1, mg_number=a.mg_number)
    d = Driver(p, device=dev, n_ranks=world, rank=rank)
    rec = []
    for _ in range(a.steps):
        r = d.step()
        sol = d.solution.get_current_solution()
        n2 = torch.dot(sol[:d.ns_operator.n_owned], sol[:d.ns_operator.n_owned]).reshape(1)
        ghosts_zero = bool((sol[d.ns_operator.n_owned:] == 0).all())
        if world > 1:
            dist.all_reduce(n2)
        rec.append(dict(newton=r["newton_iterations"], linear=r["linear_iterations"], dt=r["dt"],
                        residuals=[float(x) for x in r["newton_residuals"]], l2=float(n2.sqrt()), ghosts_zero=ghosts_zero))
    ok = True
    if world == 1:
        with open(a.record, "w") as f:
            json.dump(rec, f)
        print("N=1 record:", [(s["newton"], s["linear"]) for s in rec], flush=True)
    else:
        with open(a.record) as f:
            ref = json.load(f)
        for s, t in zip(rec, ref):
            same = s["newton"] == t["newton"] and s["linear"] == t["linear"]
            close = abs(s["l2"] / t["l2"] - 1) < 1e-6 and abs(s["dt"] / t["dt"] - 1) < 1e-10 and \
                np.allclose(s["residuals"][:2], t["residuals"][:2], rtol=1e-6)
            ok = ok and same and close and s["ghosts_zero"]
        print(f"rank {rank}/{world}: counts {[(s['newton'], s['linear']) for s in rec]} vs N=1 "
              f"{[(t['newton'], t['linear']) for t in ref]}  l2 {rec[-1]['l2']:.10e} vs {ref[-1]['l2']:.10e}  "
              f"{'OK' if ok else 'FAIL'}", flush=True)
        t = torch.tensor([0 if ok else 1], device=dev)
        dist.all_reduce(t)
        ok = int(t.item()) == 0
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
