"""Multi-GPU parity check, launched by torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check.py

Every rank builds the same partitioned hierarchy, joins the NCCL communicator and runs the
row-distributed apply / PCG; rank 0 compares with the CPU oracle on the same hierarchy and with the
expected communication pattern.  Exit code 0 = all checks passed on every rank.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("MAMG_DIST_MIN_ROWS", "2000")

import metric_amg_examples_b200 as mamg  # noqa: E402
from metric_amg_examples_b200 import params, problems  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import Oracle
    cases = [
        ("bidomain3d", problems.bidomain_system(3, 16, gamma=1e4), params.parameters_metric_schwarz, 1e-8),
        ("emi3d", problems.emi_system(3, 24, gamma=1e6), params.default_metric_parameters, 1e-10),
        ("bidomain2d", problems.bidomain_system(2, 96, gamma=1e3), params.parameters_metric, 1e-8),
        # x-strips: every rank owns a strip of the interface, 2-ring patches straddle the cuts (general kernel)
        ("emi3d_strips", problems.emi_system(3, 24, gamma=1e6), params.default_metric_parameters, 1e-10),
    ]
    ok = True
    for name, s, prm, tol in cases:
        nparts = 4 if world in (1, 2, 4) else world
        part = problems.slab_partition(s, nparts, axis=0 if name.endswith("strips") else None)
        H = mamg.Hierarchy(s.A, prm, s.interface_dofs, part=part)
        H.to_device(local, rank=rank, world=world)   # halo mode unless MAMG_HALO=0
        H.dist_init()
        r = np.random.default_rng(0).standard_normal(s.ndofs)
        H.collective_count(reset=True)
        z = H.apply(r)
        ncoll = H.collective_count()
        b, xt = s.random_rhs(1)
        if rank == world - 1:
            import time
            time.sleep(1.5)   # injected skew: the last rank enters the solve late; its peers spin on its flags meanwhile
        x, info = H.pcg(b, tolerance=tol, maxiter=500)
        # every rank must hold the same result (vectors are complete on every rank)
        zt = torch.from_numpy(z).cuda()
        z0 = zt.clone()
        dist.broadcast(z0, src=0)
        same = float((zt - z0).abs().max()) == 0.0
        if rank == 0:
            orc = Oracle(H.export(), "multicolor")
            zo = orc.apply(r)
            err = np.linalg.norm(z - zo) / np.linalg.norm(zo)
            _, ref = orc.pcg(b, tolerance=tol, maxiter=500)
            good = err < 1e-10 and abs(info["niters"] - ref["niters"]) <= 1 and \
                np.linalg.norm(x - xt) / np.linalg.norm(xt) < 1e-6 and (world == 1 or ncoll > 0)
            print(f"[dist_check] {name}: halo={os.environ.get('MAMG_HALO', '1')} world={world} nparts={nparts} dev_GB={H.device_bytes() / 1e9:.3f} apply rel err {err:.2e}, iters {info['niters']} "
                  f"(oracle {ref['niters']}), collectives per apply {ncoll}, {'OK' if good else 'FAIL'}", flush=True)
            ok &= good
        ok &= same
        if not same:
            print(f"[dist_check] rank {rank}: {name} result differs from rank 0", flush=True)
        del H
    flag = torch.tensor([0 if ok else 1], device="cuda")
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
