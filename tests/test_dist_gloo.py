"""N > 1 host logic on CPU: two gloo ranks (world_size 2, 127.0.0.1) build the partitioned hierarchy,
exchange the byte payloads dist_init() uses (NCCL id broadcast, IPC-handle all-gather) and check that
every rank holds the same hierarchy and a disjoint, complete share of the row blocks."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import metric_amg_examples_b200 as mamg
        from metric_amg_examples_b200 import params, problems
        from metric_amg_examples_b200.hierarchy import allgather_bytes, broadcast_bytes, owned_blocks
        s = problems.emi_system(3, 12, gamma=1e6)
        nparts = 4
        part = problems.slab_partition(s, nparts)
        H = mamg.Hierarchy(s.A, params.default_metric_parameters, s.interface_dofs, part=part)
        ex = H.export()
        h = hashlib.sha256()
        for L in ex["levels"]:
            for k in ("indptr", "indices", "data", "agg", "color", "part", "patch_dofs", "patch_color"):
                h.update(np.ascontiguousarray(L[k]).tobytes())
        digests = allgather_bytes(h.digest(), None)
        assert len(set(digests)) == 1, "ranks built different hierarchies"
        # the NCCL-id broadcast path (payload made on rank 0 only)
        payload = bytes(range(128)) if rank == 0 else bytes(128)
        assert broadcast_bytes(payload, 0, None) == bytes(range(128))
        # block ownership: disjoint and complete; rows of a rank's blocks = rows of its parts
        mine = list(owned_blocks(nparts, rank, world))
        alls = allgather_bytes(bytes(mine), None)
        flat = sorted(b for blk in alls for b in blk)
        assert flat == list(range(nparts))
        rows = int(np.isin(ex["levels"][0]["part"], mine).sum())
        counts = allgather_bytes(rows.to_bytes(8, "little"), None)
        assert sum(int.from_bytes(c, "little") for c in counts) == s.ndofs
        # no Schwarz patch straddles parts for the EMI slabs (interface kept inside one part)
        L0 = ex["levels"][0]
        for p in range(len(L0["patch_seed"])):
            d = L0["patch_dofs"][L0["patch_ptr"][p]:L0["patch_ptr"][p + 1]]
            assert len(set(L0["part"][d])) == 1
        with pytest.raises(ValueError):
            owned_blocks(3, rank, world)
        # x-strips cut across the interface: every rank owns patches (by the part of the seed), some
        # patches straddle parts, aggregates still do not; the hierarchy is the same on both ranks
        px = problems.slab_partition(s, nparts, axis=0)
        Hx = mamg.Hierarchy(s.A, params.default_metric_parameters, s.interface_dofs, part=px)
        X0 = Hx.export()["levels"][0]
        seeds_mine = int(np.isin(X0["part"][X0["patch_seed"]], mine).sum())
        assert seeds_mine > 0
        got = allgather_bytes(seeds_mine.to_bytes(8, "little"), None)
        assert sum(int.from_bytes(c, "little") for c in got) == len(X0["patch_seed"])
        straddle = sum(len(set(X0["part"][X0["patch_dofs"][X0["patch_ptr"][p]:X0["patch_ptr"][p + 1]]])) > 1
                       for p in range(len(X0["patch_seed"])))
        assert straddle > 0
        ok = X0["agg"] >= 0
        assert np.array_equal(Hx.export()["levels"][1]["part"][X0["agg"][ok]], X0["part"][ok])
        hx = hashlib.sha256(np.ascontiguousarray(X0["color"]).tobytes() + np.ascontiguousarray(X0["patch_color"]).tobytes())
        assert len(set(allgather_bytes(hx.digest(), None))) == 1
        # halo mode: a rank waits for exactly the ranks it sends to, so the neighbour relation of the halo
        # lists must be symmetric on every level; what rank r sends to q is what q's rows read from r
        from metric_amg_examples_b200.hierarchy import halo_send_rows
        import pickle
        for L in Hx.export()["levels"][:3]:
            send = halo_send_rows(L, nparts, rank, world)
            blob = pickle.dumps({q: len(v) for q, v in send.items()})
            blob = blob + bytes(256 - len(blob))
            tables = [pickle.loads(t) for t in allgather_bytes(blob, None)]
            for q, cnt in tables[rank].items():
                assert rank in tables[q] and cnt > 0, "halo neighbour relation is not symmetric"
            assert set(send) <= {rank - 1, rank + 1}, "x-strips: only the adjacent ranks are neighbours"
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_host_logic():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: 1, 1: 1}
