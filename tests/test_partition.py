"""Partitioned hierarchies (multi-GPU layout): host-side invariants on CPU, execution of the block
layout on one GPU, and (tests/dist_check.py, launched with torchrun on 2 GPUs) the NCCL path."""
import os

import numpy as np
import pytest

import metric_amg_examples_b200 as mamg
from metric_amg_examples_b200 import params, problems
from oracle import Oracle


def test_slab_partition_shapes():
    s = problems.bidomain_system(3, 12, gamma=10.0)
    part = problems.slab_partition(s, 4)
    nv = s.W[0].dim()
    assert part.min() == 0 and part.max() == 3 and np.array_equal(part[:nv], part[nv:])
    counts = np.bincount(part)
    assert counts.max() - counts.min() <= 2 * 13 * 13 * 2
    e = problems.emi_system(3, 16, gamma=1e3)
    pe = problems.slab_partition(e, 2)
    # the interface dofs of both sides (and their 2-ring patches) live in one part
    assert len(set(pe[e.interface_dofs])) == 1
    # x-strips cut across the interface: every part owns a strip of it, both sides of a vertex together
    px = problems.slab_partition(e, 4, axis=0)
    assert np.all(np.bincount(px[e.interface_dofs], minlength=4) > 0)
    ni = len(e.interface_dofs) // 2
    assert np.array_equal(px[e.interface_dofs[:ni]], px[e.interface_dofs[ni:]])
    with pytest.raises(ValueError):
        problems.slab_partition(e, 4, axis=3)


@pytest.mark.parametrize("prm", ["parameters_metric", "parameters_standard"])
def test_aggregates_do_not_cross_parts(prm):
    s = problems.bidomain_system(2, 32, gamma=1e3)
    part = problems.slab_partition(s, 4)
    H = mamg.Hierarchy(s.A, getattr(params, prm), s.interface_dofs if "metric" in prm else None, part=part)
    ex = H.export()
    for l in range(len(ex["levels"]) - 1):
        L, Lc = ex["levels"][l], ex["levels"][l + 1]
        agg, p = L["agg"], L["part"]
        ok = agg >= 0
        assert np.array_equal(Lc["part"][agg[ok]], p[ok])        # every member sits in its aggregate's part
    orc = Oracle(ex, "multicolor")
    b, xt = s.random_rhs(0)
    x, info = orc.pcg(b, tolerance=1e-8)
    assert info["residuals"][-1] <= 1e-8 and info["niters"] < 40


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["bidomain2d", "emi3d", "emi3d_strips", "bidomain3d_sa"])
def test_block_layout_on_one_gpu_matches_oracle(case, monkeypatch):
    """world == 1 on a partitioned hierarchy: the (part, colour) row layout and per-block launches must
    reproduce the oracle on the same hierarchy (the multi-GPU arithmetic without NCCL)."""
    monkeypatch.setenv("MAMG_DIST_MIN_ROWS", "200")
    if case == "bidomain2d":
        s, prm = problems.bidomain_system(2, 32, gamma=1e3), params.parameters_metric_schwarz
    elif case.startswith("emi3d"):
        s, prm = problems.emi_system(3, 12, gamma=1e6), params.default_metric_parameters
    else:
        from metric_amg_examples_b200 import haznics_compat as hz
        s = problems.bidomain_system(3, 8, gamma=10.0)
        prm = dict(params.parameters_standard, AMG_type=hz.SA_AMG, cycle_type=hz.V_CYCLE, coarse_dof=40, max_aggregation=8)
    part = problems.slab_partition(s, 4, axis=0 if case == "emi3d_strips" else None)
    H = mamg.Hierarchy(s.A, prm, s.interface_dofs if "standard" not in str(prm.get("aggregation_type")) and case != "bidomain3d_sa" else None, part=part)
    H.to_device(0)
    H.dist_init(0, 1)
    orc = Oracle(H.export(), "multicolor")
    r = np.random.default_rng(0).standard_normal(s.ndofs)
    z, zo = H.apply(r), orc.apply(r)
    assert np.linalg.norm(z - zo) / np.linalg.norm(zo) < 1e-10
    b, xt = s.random_rhs(1)
    x, info = H.pcg(b, tolerance=1e-8, relative=True)
    _, ref = orc.pcg(b, tolerance=1e-8, relative=True)
    assert abs(info["niters"] - ref["niters"]) <= 1
