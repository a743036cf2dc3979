"""F16 - the weighted partitioner (the reference's distribute_weights_swapping, MPMP.jl:425-465): the library's C++
implementation (clrsdp_partition; no GPU needed) against the line-by-line Python restatement in oracle/partition.py, plus
the properties the reference states for it (:408-415): cardinalities differ by at most one, disjoint cover, and the
largest set weight is no worse than that of the contiguous split it starts from."""
import random

import numpy as np
import pytest

from clrsdp import capi, instances, solver
from oracle.partition import distribute_weights_swapping


def _sets(set_of, parts):
    return [sorted(int(i) for i in np.nonzero(set_of == p)[0]) for p in range(parts)]


@pytest.mark.parametrize("seed", range(12))
def test_partition_equals_the_reference_restatement(seed):
    rng = random.Random(seed)
    n = rng.randint(1, 60)
    parts = rng.randint(1, min(9, n))
    kind = seed % 3
    if kind == 0:
        w = [float(rng.randint(1, 50)) ** 3 for _ in range(n)]         # nb^3 weights like the reference's (:494-495)
    elif kind == 1:
        w = [float(rng.choice([1, 1, 81, 243, 3])) for _ in range(n)]  # many ties, a few heavy items
    else:
        w = [rng.uniform(0.0, 1.0) for _ in range(n)]
    set_of, mx = capi.partition(w, parts)
    ref_sets, ref_w, _ = distribute_weights_swapping(w, parts)
    assert _sets(set_of, parts) == [sorted(s) for s in ref_sets]
    assert mx == pytest.approx(max(ref_w), rel=1e-12)


@pytest.mark.parametrize("seed", range(8))
def test_partition_properties(seed):
    rng = random.Random(100 + seed)
    n, parts = rng.randint(5, 200), rng.randint(2, 8)
    w = [float(rng.randint(1, 128)) ** 3 for _ in range(n)]
    set_of, mx = capi.partition(w, parts)
    sets = _sets(set_of, parts)
    assert sorted(sum(sets, [])) == list(range(n))                       # disjoint cover
    sizes = [len(s) for s in sets]
    assert max(sizes) - min(sizes) <= 1                                  # (:411) cardinalities differ by at most 1
    step = n // parts + 1
    nstep = parts - (step * parts - n)
    bounds = np.cumsum([0] + [step] * nstep + [step - 1] * (parts - nstep))
    contiguous = max(sum(w[bounds[i]:bounds[i + 1]]) for i in range(parts))
    assert mx <= contiguous * (1 + 1e-12)                                # swaps never raise the maximum
    assert mx == pytest.approx(max(sum(w[i] for i in s) for s in sets))


def test_sphere_packing_clusters_are_spread_by_weight():
    """The instance VERDICT r1 names: sphere packing has clusters of very different size (dim_S = 2 x big + 3 x medium +
    2 x tiny); the partition of its cluster weights over two parts must beat the contiguous split."""
    cons, b, _ = instances.sphere_packing_2point(n=3, d=12, prec=256)
    bi = solver.get_block_info(cons)
    w = [capi.cluster_weight(bi.m[j], bi.L[j], bi.n_samples[j], bi.delta[j], bi.n_y) for j in range(bi.J)]
    owner, mx = solver.partition_clusters(bi, 2)
    assert set(owner) == {0, 1}
    assert abs(int((owner == 0).sum()) - int((owner == 1).sum())) <= 1
    contiguous = max(sum(w[:4]), sum(w[4:]))
    assert mx <= contiguous
    assert mx == pytest.approx(max(sum(wi for wi, o in zip(w, owner) if o == p) for p in (0, 1)))
    # the formula of SURVEY §8e
    j = 1
    dimS, nbs = bi.dim_S[j], bi.Y_blocksizes[j]
    assert w[j] == pytest.approx(40 * sum(nb ** 3 for nb in nbs) + dimS ** 3 / 3 + dimS ** 2 * bi.n_y + bi.n_y ** 2 * dimS)
