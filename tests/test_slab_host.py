"""Host side of the slab decomposition (active_particle_jamming_b200/slab.py) on the CPU: column
ownership, particle -> rank, merging per-rank answers, and -- with two gloo processes -- that owned
columns + ONE ghost column per side are enough to reproduce the oracle's neighbour-pair set with every
pair reported by exactly one rank (the rule the device code follows: a rank reports (i, j) for its
owned i and any partner j with id_j > id_i, ghost or not)."""
import os
import sys

import numpy as np
import pytest

from _util import random_system
from active_particle_jamming_b200.slab import grid_b, merge_by_id, merge_pairs, owner_of, slab_columns


@pytest.mark.parametrize("b,n", [(10, 1), (10, 3), (21, 4), (85, 8), (1373, 8), (8, 8)])
def test_slab_columns_partition_the_grid(b, n):
    cols = slab_columns(b, n)
    assert cols[0][0] == 0 and sum(w for _, w in cols) == b
    for r in range(1, n):
        assert cols[r][0] == cols[r - 1][0] + cols[r - 1][1]
    assert min(w for _, w in cols) >= 1 and max(w for _, w in cols) - min(w for _, w in cols) <= 1


def test_owner_of_matches_columns():
    L, n = 120.3, 4
    b = grid_b(L)
    assert b == 21
    x = np.random.default_rng(0).uniform(-L / 2, L / 2, 5000)
    x[:3] = [-L / 2, np.nextafter(L / 2, 0), 0.0]
    own = owner_of(x, L, b, n)
    col = np.clip(np.floor((x + L / 2) / (L / b)).astype(int), 0, b - 1)
    for r, (c0, w) in enumerate(slab_columns(b, n)):
        assert np.array_equal(own == r, (col >= c0) & (col < c0 + w))


def test_merge_by_id_detects_loss_and_duplication():
    ids = [np.array([2, 0], dtype=np.int32), np.array([1, 3], dtype=np.int32)]
    parts = [(ids[0], {"x": np.array([20., 0.])}), (ids[1], {"x": np.array([10., 30.])})]
    assert np.array_equal(merge_by_id(parts, 4, ["x"])["x"], [0., 10., 20., 30.])
    with pytest.raises(ValueError):
        merge_by_id(parts[:1], 4, ["x"])
    with pytest.raises(ValueError):
        merge_by_id(parts + parts[:1], 4, ["x"])
    p = merge_pairs([np.array([[3, 5], [0, 9]]), np.array([[0, 2]])])
    assert p.tolist() == [[0, 2], [0, 9], [3, 5]]


def _rank_pairs(rank, nranks, x, y, L, rs2):
    """Pairs this rank reports: owned i x (owned + the two ghost columns), d2 < rs2, id_j > id_i."""
    b = grid_b(L)
    col = np.clip(np.floor((x + L / 2) / (L / b)).astype(int), 0, b - 1)
    c0, w = slab_columns(b, nranks)[rank]
    owned = np.nonzero((col >= c0) & (col < c0 + w))[0]
    ghost = np.nonzero((col == (c0 - 1) % b) | (col == (c0 + w) % b))[0]
    cand = np.union1d(owned, ghost)
    out = []
    for i in owned:
        dx = x[cand] - x[i]; dy = y[cand] - y[i]
        dx -= L * np.round(dx / L); dy -= L * np.round(dy / L)
        hit = cand[(dx * dx + dy * dy < rs2) & (cand > i)]
        out += [(int(i), int(j)) for j in hit]
    return np.array(out, dtype=np.int64).reshape(-1, 2)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        N, rho = 1500, 0.9
        R, L, x, y, phi = random_system(N, rho, 11)
        own = owner_of(x, L, grid_b(L), world)
        mine = _rank_pairs(rank, world, x, y, L, 4.2 * 4.2)
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        counts = [None] * world
        dist.all_gather_object(counts, int((own == rank).sum()))
        # state hand-over: every rank contributes its owned particles, all receive the whole box
        st = [None] * world
        ids = np.nonzero(own == rank)[0].astype(np.int32)
        dist.all_gather_object(st, (ids, {"x": x[ids], "y": y[ids]}))
        merged = merge_by_id(st, N, ["x", "y"])
        if rank == 0:
            q.put((merge_pairs(parts), sum(counts), bool(np.array_equal(merged["x"], x) and np.array_equal(merged["y"], y))))
    finally:
        dist.destroy_process_group()


def test_two_gloo_ranks_reproduce_the_oracle_pair_set():
    import torch.multiprocessing as mp
    from oracle.pyoracle import OracleSim
    N, rho = 1500, 0.9
    R, L, x, y, phi = random_system(N, rho, 11)
    o = OracleSim.from_arrays(R, x, y, phi, rho)
    o.topology(); o.assign(); o.build()
    want = o.pair_set()
    o.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, total, same = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert total == N and same
    assert np.array_equal(got, want)
