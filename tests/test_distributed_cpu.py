"""Host-side logic of the multi-GPU path on CPU: partition arithmetic, and the exchange steps (row delivery to the
owner only, row all-gather, best-hypothesis reduction, result gather, the shared work counter, raising together)
over a world_size-2 gloo group."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from platymatch_b200 import distributed as PD


def test_shard_rows_cover_and_align():
    for n in (1, 127, 128, 7200, 18000, 20000):
        for world in (1, 2, 4, 8):
            got, per = [], None
            for r in range(world):
                b, e, per = PD.shard_rows(n, r, world)
                assert (b % 4 == 0 or b == e) and per % 128 == 0 and 0 <= b <= e <= n and e - b <= per
                got.extend(range(b, e))
            assert got == list(range(n))


def test_pair_and_hypothesis_partitions():
    pairs = PD.all_pairs(12)
    assert len(pairs) == 66 and pairs[0] == (0, 1) and pairs[-1] == (10, 11)
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            mine = PD.pairs_for_rank(12, r, world)
            assert abs(len(mine) - 66 / world) < 1
            seen += mine
        assert sorted(seen) == pairs
        hyp = sorted(q for r in range(world) for q in PD.hypotheses_for_rank(4, r, world))
        assert hyp == [0, 1, 2, 3]
        for n_hyp in (4, 8):
            owners = [PD.owner_of_hypothesis(q, n_hyp, world) for q in range(n_hyp)]
            assert all(0 <= o < world for o in owners)
            load = [owners.count(r) for r in range(world)]
            assert max(load) == -(-n_hyp // world)              # as even as it gets
            assert all(q in PD.hypotheses_for_rank(n_hyp, owners[q], world) for q in range(n_hyp))
    assert [PD.owner_of_hypothesis(q, 4, 8) for q in range(4)] == [0, 2, 4, 6]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # --- row all-gather: every rank fills its shard of a known matrix
        n_rows, ld = 300, 8
        truth = torch.arange(n_rows * ld, dtype=torch.float32).reshape(n_rows, ld)
        b, e, per = PD.shard_rows(n_rows, rank, world)
        local = torch.zeros((per, ld), dtype=torch.float32)
        local[: e - b] = truth[b:e]
        full = PD.allgather_rows(local, n_rows, per)
        ok_rows = bool(torch.equal(full, truth))
        # --- row delivery to the owner only: each matrix lands on its owner, nobody else holds it
        for owner in range(world):
            got = PD.gather_rows_to_owner(local * (owner + 1), n_rows, per, owner)
            ok_rows &= (got is None) if rank != owner else bool(torch.equal(got, truth * (owner + 1)))
        # --- shared work counter: every index exactly once over all ranks and threads
        import threading
        counter = PD.WorkCounter()
        taken = []

        def pull():
            while True:
                k = counter.next()
                if k >= 40:
                    return
                taken.append(k)
        ts = [threading.Thread(target=pull) for _ in range(3)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        mine_t = torch.zeros(40, dtype=torch.int32)
        mine_t[taken] += 1
        dist.all_reduce(mine_t)
        ok_rows &= bool(torch.all(mine_t == 1)) and len(set(taken)) == len(taken)
        # --- raise together: rank 1 fails, BOTH raise (nobody is left waiting in the next collective)
        try:
            PD._raise_together(ValueError("boom") if rank == 1 else None)
            ok_rows = False
        except ValueError:
            ok_rows &= rank == 1
        except RuntimeError:
            ok_rows &= rank == 0
        PD._raise_together(None)
        # --- best-hypothesis reduction: ties resolve to the first hypothesis (np.argmax)
        inl_all = [17, 42, 42, 5]
        mine = PD.hypotheses_for_rank(4, rank, world)
        inl = [torch.tensor(inl_all[h], dtype=torch.int32) for h in mine]
        mats = torch.stack([torch.full((16,), float(h), dtype=torch.float64) for h in mine])
        inliers, transforms, best = PD.reduce_best_hypothesis(inl, mats, mine, 4)
        ok_best = best == 1 and inliers.tolist() == inl_all and all(float(transforms[h, 0]) == h for h in range(4))
        # --- result gather
        pairs = PD.all_pairs(5)
        local_res = {k: np.full((4, 4), float(k)) for k in range(len(pairs)) if k % world == rank}
        out = PD.gather_results(local_res, len(pairs), None)
        ok_gather = all(float(out[k, 0]) == k for k in range(len(pairs)))
        q.put((rank, ok_rows, ok_best, ok_gather))
    finally:
        dist.destroy_process_group()


def test_exchange_steps_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_rows, ok_best, ok_gather in results:
        assert ok_rows and ok_best and ok_gather, (rank, ok_rows, ok_best, ok_gather)


def test_single_process_degenerates():
    local = torch.ones((128, 4))
    assert PD.allgather_rows(local, 100, 128).shape == (100, 4)
    assert PD.gather_rows_to_owner(local, 100, 128, 0).shape == (100, 4)
    c = PD.WorkCounter()
    assert [c.next() for _ in range(3)] == [0, 1, 2]
    with pytest.raises(ValueError):
        PD._raise_together(ValueError("x"))
    inl, tr, best = PD.reduce_best_hypothesis([torch.tensor(3), torch.tensor(9)], torch.zeros((2, 16), dtype=torch.float64),
                                              [0, 1], 2)
    assert best == 1 and inl.tolist() == [3, 9]
