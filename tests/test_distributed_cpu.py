"""Host-side logic of the sharded run, on CPU: the re-partitioning plan, and the
communicator plumbing over a world_size-2 gloo process group."""
import os
import socket

import numpy as np
import pytest

from particles_b200.distributed import LocalComm, partition_plan


def _brute(n_local_all, rank, bnd):
    G = len(n_local_all)
    n = sum(n_local_all)
    owner_old = np.repeat(np.arange(G), n_local_all)
    owner_new = np.searchsorted(np.asarray(bnd[1:]), np.arange(n), side="right")
    send = np.array([np.sum((owner_old == rank) & (owner_new == d)) if d != rank else 0 for d in range(G)])
    recv = np.array([np.sum((owner_old == s) & (owner_new == rank)) if s != rank else 0 for s in range(G)])
    return send, recv, int(np.sum(owner_new == rank))


def test_ownership_rule():
    from particles_b200.distributed import ownership
    # balanced shares: boundaries move to the balanced positions, only neighbours exchange
    bnd, first, q, mode = ownership([100, 103, 98, 99], stay_cap=150, window=10)
    assert mode == 2 and bnd == [0, 100, 200, 300, 400]
    # the window limits the shift per step
    bnd, first, q, mode = ownership([100, 140, 80, 80], stay_cap=150, window=10)
    assert mode == 2 and bnd == [0, 100, 230, 310, 400]
    # a share above the capacity: equal blocks through the host
    bnd, first, q, mode = ownership([100, 180, 60, 60], stay_cap=150, window=10)
    assert mode == 1 and bnd == [0, 100, 200, 300, 400]
    # everything on rank 0 (start of a run): the first boundary gives away one window per step
    bnd, first, q, mode = ownership([52, 0, 0, 0], stay_cap=60, window=4)
    assert mode == 2 and bnd[0] == 0 and bnd[1] == 48 and bnd[-1] == 52


def test_partition_plan_matches_brute_force():
    from particles_b200.distributed import ownership
    rng = np.random.default_rng(0)
    cases = [[1, 0], [0, 0, 0, 0], [52, 0, 0, 0], [100, 103, 98, 99], [5, 0, 17, 1, 0, 0, 3, 9]]
    cases += [list(rng.integers(0, 50, int(g))) for g in rng.integers(1, 9, 40)]
    for case in cases:
        G = len(case)
        for cap, win in ((10 ** 9, 5), (30, 3), (0, 0)):
            bnd, first, q, mode = ownership(case, cap, win)
            assert bnd[0] == 0 and bnd[-1] == sum(case) and all(bnd[r] <= bnd[r + 1] for r in range(G))
            tot_send = np.zeros((G, G), dtype=np.int64)
            for r in range(G):
                send, recv, q2, mine, mode2 = partition_plan(case, r, cap, win)
                bs, br, bm = _brute(case, r, bnd)
                assert np.array_equal(send, bs) and np.array_equal(recv, br) and mine == bm and mode2 == mode
                if mode == 2:        # neighbour mode: adjacent ranks only, at most `window` particles per message
                    assert all(send[d] == 0 for d in range(G) if abs(d - r) > 1) and send.max(initial=0) <= win
                tot_send[r] = send
            for r in range(G):
                assert np.array_equal(tot_send[:, r], partition_plan(case, r, cap, win)[1])


def test_local_comm_all_to_all_matches_plan():
    import torch
    rows, case = 3, [7, 0, 5, 2]
    G = len(case)
    comm = LocalComm(G)
    first = np.concatenate([[0], np.cumsum(case)])
    plans = [partition_plan(case, r) for r in range(G)]        # capacity 0: equal blocks
    sends, recvs = [], []
    for r in range(G):
        send, recv, q, mine, _ = plans[r]
        buf = []
        for d in range(G):                   # packed [row][k] per destination segment
            lo = max(first[r], d * q)
            seg = np.arange(lo, lo + send[d])
            buf.append(np.stack([seg * 10 + row for row in range(rows)]).reshape(-1))
        sends.append(torch.tensor(np.concatenate(buf) if buf else np.zeros(0), dtype=torch.float64))
        recvs.append(torch.full((rows * 32,), -1.0, dtype=torch.float64))
    comm.all_to_all_v(sends, [p[0] for p in plans], recvs, [p[1] for p in plans], rows)
    for r in range(G):
        send, recv, q, mine, _ = plans[r]
        o = 0
        for s in range(G):
            n = int(recv[s])
            lo = max(first[s], r * q)
            want = np.stack([np.arange(lo, lo + n) * 10 + row for row in range(rows)]).reshape(-1) if n else np.zeros(0)
            assert np.array_equal(recvs[r][o:o + n * rows].numpy(), want)
            o += n * rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ok):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from particles_b200.distributed import TorchComm
        comm = TorchComm()
        # all_reduce / all_gather as the step uses them
        sc = torch.tensor([10 + rank, 100 * (rank + 1)], dtype=torch.int64)
        comm.all_reduce([sc[0:1]], "sum")
        comm.all_reduce([sc[1:2]], "max")
        assert sc.tolist() == [21, 200]
        g = comm.all_gather([torch.arange(4, dtype=torch.int64) + 10 * rank])[0]
        assert g.shape == (2, 4) and g[1, 0].item() == 10
        assert comm.max_int([3 + rank], "cpu") == 4
        # exchange plan: rank 0 produced 9 particles, rank 1 produced 1 -> q = 5
        case, rows = [9, 1], 2
        send, recv, q, mine, _ = partition_plan(case, rank)
        first = [0, 9]
        buf = []
        for d in range(world):
            lo = max(first[rank], d * q)
            seg = np.arange(lo, lo + send[d], dtype=np.float64)
            buf.append(np.stack([seg, -seg]).reshape(-1))
        sendt = torch.tensor(np.concatenate(buf), dtype=torch.float64) if sum(send) else torch.zeros(0, dtype=torch.float64)
        recvt = torch.zeros(rows * 16, dtype=torch.float64)
        comm.all_to_all_v([sendt], [send], [recvt], [recv], rows)
        if rank == 1:
            assert recv.tolist() == [4, 0] and mine == 5
            assert recvt[:8].tolist() == [5., 6., 7., 8., -5., -6., -7., -8.]
        else:
            assert recv.tolist() == [0, 0] and mine == 5
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


def test_torch_comm_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    ok = mp.get_context("spawn").Array("i", [0, 0])
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]
