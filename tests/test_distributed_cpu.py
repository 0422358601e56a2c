"""Host-side logic of the sharded run, on CPU: the re-partitioning plan, and the
communicator plumbing over a world_size-2 gloo process group."""
import os
import socket

import numpy as np
import pytest

from particles_b200.distributed import LocalComm, partition_plan


def _brute(n_local_all, rank):
    G = len(n_local_all)
    n = sum(n_local_all)
    q = max(-(-n // G), 1)
    owner_old = np.repeat(np.arange(G), n_local_all)
    owner_new = np.arange(n) // q
    send = np.array([np.sum((owner_old == rank) & (owner_new == d)) if d != rank else 0 for d in range(G)])
    recv = np.array([np.sum((owner_old == s) & (owner_new == rank)) if s != rank else 0 for s in range(G)])
    return send, recv, q, int(np.sum(owner_new == rank))


def test_partition_plan_stay_rule():
    # balanced shares that fit the capacity: nothing moves, ownership follows the parents
    send, recv, q, mine = partition_plan([100, 103, 98, 99], 1, stay_cap=120)
    assert send.sum() == 0 and recv.sum() == 0 and mine == 103
    # a share above the capacity, or too unbalanced: equal blocks
    for case, cap in (([100, 130, 98, 99], 120), ([10, 200, 200, 200], 1000)):     # above capacity / below q/2
        tot = [partition_plan(case, r, stay_cap=cap) for r in range(4)]
        assert sum(int(t[0].sum()) for t in tot) > 0
        assert [t[3] for t in tot] == [t[2] for t in tot][:3] + [sum(case) - 3 * tot[0][2]]


def test_partition_plan_matches_brute_force():
    rng = np.random.default_rng(0)
    cases = [[1, 0], [0, 0, 0, 0], [52, 0, 0, 0], [100, 103, 98, 99], [5, 0, 17, 1, 0, 0, 3, 9]]
    cases += [list(rng.integers(0, 50, int(g))) for g in rng.integers(1, 9, 40)]
    for case in cases:
        G = len(case)
        tot_send = np.zeros((G, G), dtype=np.int64)
        for r in range(G):
            send, recv, q, mine = partition_plan(case, r)
            bs, br, bq, bm = _brute(case, r)
            assert np.array_equal(send, bs) and np.array_equal(recv, br) and q == bq and mine == bm
            tot_send[r] = send
        for r in range(G):
            assert np.array_equal(tot_send[:, r], partition_plan(case, r)[1])       # what r receives is what the others send


def test_local_comm_all_to_all_matches_plan():
    import torch
    rows, case = 3, [7, 0, 5, 2]
    G = len(case)
    comm = LocalComm(G)
    first = np.concatenate([[0], np.cumsum(case)])
    plans = [partition_plan(case, r) for r in range(G)]
    sends, recvs = [], []
    for r in range(G):
        send, recv, q, mine = plans[r]
        buf = []
        for d in range(G):                   # packed [row][k] per destination segment
            lo = max(first[r], d * q)
            seg = np.arange(lo, lo + send[d])
            buf.append(np.stack([seg * 10 + row for row in range(rows)]).reshape(-1))
        sends.append(torch.tensor(np.concatenate(buf) if buf else np.zeros(0), dtype=torch.float64))
        recvs.append(torch.full((rows * 32,), -1.0, dtype=torch.float64))
    comm.all_to_all_v(sends, [p[0] for p in plans], recvs, [p[1] for p in plans], rows)
    for r in range(G):
        send, recv, q, mine = plans[r]
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
        send, recv, q, mine = partition_plan(case, rank)
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
