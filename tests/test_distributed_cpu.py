"""Host-side logic of the sharded run, on CPU: the ownership rule the device applies at every step
(equal blocks), and the one-time plumbing a multi-process host program owes the library (the
all-gather of the CUDA-IPC blobs and the host gathers of the read-outs) over a world_size-2 gloo
process group.  The data path itself needs GPUs (tests/test_gpu_sharded.py, tests/test_gpu_multiproc.py)."""
import os
import socket

import numpy as np

from particles_b200.distributed import LocalGroup, equal_blocks, owner_of


def test_equal_blocks_ownership_rule():
    # k_rank_offsets: q = ceil(n / G), rank r owns [r q, (r+1) q) clipped to n
    assert equal_blocks(400, 4) == [0, 100, 200, 300, 400]
    assert equal_blocks(401, 4) == [0, 101, 202, 303, 401]
    assert equal_blocks(3, 4) == [0, 1, 2, 3, 3]              # fewer particles than ranks: trailing ranks are empty
    assert equal_blocks(1, 8) == [0, 1, 1, 1, 1, 1, 1, 1, 1]  # the filter's first particle lives on rank 0
    assert equal_blocks(0, 2) == [0, 0, 0]
    rng = np.random.default_rng(0)
    for _ in range(200):
        G = int(rng.integers(1, 17))
        n = int(rng.integers(0, 5000))
        b = equal_blocks(n, G)
        assert b[0] == 0 and b[-1] == n and all(b[r] <= b[r + 1] for r in range(G))
        sizes = np.diff(b)
        assert sizes.max(initial=0) <= -(-max(n, 1) // G)
        if n:
            g = np.arange(n)
            own = owner_of(g, n, G)
            assert np.array_equal(own, np.searchsorted(np.asarray(b[1:]), g, side="right"))


def test_local_group_gathers_in_rank_order():
    grp = LocalGroup(3)
    assert grp.size == 3 and grp.local_ranks == [0, 1, 2]
    parts = [np.arange(2), np.arange(3) + 10, np.arange(1) + 20]
    assert np.concatenate(grp.gather_host(parts)).tolist() == [0, 1, 10, 11, 12, 20]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ok):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from particles_b200._lib import PF_MG_IPC_BYTES, PF_MG_NARRAYS
        from particles_b200.distributed import TorchGroup
        grp = TorchGroup()
        assert grp.size == world and grp.rank == rank and grp.local_ranks == [rank]
        # the IPC blobs: every rank receives every rank's blob, in rank order, bytes intact
        blob = bytes([rank * 16 + (k % 16) for k in range(PF_MG_NARRAYS * PF_MG_IPC_BYTES)])
        blobs = grp.all_gather_blobs(blob)
        assert len(blobs) == world and all(len(b) == PF_MG_NARRAYS * PF_MG_IPC_BYTES for b in blobs)
        assert blobs[0][0] == 0 and blobs[1][0] == 16 and blobs[rank] == blob
        # read-outs: shards concatenate in global (rank-major) order
        part = np.arange(3 + rank) + 100 * rank
        whole = np.concatenate(grp.gather_host([part]))
        assert whole.tolist() == [0, 1, 2, 100, 101, 102, 103]
        grp.barrier()
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


def test_torch_group_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    ok = mp.get_context("spawn").Array("i", [0, 0])
    port = _free_port()
    mp.spawn(_worker, args=(2, port, ok), nprocs=2, join=True)
    assert list(ok) == [1, 1]
