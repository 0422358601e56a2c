"""The REAL multi-process path: one process per GPU (torch.multiprocessing.spawn), CUDA-IPC peer
mappings, kernels that spin on flags written by other GPUs over NVLink.  Asserts sharded ==
unsharded BITWISE (log-weights, K, ancestors, assignments, log-evidence) on every rank.  Needs
>= 2 devices (skipped otherwise; run on the box with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ndev():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _mix2d(rng, T, k=8, radius=6.0):
    ang = 2 * np.pi * rng.integers(0, k, T) / k
    return radius * np.stack([np.cos(ang), np.sin(ang)], axis=1) + rng.standard_normal((T, 2))


def _worker(rank, world, port, N, T, margin, ok):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["PF_MG_TIMEOUT_MS"] = "30000"
    if margin:
        os.environ["PF_SEL_MARGIN"] = str(margin)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import particles_b200 as pb
        from particles_b200.distributed import ShardedFearnheadParticles, TorchGroup
        rng = np.random.default_rng(1234)
        ys, us = _mix2d(rng, T), rng.random(T)
        prior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
        crp = pb.ChineseRestaurantProcess(1.0)
        sh = ShardedFearnheadParticles(N, prior, crp, TorchGroup(), history=T, k_max=24, device=rank)
        half = T // 2
        sh.filter_(ys[:half], us[:half])                 # two calls: the flags' sequence numbers carry over
        le0 = sh.log_evidence.copy()
        sh.filter_(ys[half:], us[half:])
        le = np.concatenate([le0, sh.log_evidence])
        lw, K, anc, asg = sh.logweights(), sh.ncomponents(), sh.last_ancestors(), sh.last_assignments()
        n = len(sh)
        # every rank checks against its own unsharded run on its own GPU (margin does not apply: default bracket)
        os.environ.pop("PF_SEL_MARGIN", None)
        ref = pb.FearnheadParticles(N, prior, crp, history=T, k_max=24, device=rank)
        ref.filter_(ys, uniforms=us)
        assert n == len(ref)
        assert np.array_equal(K, ref.ncomponents())
        assert np.array_equal(anc, ref.last_ancestors())
        assert np.array_equal(asg, ref.last_assignments())
        if N <= (1 << 17):                               # T x N labels through the peers' genealogy arrays (CUDA IPC)
            assert np.array_equal(sh.assignments(), ref.assignments())
        if margin:
            np.testing.assert_allclose(lw, ref.logweights(), rtol=0, atol=1e-12)
            np.testing.assert_allclose(le, ref.log_evidence, rtol=1e-13)
        else:
            assert np.array_equal(lw, ref.logweights())
            assert np.array_equal(le, ref.log_evidence)
        dist.barrier()
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ndev() < 2, reason="needs >= 2 CUDA devices")
@pytest.mark.parametrize("world,N,T,margin", [(2, 1 << 17, 40, 0), (2, 1 << 20, 36, 0), (2, 1 << 17, 30, 1)])
def test_one_process_per_gpu_equals_unsharded(world, N, T, margin):
    import torch.multiprocessing as mp
    world = min(world, _ndev())
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_worker, args=(world, _free_port(), N, T, margin, ok), nprocs=world, join=True)
    assert list(ok) == [1] * world


@pytest.mark.skipif(_ndev() < 4, reason="needs >= 4 CUDA devices")
def test_four_processes():
    import torch.multiprocessing as mp
    ok = mp.get_context("spawn").Array("i", [0] * 4)
    mp.spawn(_worker, args=(4, _free_port(), 1 << 19, 36, 0, ok), nprocs=4, join=True)
    assert list(ok) == [1] * 4


@pytest.mark.skipif(_ndev() < 2, reason="needs >= 2 CUDA devices")
def test_one_process_driving_all_gpus():
    """pf_mg_create_local: one host thread, one shard per device, peer access instead of IPC."""
    import particles_b200 as pb
    from particles_b200.distributed import LocalGroup, ShardedFearnheadParticles
    G = min(_ndev(), 8)
    rng = np.random.default_rng(99)
    N, T = 1 << 18, 36
    ys, us = _mix2d(rng, T), rng.random(T)
    prior = pb.NormalInverseWishart(np.zeros(2), 0.1, np.eye(2), 3.0)
    crp = pb.ChineseRestaurantProcess(1.0)
    sh = ShardedFearnheadParticles(N, prior, crp, LocalGroup(G, devices=list(range(G))), history=T, k_max=24)
    sh.filter_(ys, us)
    ref = pb.FearnheadParticles(N, prior, crp, history=T, k_max=24)
    ref.filter_(ys, uniforms=us)
    assert len(sh) == len(ref)
    assert np.array_equal(sh.logweights(), ref.logweights())
    assert np.array_equal(sh.ncomponents(), ref.ncomponents())
    assert np.array_equal(sh.last_ancestors(), ref.last_ancestors())
    assert np.array_equal(sh.log_evidence, ref.log_evidence)
    assert np.array_equal(sh.assignments(), ref.assignments())


def _worker_chenliu(rank, world, port, N, T, d, ok):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["PF_MG_TIMEOUT_MS"] = "30000"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import particles_b200 as pb
        from particles_b200.distributed import ShardedChenLiuParticles, TorchGroup
        rng = np.random.default_rng(4321)
        if d == 2:
            ys = _mix2d(rng, T)
        else:
            ys = 6.0 * np.sign(rng.standard_normal((32, d)))[rng.integers(0, 32, T)] + rng.standard_normal((T, d))
        prior = pb.NormalInverseWishart(np.zeros(d), 0.1, np.eye(d), d + 1.0)
        crp = pb.ChineseRestaurantProcess(1.0)
        sh = ShardedChenLiuParticles(N, prior, crp, TorchGroup(), history=T, k_max=40, device=rank, seed=7)
        half = T // 2
        sh.filter_(ys[:half])                            # device uniforms, two calls
        le0 = sh.log_evidence.copy()
        sh.filter_(ys[half:])
        le = np.concatenate([le0, sh.log_evidence])
        ref = pb.ChenLiuParticles(N, prior, crp, history=T, k_max=40, device=rank, seed=7)
        ref.filter_(ys)
        assert np.array_equal(sh.ncomponents(), ref.ncomponents())
        assert np.array_equal(sh.last_ancestors(), ref.last_ancestors())
        assert np.array_equal(sh.logweights(), ref.logweights())
        assert np.array_equal(le, ref.log_evidence)
        assert np.array_equal(sh.assignments(), ref.assignments())
        dist.barrier()
        ok[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(_ndev() < 2, reason="needs >= 2 CUDA devices")
@pytest.mark.parametrize("N,T,d", [(1 << 17, 40, 2), (1 << 14, 40, 8)])
def test_chenliu_one_process_per_gpu_equals_unsharded(N, T, d):
    """Sharded Chen-Liu over CUDA IPC: scalar exchanges through peer memory, ancestors pulled from the peers' stores on
    rejuvenation steps; bitwise equal to the one-GPU run."""
    import torch.multiprocessing as mp
    world = min(4, _ndev())
    ok = mp.get_context("spawn").Array("i", [0] * world)
    mp.spawn(_worker_chenliu, args=(world, _free_port(), N, T, d, ok), nprocs=world, join=True)
    assert list(ok) == [1] * world
