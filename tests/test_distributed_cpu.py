"""world_size-2 tests of the multi-GPU host logic on CPU (gloo): the data-parallel gradient mean and
the sharded sliding-window accumulate + reduce.  The kernels themselves are CUDA-only; here the
per-rank accumulate is the numpy oracle, so the test pins the partition / exchange logic."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _grad_mean(rank, world):
    from e2enet_medical_b200.training import allreduce_mean_grads
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv3d(2, 4, (1, 3, 3)), torch.nn.InstanceNorm3d(4, affine=True),
                              torch.nn.Conv3d(4, 3, 1, bias=False))
    g = torch.Generator().manual_seed(100 + rank)
    for p in net.parameters():
        p.grad = torch.randn(p.shape, generator=g)
    net[2].weight.grad = None if False else net[2].weight.grad        # all grads present
    allreduce_mean_grads(list(net.parameters()), world)
    return [p.grad.clone().numpy() for p in net.parameters()]


def test_dp_gradient_mean_gloo():
    out = _run(_grad_mean)
    # expected mean computed independently
    exp = None
    for rank in range(2):
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Conv3d(2, 4, (1, 3, 3)), torch.nn.InstanceNorm3d(4, affine=True),
                                  torch.nn.Conv3d(4, 3, 1, bias=False))
        g = torch.Generator().manual_seed(100 + rank)
        gs = [torch.randn(p.shape, generator=g).numpy() for p in net.parameters()]
        exp = gs if exp is None else [a + b for a, b in zip(exp, gs)]
    exp = [a / 2 for a in exp]
    for r in range(2):
        for a, b in zip(out[r], exp):
            np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-7)
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b), "replicas must hold identical gradients after the all-reduce"


def _fake_net(x):
    # deterministic 3-class 'network': logits from simple voxel functions
    return np.stack([x[0], -x[0], 0.5 * x[0] * x[0]], 0).astype(np.float32)


def _sharded_window_root(rank, world):
    return _sharded_window(rank, world, dst=0)


def _sharded_window(rank, world, dst=None):
    from e2enet_medical_b200.network_architecture.neural_network import SegmentationNetwork
    from oracle import window as owin
    rs = np.random.RandomState(3)
    vol = rs.randn(1, 20, 30, 26).astype(np.float32)
    patch = (8, 16, 12)
    steps = owin.compute_steps(patch, vol.shape[1:], 0.5)
    gauss = owin.gaussian_map(patch)
    tiles = [(a, b, c) for a in steps[0] for b in steps[1] for c in steps[2]]
    mine = SegmentationNetwork._shard_tiles(tiles, rank, world)
    agg = np.zeros((3,) + vol.shape[1:], np.float32)
    wsum = np.zeros(vol.shape[1:], np.float32)
    for (a, b, c) in mine:
        t = vol[:, a:a + patch[0], b:b + patch[1], c:c + patch[2]]
        pr = owin.softmax0(_fake_net(t)) * gauss
        agg[:, a:a + patch[0], b:b + patch[1], c:c + patch[2]] += pr
        wsum[a:a + patch[0], b:b + patch[1], c:c + patch[2]] += gauss
    ta, tw = torch.from_numpy(agg), torch.from_numpy(wsum)
    SegmentationNetwork._reduce_accumulators(ta, tw, None, dst)     # all-reduce, or reduce to rank dst
    if dst is not None and rank != dst:
        return None, [tuple(t) for t in mine], len(tiles)
    return (ta / tw).numpy(), [tuple(t) for t in mine], len(tiles)


def test_sharded_sliding_window_gloo():
    from oracle import window as owin
    out = _run(_sharded_window)
    (p0, t0, n), (p1, t1, _) = out
    assert len(t0) + len(t1) == n and not set(t0) & set(t1), "tiles must be partitioned exactly once"
    assert np.array_equal(p0, p1), "every rank returns the full result"
    # single-process oracle on the same volume
    rs = np.random.RandomState(3)
    vol = rs.randn(1, 20, 30, 26).astype(np.float32)
    seg, probs = owin.predict_tiled(lambda t: _fake_net(t), vol, 3, (8, 16, 12), 0.5, do_mirroring=False,
                                    use_gaussian=True)[:2]
    np.testing.assert_allclose(p0, probs, rtol=2e-6, atol=1e-7)
    assert (p0.argmax(0) == seg).mean() > 0.9999


def test_sharded_sliding_window_reduce_to_root_gloo():
    from oracle import window as owin
    (p0, t0, n), (p1, t1, _) = _run(_sharded_window_root)
    assert p1 is None and len(t0) + len(t1) == n
    rs = np.random.RandomState(3)
    vol = rs.randn(1, 20, 30, 26).astype(np.float32)
    _, probs = owin.predict_tiled(lambda t: _fake_net(t), vol, 3, (8, 16, 12), 0.5, do_mirroring=False,
                                  use_gaussian=True)[:2]
    np.testing.assert_allclose(p0, probs, rtol=2e-6, atol=1e-7)


def _slab_window(rank, world):
    """slab ownership (SURVEY 8(e) option B): contiguous tile ranges, local accumulators over the own
    x-extent only, neighbour exchange of the overlap planes"""
    from e2enet_medical_b200.network_architecture.neural_network import SegmentationNetwork as S
    from oracle import window as owin
    rs = np.random.RandomState(3)
    vol = rs.randn(1, 40, 30, 26).astype(np.float32)
    patch = (8, 16, 12)
    steps = owin.compute_steps(patch, vol.shape[1:], 0.5)
    gauss = owin.gaussian_map(patch)
    tiles = [(a, b, c) for a in steps[0] for b in steps[1] for c in steps[2]]
    cut, bx, hi = S._slab_plan(tiles, patch[0], vol.shape[1], world)
    mine = tiles[cut[rank]:cut[rank + 1]]
    x0, ext = bx[rank], max(hi[rank], bx[rank + 1]) - bx[rank]
    agg = np.zeros((3, ext) + vol.shape[2:], np.float32)
    wsum = np.zeros((ext,) + vol.shape[2:], np.float32)
    for (a, b, c) in mine:
        t = vol[:, a:a + patch[0], b:b + patch[1], c:c + patch[2]]
        pr = owin.softmax0(_fake_net(t)) * gauss
        agg[:, a - x0:a - x0 + patch[0], b:b + patch[1], c:c + patch[2]] += pr
        wsum[a - x0:a - x0 + patch[0], b:b + patch[1], c:c + patch[2]] += gauss
    ta, tw = torch.from_numpy(agg), torch.from_numpy(wsum)
    S._slab_exchange(ta, tw, rank, bx, hi)
    own = bx[rank + 1] - bx[rank]
    return (ta[:, :own] / tw[:own]).numpy(), (bx[rank], bx[rank + 1]), len(mine), len(tiles)


def test_slab_sharded_sliding_window_gloo():
    from oracle import window as owin
    for world in (2, 3):
        out = _run(_slab_window, world)
        assert sum(o[2] for o in out) == out[0][3], "every tile predicted exactly once"
        rs = np.random.RandomState(3)
        vol = rs.randn(1, 40, 30, 26).astype(np.float32)
        _, probs = owin.predict_tiled(lambda t: _fake_net(t), vol, 3, (8, 16, 12), 0.5, do_mirroring=False,
                                      use_gaussian=True)[:2]
        covered = 0
        for p_, (lo, up), _, _ in out:
            assert lo == covered, "owned slabs tile the x axis"
            covered = up
            np.testing.assert_allclose(p_, probs[:, lo:up], rtol=2e-6, atol=1e-7)
        assert covered == vol.shape[1]
