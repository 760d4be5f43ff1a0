"""CPU test of the N>1 path: frame-wise sharding + the final gather over torch.distributed (gloo, world_size 2).
The per-frame work is done by the CPU oracle here (no GPU in this test); the sharding / gather code is the one the
multi-GPU tools use with NCCL."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "orb_slam2_modification_with-point-and-line-feature_b200"


def _worker(rank, world, port, n_frames, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module(PKG)
    sh = importlib.import_module(PKG + ".sharding")
    import pyoracle
    b, e = sh.shard_range(n_frames, world, rank)
    o = pyoracle.OrbOracle(300)
    counts, sums = [], []
    for i in range(b, e):
        k, d = o.extract(pkg.synth.frame(6000 + i, 320, 240))
        counts.append(len(k))
        sums.append(sh.frame_checksum(k, d))
    res = sh.gather_results(counts, sums, n_frames, world, rank, dist)
    if rank == 0:
        q.put((res[0].tolist(), res[1].tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges():
    sh = importlib.import_module(PKG + ".sharding")
    for n in (0, 1, 7, 300, 4096):
        for w in (1, 2, 3, 4, 8):
            r = [sh.shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def test_gloo_world2_matches_serial(oracle, synth):
    import torch.multiprocessing as mp
    sh = importlib.import_module(PKG + ".sharding")
    n = 5  # ragged: 2 + 3 frames
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, sums = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = oracle.OrbOracle(300)
    ref = [o.extract(synth.frame(6000 + i, 320, 240)) for i in range(n)]
    assert counts == [len(k) for k, d in ref]
    assert sums == [sh.frame_checksum(k, d) for k, d in ref]


def _feature_worker(rank, world, port, n_frames, q):
    """The final gather of config 5 (sharding.pack_features + gather_features) over gloo: synthetic padded outputs with ragged counts."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module(PKG + ".sharding")
    full = _synthetic_outputs(n_frames)
    b, e = sh.shard_range(n_frames, world, rank)
    parts = sh.pack_features(*[t[b:e] for t in full])
    got, nbytes = sh.gather_features(parts, world, rank, dist)
    if rank == 0:
        q.put(([g.numpy() for g in got], nbytes, sh.features_checksum(got)))
    dist.barrier()
    dist.destroy_process_group()


def _synthetic_outputs(n_frames, cap=40, L=9):
    import torch
    g = torch.Generator().manual_seed(7)
    n = torch.randint(0, cap + 1, (n_frames,), generator=g, dtype=torch.int32)
    ln = torch.randint(0, L + 1, (n_frames,), generator=g, dtype=torch.int32)
    n[1] = 0                                              # a frame without key points
    kps = torch.rand((n_frames, cap, 7), generator=g)
    desc = torch.randint(0, 256, (n_frames, cap, 32), generator=g, dtype=torch.uint8)
    kls = torch.rand((n_frames, L, 17), generator=g)
    ldesc = torch.randint(0, 256, (n_frames, L, 32), generator=g, dtype=torch.uint8)
    lco = torch.rand((n_frames, L, 3), generator=g, dtype=torch.float64)
    return kps, desc, n, kls, ldesc, lco, ln


@pytest.mark.parametrize("world", [2, 3])
def test_feature_gather_equals_the_unsharded_result(world):
    import torch.multiprocessing as mp
    sh = importlib.import_module(PKG + ".sharding")
    n = 7  # ragged shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_feature_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    got, nbytes, chk = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = sh.pack_features(*_synthetic_outputs(n))       # what one rank alone produces
    assert len(got) == len(want) and all(np.array_equal(g, w.numpy()) for g, w in zip(got, want))
    assert nbytes == sum(w.numel() * w.element_size() for w in want)
    assert chk == sh.features_checksum(want)              # the checksum does not depend on the sharding
    alone, nb1 = sh.gather_features(want, 1, 0, None)
    assert alone is want and nb1 == nbytes
