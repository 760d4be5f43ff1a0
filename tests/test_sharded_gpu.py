"""GPU test (-m gpu) for the batched offline extraction of config 5 (thousands of frames, sharded frame-wise): a frame's
features do not depend on its position in a batch, on the chunking, or on which shard it falls into.  Size-independent
property: 1200 frames built from 8 distinct images; every copy must reproduce its source bit for bit (ORB keypoints and
descriptors, line keylines and LBD descriptors), and the checksum of checksums is the same for 1, 2, 4 and 8 shards."""
import importlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PKG = "orb_slam2_modification_with-point-and-line-feature_b200"


def test_sharded_batched_extraction_is_position_independent(api, synth, oracle):
    sh = importlib.import_module(PKG + ".sharding")
    base = synth.frames(7000, 8)
    n = 1200
    src = (np.arange(n) * 5 + (np.arange(n) // 7)) % 8
    frames = base[src]
    orb = api.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=300)
    line = api.LineExtractor(max_batch=300)
    # reference results of the 8 distinct images, one at a time, checked against the oracle for one of them
    single_orb, single_line = [], []
    for i in range(8):
        single_orb.append(orb(base[i]))
        single_line.append(line.ExtractLineSegment(base[i]))
    ok, od = oracle.OrbOracle(1000).extract(base[3])
    assert np.array_equal(single_orb[3][0], ok) and np.array_equal(single_orb[3][1], od)
    ref_sum = [sh.frame_checksum(*single_orb[i]) for i in range(8)]
    ref_lsum = [sh.frame_checksum(single_line[i][0], single_line[i][1]) for i in range(8)]

    def run(world):
        sums = np.zeros(n, np.int64)
        lsums = np.zeros(n, np.int64)
        for rank in range(world):
            b, e = sh.shard_range(n, world, rank)
            if e == b:
                continue
            kps, desc, cnt = orb.extract_batch(frames[b:e])
            kls, ldesc, _, lcnt = line.extract_batch(frames[b:e])
            for k in range(e - b):
                sums[b + k] = sh.frame_checksum(kps[k, :cnt[k]], desc[k, :cnt[k]])
                lsums[b + k] = sh.frame_checksum(kls[k, :lcnt[k]], ldesc[k, :lcnt[k]])
        return sums, lsums

    expect = np.array([ref_sum[s] for s in src], np.int64)
    lexpect = np.array([ref_lsum[s] for s in src], np.int64)
    totals = set()
    for world in (1, 2, 4, 8):
        sums, lsums = run(world)
        assert np.array_equal(sums, expect), f"ORB features depend on the batch position (world={world})"
        assert np.array_equal(lsums, lexpect), f"line features depend on the batch position (world={world})"
        totals.add((int(sums.sum() & 0x7FFFFFFFFFFFFFFF), int(lsums.sum() & 0x7FFFFFFFFFFFFFFF)))
    assert len(totals) == 1
