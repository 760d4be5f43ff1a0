"""GPU parity tests (-m gpu): LSD + LBD through the C ABI vs the CPU oracle (LSD itself is pinned to cv2 4.13).

Tolerances (north_star): LSD endpoints <= 1e-3 px, KeyLine angles / LBD floats <= 1e-4 relative.  In practice the
integer stages (scaled image, angle map) are bit-exact and so are the segments; the tolerances only absorb the
CUDA-vs-glibc differences of double cos/sin/log/exp/atan2."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [("cfgA_seed1000", (1000, 640, 480)), ("small_seed7", (7, 320, 240)), ("kitti_seed3000", (3000, 1241, 376))]
ENDPOINT_TOL_PX = 1e-3
REL_TOL = 1e-4


@pytest.mark.parametrize("name,args", CASES, ids=[c[0] for c in CASES])
def test_lsd_matches_oracle_and_golden(name, args, synth, oracle, api, golden_dir):
    img = synth.frame(*args)
    ex = api.LineExtractor(max_cols=img.shape[1], max_rows=img.shape[0])
    kls, desc, co = ex.ExtractLineSegment(img)
    assert np.array_equal(ex.scaled_image(), oracle.lsd_scaled(img))          # integer stage: bit-exact
    lines, width, prec, nfa = ex.lsd_segments()
    g = np.load(os.path.join(golden_dir, f"lsd_{name}.npz"))
    assert lines.shape == g["lines"].shape, f"{lines.shape} vs {g['lines'].shape}"
    assert np.abs(lines - g["lines"]).max() <= ENDPOINT_TOL_PX
    assert np.abs(width - g["width"]).max() <= 1e-6 and np.array_equal(prec, g["prec"])
    # NFA: a rectangle's edges pass exactly through its extreme pixels, so the pixels the NFA scan counts are decided by the last
    # bit of cos / sin(theta): the device computes them with the host libm's own arithmetic (csrc/pl_glibc_sincos.cuh), and every
    # NFA value must agree (the remaining tolerance is log / exp / log_gamma rounding of the final formula, not pixel counts)
    same = np.abs(nfa - g["nfa"]) <= 1e-6 * np.maximum(1.0, np.abs(g["nfa"]))
    assert same.all(), f"{(~same).sum()} of {len(same)} NFA values differ"
    # full ExtractLineSegment vs oracle
    okl, odesc, oco = oracle.line_extract(img, 80)
    assert len(kls) == len(okl) == 80
    assert np.array_equal(kls["class_id"], okl["class_id"]) and np.array_equal(kls["num_pixels"], okl["num_pixels"])
    for fld in ("sx", "sy", "ex", "ey", "sx_oct", "sy_oct", "ex_oct", "ey_oct", "pt_x", "pt_y"):
        assert np.abs(kls[fld] - okl[fld]).max() <= ENDPOINT_TOL_PX, fld
    for fld in ("length", "response", "angle", "size"):
        assert np.allclose(kls[fld], okl[fld], rtol=REL_TOL, atol=1e-6), fld
    assert np.allclose(co, oco, rtol=1e-9, atol=1e-12)
    fd = ex.float_descriptors(len(kls))
    _, ofd = oracle.lbd_compute(img, okl)
    assert np.allclose(fd, ofd, rtol=REL_TOL, atol=1e-6)
    assert np.array_equal(desc, odesc)                                         # bits exact given equal floats


def test_few_lines_and_empty(api, oracle):
    img = np.full((240, 320), 90, np.uint8)
    img[:, 160:] = 180
    ex = api.LineExtractor(max_cols=320, max_rows=240)
    kls, desc, co = ex.ExtractLineSegment(img)
    okl, odesc, oco = oracle.line_extract(img, 80)
    assert len(kls) == len(okl) and np.array_equal(desc, odesc)
    assert np.abs(kls["sx"] - okl["sx"]).max() <= ENDPOINT_TOL_PX
    k0 = ex.ExtractLineSegment(np.zeros((0, 0), np.uint8))
    assert len(k0[0]) == 0
    flat = np.full((240, 320), 77, np.uint8)      # no gradient anywhere: zero segments
    kz, dz, cz = ex.ExtractLineSegment(flat)
    assert len(kz) == 0


def test_batch_equals_single(api, synth, oracle):
    frames = synth.frames(6000, 10)
    ex = api.LineExtractor(max_batch=4)
    kls, desc, co, cnt = ex.extract_batch(frames)
    for i in (0, 3, 4, 9):
        okl, odesc, oco = oracle.line_extract(frames[i], 80)
        assert cnt[i] == len(okl)
        assert np.array_equal(kls[i, :cnt[i]]["class_id"], okl["class_id"]), f"frame {i}"
        assert np.abs(kls[i, :cnt[i]]["ex"] - okl["ex"]).max() <= ENDPOINT_TOL_PX
        assert np.array_equal(desc[i, :cnt[i]], odesc), f"frame {i}"
    k2, d2, c2, n2 = ex.extract_batch(frames)
    assert np.array_equal(d2, desc) and np.array_equal(n2, cnt) and np.array_equal(k2, kls)


def test_line_extractor_reads_the_frames_the_orb_extractor_staged(api, synth):
    """One upload per frame (Frame.cc:152-155 hands both extractors the same cv::Mat): pl_orb_staged_images_dev +
    pl_line_extract_batch_from_dev give what pl_line_extract_batch gives from host memory; odd width (the staging pitch is padded)."""
    N = api.N
    for w, h in ((640, 480), (333, 250)):
        frames = synth.frames(6100, 5, w, h)
        orb = api.ORBextractor(500, 1.2, 8, 20, 7, max_cols=w, max_rows=h, max_batch=8)
        cap = orb.max_keypoints()
        kp, dd, nn = np.zeros((5, cap), N.KP_DTYPE), np.zeros((5, cap, 32), np.uint8), np.zeros(5, np.int32)
        with pytest.raises(Exception):
            orb.staged_images()          # nothing staged before the first host-pointer call
        orb.extract_batch_into(frames, kp, dd, nn)
        d, n, r, c, st, fs = orb.staged_images()
        assert (n, r, c) == (5, h, w) and st % 16 == 0 and st >= w and fs == st * h
        ex = api.LineExtractor(max_cols=w, max_rows=h, max_batch=8)
        ref = ex.extract_batch(frames)
        kl, ld, lc, ln = np.zeros((5, 80), N.KL_DTYPE), np.zeros((5, 80, 32), np.uint8), np.zeros((5, 80, 3), np.float64), np.zeros(5, np.int32)
        ex.extract_batch_from_dev_into(d, n, r, c, st, fs, 80, kl, ld, lc, ln)
        assert np.array_equal(ln, ref[3]) and ln.sum() > 50
        for i in range(5):
            assert np.array_equal(kl[i, :ln[i]], ref[0][i, :ln[i]]) and np.array_equal(ld[i, :ln[i]], ref[1][i, :ln[i]])
            assert np.array_equal(lc[i, :ln[i]], ref[2][i, :ln[i]])


@pytest.mark.gpu
def test_orb_call_in_two_halves_with_the_line_extractor_started_in_between(api, synth):
    """pl_orb_stage_batch + pl_orb_extract_staged give what pl_orb_extract_batch gives, and a line extractor whose stream waits
    for the staging copy (pl_orb_stream_wait_staged) and that is called from a second host thread while the ORB half is still
    running reads the same frames (Frame.cc:152-155: one image, two extractor threads)."""
    import threading
    N = api.N
    for w, h in ((640, 480), (333, 250)):
        frames = synth.frames(6200, 6, w, h)
        orb = api.ORBextractor(700, 1.2, 8, 20, 7, max_cols=w, max_rows=h, max_batch=8)
        cap = orb.max_keypoints()
        ref = [np.zeros((6, cap), N.KP_DTYPE), np.zeros((6, cap, 32), np.uint8), np.zeros(6, np.int32)]
        orb.extract_batch_into(frames, *ref)
        ex = api.LineExtractor(max_cols=w, max_rows=h, max_batch=8)
        lref = ex.extract_batch(frames)
        with pytest.raises(Exception):
            orb.extract_staged_into(*ref)        # nothing staged (the one-call entry leaves nothing pending)
        with pytest.raises(Exception):
            orb.stream_wait_staged(ex.stream())
        for _ in range(2):
            kp, dd, nn = np.zeros((6, cap), N.KP_DTYPE), np.zeros((6, cap, 32), np.uint8), np.zeros(6, np.int32)
            kl, ld, lc, ln = np.zeros((6, 80), N.KL_DTYPE), np.zeros((6, 80, 32), np.uint8), np.zeros((6, 80, 3), np.float64), np.zeros(6, np.int32)
            orb.stage_batch(frames)
            orb.stream_wait_staged(ex.stream())
            d, n, r, c, st, fs = orb.staged_images()
            assert (n, r, c) == (6, h, w)
            err = []

            def lines():
                try:
                    ex.extract_batch_from_dev_into(d, n, r, c, st, fs, 80, kl, ld, lc, ln)
                except BaseException as e:
                    err.append(e)
            th = threading.Thread(target=lines)
            th.start()
            orb.extract_staged_into(kp, dd, nn)
            th.join()
            assert not err
            assert np.array_equal(nn, ref[2]) and nn.sum() > 1000
            for i in range(6):
                assert np.array_equal(kp[i, :nn[i]], ref[0][i, :nn[i]]) and np.array_equal(dd[i, :nn[i]], ref[1][i, :nn[i]])
                assert np.array_equal(kl[i, :ln[i]], lref[0][i, :ln[i]]) and np.array_equal(ld[i, :ln[i]], lref[1][i, :ln[i]])
            assert np.array_equal(ln, lref[3])
        with pytest.raises(Exception):
            orb.stage_batch(synth.frames(6200, 9, w, h))   # more than one chunk



def test_orb_stream_can_wait_for_the_grower_launch(api, synth):
    """pl_line_stream_wait_grow_start: the ORB extractor's stream waits for the point where the line extractor's streaming stages end;
    both results are what they are without the dependency (and waiting on a handle that never extracted is a no-op)."""
    import torch
    N = api.N
    frames = synth.frames(6200, 6)
    d_in = torch.from_numpy(frames).cuda()
    orb = api.ORBextractor(800, 1.2, 8, 20, 7, max_batch=8)
    ex = api.LineExtractor(max_batch=8)
    ex.stream_wait_grow_start(orb.stream())
    cap = orb.max_keypoints()
    ref_l = ex.extract_batch(frames)
    ref_o = orb.extract_batch(frames)
    d_kps = torch.zeros((6, cap, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((6, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(6, dtype=torch.int32, device="cuda")
    d_kls = torch.zeros((6, 80, 17), dtype=torch.float32, device="cuda")
    d_ld = torch.zeros((6, 80, 32), dtype=torch.uint8, device="cuda")
    d_lc = torch.zeros((6, 80, 3), dtype=torch.float64, device="cuda")
    d_ln = torch.zeros(6, dtype=torch.int32, device="cuda")
    for _ in range(2):
        ex.extract_batch_dev(d_in.data_ptr(), 6, 480, 640, 640, 640 * 480, 80, d_kls.data_ptr(), d_ld.data_ptr(), d_lc.data_ptr(), d_ln.data_ptr())
        ex.stream_wait_grow_start(orb.stream())
        orb.extract_batch_dev(d_in.data_ptr(), 6, 480, 640, 640, 640 * 480, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr())
        orb.sync()
        ex.sync()
    n = d_n.cpu().numpy()
    assert np.array_equal(n, ref_o[2]) and np.array_equal(d_ln.cpu().numpy(), ref_l[3])
    dd, ld = d_desc.cpu().numpy(), d_ld.cpu().numpy()
    for i in range(6):
        assert np.array_equal(dd[i, :n[i]], ref_o[1][i, :n[i]]) and np.array_equal(ld[i, :ref_l[3][i]], ref_l[1][i, :ref_l[3][i]])


def test_graph_replay_gives_the_same_lines(api, synth):
    """pl_line_set_graph: a single-frame call captured once and replayed (host-pointer entry: same staging buffers every call), a call
    with other parameters in between (re-capture), and a batch (direct launches) all give what the direct path gives."""
    frames = synth.frames(6300, 6)
    ref = api.LineExtractor(max_batch=8)
    want = [ref.ExtractLineSegment(f) for f in frames]
    ex = api.LineExtractor(max_batch=8)
    ex.set_graph(True)
    for rnd in range(2):
        for f, w in zip(frames, want):
            kl, ld, lc = ex.ExtractLineSegment(f)
            assert np.array_equal(kl, w[0]) and np.array_equal(ld, w[1]) and np.array_equal(lc, w[2])
        small = np.ascontiguousarray(frames[0][:240, :320])
        ks, ds, cs = ex.ExtractLineSegment(small)          # other geometry: re-capture
        kr, dr, cr = ref.ExtractLineSegment(small)
        assert np.array_equal(ks, kr) and np.array_equal(ds, dr)
        kb = ex.extract_batch(frames)                       # six frames: launched directly
        rb = ref.extract_batch(frames)
        assert np.array_equal(kb[3], rb[3]) and np.array_equal(kb[1], rb[1])


def test_max_lines_parameter(api, synth, oracle):
    img = synth.frame(1000, 640, 480)
    ex = api.LineExtractor()
    for m in (1, 17, 200):
        kls, desc, co = ex.ExtractLineSegment(img, max_lines=m)
        okl, odesc, oco = oracle.line_extract(img, m)
        assert len(kls) == len(okl) == m
        assert np.array_equal(kls["class_id"], okl["class_id"]) and np.array_equal(desc, odesc)


@pytest.mark.parametrize("env", [{"PLSLAM_LSD_POOL_TILES": "2"}, {"PLSLAM_LSD_POOL_TILES": "1", "PLSLAM_LSD_WINDOW": "1"},
                                 {"PLSLAM_LSD_GROW2": "64,64,1"}, {"PLSLAM_LSD_GROW2": "1024,384,2", "PLSLAM_LSD_WINDOW": "128", "PLSLAM_LSD_LOOKAHEAD": "32"},
                                 {"PLSLAM_LSD_GROW2": "512,256,3", "PLSLAM_LSD_FORCE_MANY": "1"},
                                 {"PLSLAM_LSD_GROW2": "128,160,4", "PLSLAM_LSD_WINDOW": "7", "PLSLAM_LSD_LOOKAHEAD": "1"},
                                 {"PLSLAM_LSD_GROW2": "512,256,3", "PLSLAM_LSD_TAIL_NFA": "0", "PLSLAM_LSD_POLL_NS": "5000"}])
def test_grower_configuration_does_not_change_the_result(env, api, synth, monkeypatch):
    """The speculative region grower must give the sequential result whatever its shape: number of grower
    warps per CTA (PLSLAM_LSD_GROW2 = threads for up to one frame per SM, threads and CTAs per SM for more), window of
    uncommitted tickets, tickets issued ahead of the growers, and a private tile pool so small that most regions overflow
    and are re-grown at commit time.  (The knobs are read when the extractor is created.)"""
    frames = synth.frames(4242, 5)
    ref = api.LineExtractor(max_batch=5)
    k0, d0, c0, n0 = ref.extract_batch(frames)
    s0 = ref.lsd_segments(frame=4)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    ex = api.LineExtractor(max_batch=5)
    k1, d1, c1, n1 = ex.extract_batch(frames)
    s1 = ex.lsd_segments(frame=4)
    assert np.array_equal(n0, n1) and np.array_equal(k0, k1) and np.array_equal(d0, d1) and np.array_equal(c0, c1)
    for a, b in zip(s0, s1):
        assert np.array_equal(a, b)
    ex.close()
    ref.close()


def test_reserved_sms_do_not_change_the_result(api, synth):
    """pl_line_set_reserved_sms only changes how many persistent grower CTAs are launched when a batch has more frames than SMs
    (frames come from a counter); the segments must be the same, and a matcher call must be able to run next to the grower."""
    frames = synth.frames(777, 160, 320, 240)   # more frames than SMs: the reservation applies
    ex = api.LineExtractor(max_cols=320, max_rows=240, max_batch=160)
    k0, d0, c0, n0 = ex.extract_batch(frames)
    for r in (16, 100, 147):
        ex.set_reserved_sms(r)
        k1, d1, c1, n1 = ex.extract_batch(frames)
        assert np.array_equal(n0, n1) and np.array_equal(k0, k1) and np.array_equal(d0, d1) and np.array_equal(c0, c1), r
    with pytest.raises(api.N.PlError):
        ex.set_reserved_sms(100000)
    ex.close()


def test_full_hd_frame(api, synth, oracle):
    """1920x1080: the committed bitmap of a frame takes most of the shared memory, so the grower runs one frame per CTA
    with fewer warps and a smaller tile pool — same result."""
    img = synth.frame(77, 1920, 1080)
    ex = api.LineExtractor(max_cols=1920, max_rows=1080)
    kls, desc, co = ex.ExtractLineSegment(img)
    okl, odesc, oco = oracle.line_extract(img, 80)
    assert len(kls) == len(okl) and np.array_equal(kls["class_id"], okl["class_id"])
    for f in ("sx", "sy", "ex", "ey"):
        assert np.abs(kls[f] - okl[f]).max() <= ENDPOINT_TOL_PX
    assert np.array_equal(desc, odesc)
    frames = np.stack([img, synth.frame(78, 1920, 1080), img])
    exb = api.LineExtractor(max_cols=1920, max_rows=1080, max_batch=3)
    k2, d2, c2, n2 = exb.extract_batch(frames)
    assert n2[0] == len(okl) and np.array_equal(d2[0, :n2[0]], odesc) and np.array_equal(d2[2, :n2[2]], odesc)


def test_device_double_sincos_equals_host_libm(api):
    """region2rect's rec.dx / rec.dy: the device restatement of the host's sin / cos (csrc/pl_glibc_sincos.cuh) on LSD's own
    arguments (float degrees -> radians, +pi) and on uniform doubles: bit-identical to numpy (= the host libm)."""
    import ctypes as C
    rng = np.random.default_rng(5)
    deg = rng.integers(0, 3600000, 400000).astype(np.float64) / 10000.0
    x = np.concatenate([deg.astype(np.float32).astype(np.float64) * (np.pi / 180), deg.astype(np.float32).astype(np.float64) * (np.pi / 180) + np.pi,
                        rng.uniform(-20, 20, 400000), np.array([0.0, 1e-9, 0.126, 0.85546875, 2.426265, np.pi / 2, np.pi, -np.pi])])
    x = np.ascontiguousarray(x)
    s = np.empty_like(x)
    c = np.empty_like(x)
    N = api.N
    N.check(N.lib().pl_test_sincos(N.ptr(x), C.c_int(len(x)), N.ptr(s), N.ptr(c)))
    assert np.array_equal(s, np.sin(x)) and np.array_equal(c, np.cos(x)), f"{(s != np.sin(x)).sum()} sin / {(c != np.cos(x)).sum()} cos values differ"


def test_lsd_segment_sets_equal_the_oracle_over_many_frames(api, synth, oracle):
    """Segment-set equality sweep: every LSD segment (end points, width, precision, NFA) of 96 frames from four generators."""
    bad = []
    ex = api.LineExtractor(max_batch=24)
    for seed in (99, 6000, 12345, 777):
        fr = synth.frames(seed, 24)
        ex.extract_batch(fr)
        for j in range(24):
            lines, width, prec, nfa = ex.lsd_segments(frame=j)
            ol, ow, op, on = oracle.lsd_detect(fr[j])
            ok = lines.shape == ol.shape and np.abs(lines - ol).max() <= ENDPOINT_TOL_PX and np.abs(width - ow).max() <= 1e-6 and np.array_equal(prec, op) \
                and (np.abs(nfa - on) <= 1e-6 * np.maximum(1.0, np.abs(on))).all()
            if not ok:
                bad.append((seed, j))
    assert not bad, bad
