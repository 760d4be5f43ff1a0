"""GPU tests (-m gpu) of the device-pointer entry points (pl_orb_extract_batch_dev / pl_line_extract_batch_dev): same
results as the host-pointer calls, asynchronous, and capacity problems are reported by pl_*_sync."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_dev_api_matches_host_api_and_reports_capacity(api, synth):
    import torch
    N = api.N
    frames = synth.frames(909, 6)
    F, H, W = frames.shape
    orb = api.ORBextractor(1000, 1.2, 8, 20, 7, max_batch=4)
    line = api.LineExtractor(max_batch=4)
    kps, desc, cnt = orb.extract_batch(frames)
    kls, ldesc, lco, lcnt = line.extract_batch(frames)
    cap = orb.max_keypoints()
    d_gray = torch.from_numpy(frames).cuda()
    d_kps = torch.zeros((F, cap, 7), dtype=torch.float32, device="cuda")
    d_desc = torch.zeros((F, cap, 32), dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(F, dtype=torch.int32, device="cuda")
    d_kls = torch.zeros((F, 80, 17), dtype=torch.float32, device="cuda")
    d_ldesc = torch.zeros((F, 80, 32), dtype=torch.uint8, device="cuda")
    d_lco = torch.zeros((F, 80, 3), dtype=torch.float64, device="cuda")
    d_ln = torch.zeros(F, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    orb.extract_batch_dev(d_gray.data_ptr(), F, H, W, W, W * H, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr())
    line.extract_batch_dev(d_gray.data_ptr(), F, H, W, W, W * H, 80, d_kls.data_ptr(), d_ldesc.data_ptr(), d_lco.data_ptr(), d_ln.data_ptr())
    orb.sync()
    line.sync()
    n = d_n.cpu().numpy()
    assert np.array_equal(n, cnt)
    hk = d_kps.cpu().numpy().view(N.KP_DTYPE).reshape(F, cap)
    hd = d_desc.cpu().numpy()
    ln = d_ln.cpu().numpy()
    assert np.array_equal(ln, lcnt)
    hl = d_kls.cpu().numpy().view(N.KL_DTYPE).reshape(F, 80)
    for i in range(F):
        assert np.array_equal(hk[i, :n[i]], kps[i, :n[i]]) and np.array_equal(hd[i, :n[i]], desc[i, :n[i]])
        assert np.array_equal(hl[i, :ln[i]], kls[i, :ln[i]]) and np.array_equal(d_ldesc[i, :ln[i]].cpu().numpy(), ldesc[i, :ln[i]])
    # a caller capacity that is too small: the asynchronous call succeeds, the next sync reports it (once)
    small = 100
    orb.extract_batch_dev(d_gray.data_ptr(), F, H, W, W, W * H, d_kps.data_ptr(), d_desc.data_ptr(), small, d_n.data_ptr())
    with pytest.raises(Exception) as ei:
        orb.sync()
    assert "capacity" in str(ei.value)
    orb.sync()
