"""GPU: PSNR / SSIM kernels (tecogan_b200/metrics.py -> teco_metrics_*_y_u8) against the reference's own psnr / crop_8x8
outputs (tests/golden/metrics.npz) and the oracle's SSIM restatement.  float64 arithmetic: tolerance 1e-9 dB / 1e-10."""
import math
import os

import numpy as np
import pytest
import torch

from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _g():
    return np.load(os.path.join(GOLDEN, "metrics.npz"), allow_pickle=False)


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_psnr_matches_reference_function_full_and_cropped():
    from tecogan_b200 import metrics as M
    g = _g()
    for i in range(int(g["n_cases"])):
        tgt, out = _cu(g["tgt%d" % i]), _cu(g["out%d" % i])
        assert abs(M.psnr(tgt, out) - float(g["psnr_full%d" % i])) < 1e-9
        win = M.crop_window(tgt.shape[0], tgt.shape[1])
        assert win == tuple(int(v) for v in g["crop%d" % i])
        ps, _ = M.frame_metrics(tgt, out)                      # crop passed as a window, no copy
        assert abs(ps[0] - float(g["psnr_crop%d" % i])) < 1e-9
        ct, y, x = M.crop_8x8(tgt)                             # the reference's call shape: crop first, then psnr
        co, _, _ = M.crop_8x8(out)
        assert (y, x) == win[:2]
        assert abs(M.psnr(ct.contiguous(), co.contiguous()) - float(g["psnr_crop%d" % i])) < 1e-9


def test_ssim_matches_oracle_on_ragged_sizes_and_batches():
    from oracle import teco_oracle as O
    from tecogan_b200 import metrics as M
    g = _g()
    for i in range(int(g["n_cases"])):
        tgt, out = g["tgt%d" % i], g["out%d" % i]
        assert abs(M.ssim(_cu(tgt), _cu(out)) - O.ssim_y(tgt, out)) < 1e-10
        _, ss = M.frame_metrics(_cu(tgt), _cu(out))
        assert abs(ss[0] - O.ssim_y(O.crop_8x8(tgt), O.crop_8x8(out))) < 1e-10
    # tiles that are cut on both axes, window smaller than one tile, exactly one window position
    tgt, out = g["tgt2"], g["out2"]
    for h, w in ((39, 45), (7, 7), (33, 8), (71, 101)):
        a, b = np.ascontiguousarray(tgt[:h, :w]), np.ascontiguousarray(out[:h, :w])
        assert abs(M.ssim(_cu(a), _cu(b)) - O.ssim_y(a, b)) < 1e-10, (h, w)
    # a batch: every frame gets its own data range and sums
    T = np.stack([g["tgt3"], g["tgt3"][::-1].copy(), g["out3"]])
    R = np.stack([g["out3"], g["out3"][::-1].copy(), g["out3"]])
    ps, ss = M.frame_metrics(_cu(T), _cu(R), crop=False)
    for n in range(3):
        assert abs(ss[n] - O.ssim_y(T[n], R[n])) < 1e-10
        if n < 2:
            assert abs(ps[n] - O.psnr_y(T[n], R[n])) < 1e-9
    assert math.isinf(ps[2]) and abs(ss[2] - 1.0) < 1e-12     # identical frames: rmse 0 -> inf (numpy), SSIM 1


def test_result_larger_than_target_is_cut_and_sequence_averages():
    from oracle import teco_oracle as O
    from tecogan_b200 import metrics as M
    g = _g()
    tgt = g["tgt1"][:141, :178]                                # target not divisible by 4: result is 144x180 (metrics.py:134)
    out = g["out1"]
    tgt = np.ascontiguousarray(tgt)
    ps, ss = M.frame_metrics(_cu(tgt), _cu(out))
    oc = out[:141, :178]
    assert abs(ps[0] - O.psnr_y(O.crop_8x8(tgt), O.crop_8x8(oc))) < 1e-9
    assert abs(ss[0] - O.ssim_y(O.crop_8x8(tgt), O.crop_8x8(oc))) < 1e-10
    T = np.stack([np.roll(g["tgt3"], k, axis=1) for k in range(7)])
    R = np.stack([np.roll(g["out3"], k, axis=1) for k in range(7)])
    r = M.evaluate_sequence(_cu(T), _cu(R), cutfr=2)
    assert len(r["PSNR"]) == 3
    want = [O.psnr_y(O.crop_8x8(T[k]), O.crop_8x8(R[k])) for k in range(2, 5)]
    assert abs(r["FolderAvg_PSNR"] - sum(want) / 3) < 1e-9


def test_misuse_raises_like_the_reference_stack():
    from tecogan_b200 import metrics as M
    a = torch.zeros(6, 40, 3, dtype=torch.uint8, device="cuda")
    with pytest.raises(ValueError):                            # skimage: win_size exceeds image extent
        M.ssim(a, a)
    with pytest.raises(ValueError):
        M.psnr(a.float(), a.float())
    with pytest.raises(ValueError):
        M.psnr(a.cpu(), a.cpu())                               # no CPU path
    with pytest.raises(ValueError):
        M.psnr_ssim_sums(a, a, window=(0, 0, 8, 8))            # window outside the frames


def test_metrics_command_line_on_png_folders(tmp_path):
    """Root metrics.py (the reference's flags): PNG folders in, per-frame log + metrics.csv out, averages = oracle."""
    from PIL import Image
    from oracle import teco_oracle as O
    import metrics as CLI
    g = _g()
    res, tar = tmp_path / "results" / "scene", tmp_path / "HR" / "scene"
    res.mkdir(parents=True)
    tar.mkdir(parents=True)
    T = [np.roll(g["tgt0"], 3 * k, axis=0) for k in range(6)]
    R = [np.roll(g["out0"], 3 * k, axis=0) for k in range(6)]
    for k in range(6):
        Image.fromarray(T[k]).save(tar / ("%04d.png" % k))
        Image.fromarray(R[k]).save(res / ("output_%04d.png" % k))
    Image.fromarray(R[0]).save(res / "IB_0000.png")            # skipped by listPNGinDir
    r = CLI.main(["--output", str(tmp_path / "log"), "--results", str(res), "--targets", str(tar)])
    want_p = np.mean(np.float32([O.psnr_y(O.crop_8x8(T[k]), O.crop_8x8(R[k])) for k in (2, 3)]))
    want_s = np.mean(np.float32([O.ssim_y(O.crop_8x8(T[k]), O.crop_8x8(R[k])) for k in (2, 3)]))
    assert abs(r["PSNR"] - want_p) < 1e-4 and abs(r["SSIM"] - want_s) < 1e-6
    txt = open(tmp_path / "log" / "metrics.csv").read()
    assert "PSNR_00" in txt and "FolderAvg_SSIM" in txt
