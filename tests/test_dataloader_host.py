"""CPU: host half of the training input (tecogan_b200/lib/dataloader.py::HRClipLoader after reference
lib/dataloader.py:147-273): directory layout and skipping rules, window count, moving-first-frame / crop / flip
augmentation semantics, determinism under threads, rank sharding, error behaviour."""
import os

import numpy as np
import pytest
import torch


class Flags:
    def __init__(self, **kw):
        self.__dict__.update(dict(input_video_dir="", input_video_pre="scene", str_dir=2000, end_dir=2003, max_frm=11, RNN_N=4,
                                  batch_size=2, crop_size=4, rand_seed=1, queue_thread=3, movingFirstFrame=True, random_crop=True,
                                  flip=True, mode="train"))
        self.__dict__.update(kw)


def _make_dataset(root, dirs=(2000, 2001, 2003), frames=12, h=40, w=52, short=()):
    import cv2
    for d in dirs:
        p = os.path.join(root, "scene_%04d" % d)
        os.makedirs(p)
        n = 5 if d in short else frames
        for f in range(n):
            # pixel (y, x) of frame f in directory d encodes its own coordinates: R = y, G = x, B = 10*(d-2000)+f
            img = np.zeros((h, w, 3), np.uint8)
            img[..., 0] = np.arange(h)[:, None]
            img[..., 1] = np.arange(w)[None, :]
            img[..., 2] = 10 * (d - 2000) + f
            cv2.imwrite(os.path.join(p, "col_high_%04d.png" % f), img[:, :, ::-1])


def test_layout_windows_and_errors(tmp_path):
    from tecogan_b200.lib.dataloader import HRClipLoader
    with pytest.raises(ValueError, match="input_video_dir is not provided"):
        HRClipLoader(Flags())
    with pytest.raises(ValueError, match="not found"):
        HRClipLoader(Flags(input_video_dir=str(tmp_path / "nope")))
    root = str(tmp_path / "data")
    _make_dataset(root, short=(2001,))
    L = HRClipLoader(Flags(input_video_dir=root))
    assert [os.path.basename(c) for c in L.clips] == ["scene_2000", "scene_2003"]      # 2001 too short, 2002 absent
    assert L.windows == 11 - 4 + 1 and len(L) == 2 * 8 and L.steps_per_epoch == 8
    assert L.tar_size == 4 * 4 + 8
    with pytest.raises(Exception, match="No frame files"):
        HRClipLoader(Flags(input_video_dir=root, str_dir=2050, end_dir=2060))
    with pytest.raises(Exception, match="Not implemented"):
        HRClipLoader(Flags(input_video_dir=root, random_crop=False)).sample(0)


def test_augmentation_semantics(tmp_path):
    from tecogan_b200.lib.dataloader import HRClipLoader
    root = str(tmp_path / "data")
    _make_dataset(root, dirs=(2000,))
    F = Flags(input_video_dir=root, str_dir=2000, end_dir=2000)
    L = HRClipLoader(F)
    seen_moving = seen_plain = seen_flip = 0
    for idx in range(len(L)):
        for epoch in range(6):
            s = L.sample(idx, epoch)
            assert s.shape == (4, 24, 24, 3) and s.dtype == np.float32 and 0.0 <= s.min() and s.max() <= 1.0
            y = np.rint(s[..., 0] * 255).astype(int)
            x = np.rint(s[..., 1] * 255).astype(int)
            fidx = np.rint(s[..., 2] * 255).astype(int)
            for t in range(4):
                assert len(np.unique(fidx[t])) == 1                   # one source frame per output frame
                assert (np.diff(y[t], axis=0) == 1).all()             # contiguous rows of the source
                dx = np.diff(x[t], axis=1)
                assert (dx == 1).all() or (dx == -1).all()            # contiguous columns, possibly mirrored
            flipped = bool((np.diff(x[0], axis=1) == -1).all())
            seen_flip += flipped
            frames = fidx[:, 0, 0]
            start = idx % L.windows
            if (frames == start).all():                               # moving first frame: crops of frame `start`
                seen_moving += 1
                oy, ox = y[:, 0, 0], (x[:, 0, -1] if flipped else x[:, 0, 0])
                steps = np.stack([np.diff(ox), np.diff(oy)], axis=1)
                assert np.abs(steps).max() <= 4                       # floor(U(-3.5, 4.5)) in [-4, 4]
            else:
                seen_plain += 1
                assert (frames == start + np.arange(4)).all()         # consecutive frames of the window
                assert (y[:, 0, 0] == y[0, 0, 0]).all() and (x[:, 0, 0] == x[0, 0, 0]).all()   # one crop for all frames
    n = len(L) * 6
    assert 0.15 < seen_moving / n < 0.45 and 0.35 < seen_flip / n < 0.65 and seen_plain > 0
    L2 = HRClipLoader(Flags(input_video_dir=root, str_dir=2000, end_dir=2000, movingFirstFrame=False, flip=False))
    s = L2.sample(3, 0)
    assert (np.rint(s[:, 0, 0, 2] * 255).astype(int) == 3 + np.arange(4)).all() and (np.diff(s[0, 0, :, 1]) > 0).all()


def test_batches_are_deterministic_sharded_and_cover_an_epoch(tmp_path):
    from tecogan_b200.lib.dataloader import HRClipLoader
    root = str(tmp_path / "data")
    _make_dataset(root, dirs=(2000, 2001))
    F = Flags(input_video_dir=root, str_dir=2000, end_dir=2001)
    it1, it2 = HRClipLoader(F).batches(), HRClipLoader(Flags(input_video_dir=root, str_dir=2000, end_dir=2001, queue_thread=1)).batches()
    a = [next(it1) for _ in range(9)]
    b = [next(it2) for _ in range(9)]
    assert a[0].shape == (2, 4, 24, 24, 3) and a[0].dtype == torch.float32
    for x, y in zip(a, b):
        assert torch.equal(x, y)                                       # independent of the number of worker threads
    L = HRClipLoader(F)
    assert sorted(L._order(0).tolist()) == list(range(16)) and L._order(0).tolist() != L._order(1).tolist()
    r0, r1 = HRClipLoader(F, rank=0, world=2), HRClipLoader(F, rank=1, world=2)
    assert sorted(r0._order(3).tolist() + r1._order(3).tolist()) == list(range(16))      # ranks split every epoch's clips
    assert r0.steps_per_epoch == 4
    resumed = HRClipLoader(F).batches(start_step=5)
    assert torch.equal(next(resumed), a[5])
