#!/usr/bin/env python
"""Mirror of the reference's metrics.py command line (flags --output / --results / --targets, metrics.py:9-12) for the two
metrics that are image arithmetic: PSNR and SSIM on the Y channel, computed on the GPU by tecogan_b200/metrics.py
(teco_metrics_psnr_y_u8 / teco_metrics_ssim_y_u8).  Same frame selection (cutfr = 2, metrics.py:117,131), same cut of the
result to the target size (:134-135), same crop_8x8 (:171-172), same per-frame / per-folder / total printout and
metrics.csv layout (:205-242) restricted to keys = ["PSNR", "SSIM"].
LPIPS / tLP (AlexNet weights) and tOF (cv2 Farneback on the CPU) are not part of this path: DESIGN.md section 6.

  python metrics.py --output results/metric_log/ --results results/calendar --targets HR/calendar
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

KEYS = ["PSNR", "SSIM"]
CUTFR = 2


def listPNGinDir(dirpath):
    """metrics.py:28-35: *.png, not starting with IB, ordered by the digits in the name."""
    names = sorted(n for n in os.listdir(dirpath) if n.endswith(".png") and not n.startswith("IB"))
    names.sort(key=lambda f: int(''.join(filter(str.isdigit, f)) or -1))
    return [os.path.join(dirpath, n) for n in names]


def read_rgb(path):
    try:
        import cv2
        return np.ascontiguousarray(cv2.imread(path)[:, :, ::-1])
    except ImportError:
        from PIL import Image
        return np.ascontiguousarray(np.asarray(Image.open(path).convert("RGB")))


def evaluate_folder(result_dir, target_dir, log=print):
    """Per-frame PSNR / SSIM lists of one scene (frames cutfr .. n-cutfr-1), frames grouped by size into batched launches."""
    import torch
    from tecogan_b200 import metrics as M
    result, target = listPNGinDir(result_dir), listPNGinDir(target_dir)
    lists = {k: [] for k in KEYS}
    for i in range(CUTFR, len(target) - CUTFR):
        out, tar = read_rgb(result[i]), read_rgb(target[i])
        msg = "frame %d, tar %s, out %s, " % (i, str(tar.shape), str(out.shape))
        t, o = torch.from_numpy(tar).cuda(), torch.from_numpy(out).cuda()
        ps, ss = M.frame_metrics(t, o)            # window = crop_8x8 of the common area; no cropped copies
        y, x, _, _ = M.crop_window(min(tar.shape[0], out.shape[0]), min(tar.shape[1], out.shape[1]))
        lists["PSNR"].append(ps[0])
        lists["SSIM"].append(ss[0])
        log(result[i])
        log(msg + "psnr %02.2f, ssim %02.2f, crop (%d, %d)" % (ps[0], ss[0], y, x))
    return lists


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--output", required=True, help="the path of output directory")
    ap.add_argument("--results", required=True, help="the list of paths of result directory")
    ap.add_argument("--targets", required=True, help="the list of paths of target directory")
    F = ap.parse_args(argv)
    import pandas as pd
    os.makedirs(F.output, exist_ok=True)
    logf = open(os.path.join(F.output, "metricsfile.txt"), "a")

    def log(m):
        print(m)
        logf.write(m + "\n")
    result_list, target_list = F.results.split(','), F.targets.split(',')
    folder_n = len(result_list)
    sum_d = {"FrameAvg_" + k: 0.0 for k in KEYS}
    len_d = {k: 0 for k in KEYS}
    avg_d = {"Avg_" + k: [] for k in KEYS}
    folder_d = {"FolderAvg_" + k: 0.0 for k in KEYS}
    csv = os.path.join(F.output, "metrics.csv")
    for fi in range(folder_n):
        lists = evaluate_folder(result_list[fi], target_list[fi], log)
        pd_dict = {}
        for k in KEYS:
            cur = np.float32(lists[k])
            pd_dict["%s_%02d" % (k, fi)] = pd.Series(cur)
            mean = cur.sum() / cur.shape[0]
            log("%s_%02d, max %02.4f, min %02.4f, avg %02.4f" % (k, fi, cur.max(), cur.min(), mean))
            avg_d["Avg_" + k].append(mean)
            sum_d["FrameAvg_" + k] += cur.sum()
            len_d[k] += cur.shape[0]
            folder_d["FolderAvg_" + k] += mean
        pd.DataFrame(pd_dict).to_csv(csv, mode='w' if fi == 0 else 'a')
    for k in KEYS:
        sum_d["FrameAvg_" + k] = pd.Series([sum_d["FrameAvg_" + k] / len_d[k]])
        folder_d["FolderAvg_" + k] = pd.Series([folder_d["FolderAvg_" + k] / folder_n])
        avg_d["Avg_" + k] = pd.Series(np.float32(avg_d["Avg_" + k]))
        log("%s, total frame %d, total avg %02.4f, folder avg %02.4f" % (k, len_d[k], sum_d["FrameAvg_" + k][0], folder_d["FolderAvg_" + k][0]))
    for d in (avg_d, folder_d, sum_d):
        pd.DataFrame(d).to_csv(csv, mode='a')
    log("Finished.")
    logf.close()
    return {k: float(sum_d["FrameAvg_" + k][0]) for k in KEYS}


if __name__ == "__main__":
    main()
