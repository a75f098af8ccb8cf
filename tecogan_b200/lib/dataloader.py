"""Mirror of the hot-path parts of the reference's lib/dataloader.py: the inference PNG loader (lib/dataloader.py:11-50)
and the device half of frvsr_gpu_data_loader (lib/dataloader.py:306-332).  The TF queue-runner training loader
(lib/dataloader.py:52-273) is CPU I/O + augmentation and out of scope."""
import collections
import os

import numpy as np

from .. import kernels as K
from .ops import preprocess


def inference_data_loader(FLAGS):
    """reference lib/dataloader.py:11-50: sorted PNG list -> RGB float32 /255 frames, with the hard-coded symmetric
    warm-up padding (list indices 5,4,3,2,1 prepended)."""
    filedir = FLAGS.input_dir_LR
    downSP = False
    if (FLAGS.input_dir_LR is None) or (not os.path.exists(FLAGS.input_dir_LR)):
        if (FLAGS.input_dir_HR is None) or (not os.path.exists(FLAGS.input_dir_HR)):
            raise ValueError('Input directory not found')
        filedir = FLAGS.input_dir_HR
        downSP = True
    import cv2 as cv
    names = [_ for _ in os.listdir(filedir) if _.endswith(".png")]
    names = sorted(names)
    names.sort(key=lambda f: int(''.join(list(filter(str.isdigit, f))) or -1))
    if FLAGS.input_dir_len > 0:
        names = names[:FLAGS.input_dir_len]
    image_list_LR = [os.path.join(filedir, _) for _ in names]

    def preprocess_test(name):
        im = cv.imread(name, 3).astype(np.float32)[:, :, ::-1]
        if downSP:
            icol_blur = cv.GaussianBlur(im, (0, 0), sigmaX=1.5)
            im = icol_blur[::4, ::4, ::]
        return np.ascontiguousarray(im / 255.0)

    image_LR = [preprocess_test(_) for _ in image_list_LR]
    image_list_LR = image_list_LR[5:0:-1] + image_list_LR
    image_LR = image_LR[5:0:-1] + image_LR
    Data = collections.namedtuple('Data', 'paths_LR, inputs')
    return Data(paths_LR=image_list_LR, inputs=image_LR)


def frvsr_gpu_data_loader(HR_frames, FLAGS):
    """Device half of reference lib/dataloader.py:306-332.  HR_frames: CUDA [B,RNN_N,crop*4+8,crop*4+8,3] in [0,1]
    (the CPU queue's output).  Returns (s_inputs [B,T,crop,crop,3] in [0,1], s_targets [B,T,4crop,4crop,3] in [-1,1]):
    LR = 9x9 sigma-1.5 Gaussian stride-4 VALID; target = centre crop, preprocessed."""
    B, T, Hh, Ww, _ = HR_frames.shape
    crop = FLAGS.crop_size
    k_w_border = int(1.5 * 3.0)
    if Hh != crop * 4 + 2 * k_w_border or Ww != Hh:
        raise ValueError("frvsr_gpu_data_loader: HR frames must be %dx%d" % (crop * 4 + 2 * k_w_border, crop * 4 + 2 * k_w_border))
    flat = HR_frames.reshape(B * T, Hh, Ww, 3).contiguous()
    lr = K.gauss_down4(flat).reshape(B, T, crop, crop, 3)
    tgt = flat[:, k_w_border:k_w_border + crop * 4, k_w_border:k_w_border + crop * 4, :].contiguous()
    tgt = preprocess(tgt).reshape(B, T, crop * 4, crop * 4, 3)
    return lr, tgt
