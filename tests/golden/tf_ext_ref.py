"""Second, independent derivation of the TensorFlow-1.x primitives the reference calls but does not spell out
([TF-ext] in SURVEY.md App. A): written in numpy straight from the operators' published definitions, sharing NO code
with oracle/teco_oracle.py (which restates them with torch functional ops).  Two uses, both test infrastructure:

  * tests/golden/tf_shim.py executes the reference's own Python on top of THESE functions, so the committed goldens
    do not depend on the oracle's restatement of a primitive (the pin is no longer circular);
  * tests/test_oracle_golden.py checks every oracle primitive against them on ragged shapes and clamp cases.

Definitions followed (reference call sites in brackets):
  conv2d             slim.conv2d / tf.nn.conv2d, NHWC, HWIO filter, 'SAME' or 'VALID'      [lib/ops.py:47-56, 347-367]
                     SAME: out = ceil(n/s), pad_total = max((out-1)*s + k - n, 0), before = pad_total // 2
  conv2d_transpose   tf.nn.conv2d_transpose == conv2d_backprop_input of the forward SAME convolution that maps the
                     (stride*n)-sized output back to n; filter [kh,kw,Cout,Cin]            [lib/ops.py:35-44]
  dense_image_warp   tf.contrib.image: query = (y,x) - flow; interpolate_bilinear with floor clamped to [0, size-2]
                     and the fractional part clamped to [0,1]; flow[...,0] is the row displacement  [lib/Teco.py:120,140]
  resize_bilinear    tf.image.resize_images default (bilinear, align_corners=False, legacy coordinates src = dst*in/out)
                                                                                           [lib/frvsr.py:21-22, lib/Teco.py:244]
  batch_norm_train   slim.batch_norm(is_training=True, scale=False): batch moments over N,H,W, biased variance
                                                                                           [lib/ops.py:88-90]
  max_pool_2x2       slim.max_pool2d([2,2]) stride 2 'VALID'                               [lib/ops.py:92-94]
  space_to_depth     tf.space_to_depth: out channel = (dy*bs + dx)*C + c                   [main.py:201]
  leaky_relu         keras LeakyReLU(alpha)                                                [lib/ops.py:84-85]
"""
import numpy as np


def same_padding(n, k, s):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    return out, total // 2, total - total // 2


def conv2d(x, w, b=None, stride=1, padding="SAME"):
    """y[n,oy,ox,co] = sum_{ky,kx,ci} xpad[n, oy*s+ky, ox*s+kx, ci] * w[ky,kx,ci,co]  (cross-correlation, as TF)."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    N, H, W, C = x.shape
    kh, kw, ci, co = w.shape
    assert ci == C, (x.shape, w.shape)
    if padding == "SAME":
        OH, pt, pb = same_padding(H, kh, stride)
        OW, pl, pr = same_padding(W, kw, stride)
    else:
        OH, OW = (H - kh) // stride + 1, (W - kw) // stride + 1
        pt = pb = pl = pr = 0
    xp = np.zeros((N, H + pt + pb, W + pl + pr, C))
    xp[:, pt:pt + H, pl:pl + W] = x
    y = np.zeros((N, OH, OW, co))
    for ky in range(kh):
        for kx in range(kw):
            patch = xp[:, ky:ky + (OH - 1) * stride + 1:stride, kx:kx + (OW - 1) * stride + 1:stride, :]
            y += patch @ w[ky, kx]
    if b is not None:
        y += np.asarray(b, np.float64)
    return y.astype(np.float32)


def conv2d_transpose(x, w, b=None, stride=2):
    """Gradient-of-conv definition.  Forward conv G: [N, s*H, s*W, Cout] -> [N, H, W, Cin], SAME, filter w[kh,kw,Cout,Cin];
    the transposed conv scatters every x[n,oy,ox,:] through the taps it would have been computed from."""
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    N, H, W, Cin = x.shape
    kh, kw, Cout, ci = w.shape
    assert ci == Cin
    OH, OW = H * stride, W * stride
    _, pt, _ = same_padding(OH, kh, stride)      # padding of the forward conv on the LARGE side
    _, pl, _ = same_padding(OW, kw, stride)
    y = np.zeros((N, OH, OW, Cout))
    for oy in range(H):
        for ky in range(kh):
            iy = oy * stride + ky - pt
            if iy < 0 or iy >= OH:
                continue
            for ox in range(W):
                for kx in range(kw):
                    ix = ox * stride + kx - pl
                    if ix < 0 or ix >= OW:
                        continue
                    y[:, iy, ix, :] += x[:, oy, ox, :] @ w[ky, kx].T
    if b is not None:
        y += np.asarray(b, np.float64)
    return y.astype(np.float32)


def dense_image_warp(image, flow):
    image = np.asarray(image, np.float32)
    flow = np.asarray(flow, np.float32)
    N, H, W, C = image.shape
    assert flow.shape == (N, H, W, 2)
    gy, gx = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    qy = gy[None] - flow[..., 0]
    qx = gx[None] - flow[..., 1]

    def split(q, size):
        fl = np.minimum(np.maximum(np.floor(q), 0.0), np.float32(size - 2))
        alpha = np.minimum(np.maximum(q - fl, np.float32(0.0)), np.float32(1.0))
        return fl.astype(np.int64), alpha.astype(np.float32)
    y0, ay = split(qy, H)
    x0, ax = split(qx, W)
    out = np.empty_like(image)
    for n in range(N):
        tl = image[n, y0[n], x0[n]]
        tr = image[n, y0[n], x0[n] + 1]
        bl = image[n, y0[n] + 1, x0[n]]
        br = image[n, y0[n] + 1, x0[n] + 1]
        top = ax[n][..., None] * (tr - tl) + tl
        bot = ax[n][..., None] * (br - bl) + bl
        out[n] = ay[n][..., None] * (bot - top) + top
    return out


def resize_bilinear(x, oh, ow):
    x = np.asarray(x, np.float32)
    N, H, W, C = x.shape

    def axis(n_in, n_out):
        scale = np.float32(n_in) / np.float32(n_out)
        src = np.arange(n_out, dtype=np.float32) * scale
        lo = np.floor(src).astype(np.int64)
        hi = np.minimum(lo + 1, n_in - 1)
        return lo, hi, (src - lo.astype(np.float32)).astype(np.float32)
    y0, y1, fy = axis(H, oh)
    x0, x1, fx = axis(W, ow)
    top = x[:, y0][:, :, x0] + (x[:, y0][:, :, x1] - x[:, y0][:, :, x0]) * fx[None, None, :, None]
    bot = x[:, y1][:, :, x0] + (x[:, y1][:, :, x1] - x[:, y1][:, :, x0]) * fx[None, None, :, None]
    return (top + (bot - top) * fy[None, :, None, None]).astype(np.float32)


def batch_norm_train(x, beta, eps=1e-3):
    x64 = np.asarray(x, np.float64)
    mean = x64.mean(axis=(0, 1, 2))
    var = ((x64 - mean) ** 2).mean(axis=(0, 1, 2))
    return ((x64 - mean) / np.sqrt(var + eps) + np.asarray(beta, np.float64)).astype(np.float32)


def max_pool_2x2(x):
    x = np.asarray(x, np.float32)
    N, H, W, C = x.shape
    oh, ow = H // 2, W // 2
    v = x[:, :2 * oh, :2 * ow].reshape(N, oh, 2, ow, 2, C)
    return v.max(axis=(2, 4))


def space_to_depth(x, bs=4):
    x = np.asarray(x, np.float32)
    N, H, W, C = x.shape
    out = np.empty((N, H // bs, W // bs, bs * bs * C), np.float32)
    for dy in range(bs):
        for dx in range(bs):
            out[..., (dy * bs + dx) * C:(dy * bs + dx + 1) * C] = x[:, dy::bs, dx::bs, :]
    return out


def leaky_relu(x, alpha):
    x = np.asarray(x, np.float32)
    return np.where(x >= 0, x, np.float32(alpha) * x).astype(np.float32)


def sigmoid(x):
    x = np.asarray(x, np.float64)
    return (1.0 / (1.0 + np.exp(-x))).astype(np.float32)
