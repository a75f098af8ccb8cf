"""Seeded xavier-uniform parameter sets under the TF variable names for `--checkpoint random:<seed>`.
Product-side helper (does NOT import oracle/): weights only, same initialiser family as the reference
(tf.contrib.layers.xavier_initializer, lib/ops.py:40,52,98; zero biases)."""
import math
from collections import OrderedDict

import torch


def _xav(gen, shape, fan_in, fan_out):
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen) * 2 - 1) * lim


def xavier_params(seed, num_resblock=16, need_d=False, need_vgg=False, damp=0.5):
    """`damp` scales res-block/output weights: untrained xavier recurrences diverge over frames (DESIGN.md 'Parity')."""
    gen = torch.Generator().manual_seed(seed)
    p = OrderedDict()

    def conv(name, k, cin, cout, tr=False, bias=True, scale=1.0):
        shape = (k, k, cout, cin) if tr else (k, k, cin, cout)
        fi, fo = (k * k * cout, k * k * cin) if tr else (k * k * cin, k * k * cout)
        p[name + "/weights"] = _xav(gen, shape, fi, fo) * scale
        if bias:
            p[name + "/biases"] = torch.zeros(cout)
    g = "generator/generator_unit/"
    conv(g + "input_stage/conv/Conv", 3, 51, 64)
    for i in range(1, num_resblock + 1):
        conv(g + "resblock_%d/conv_1/Conv" % i, 3, 64, 64, scale=damp)
        conv(g + "resblock_%d/conv_2/Conv" % i, 3, 64, 64, scale=damp)
    for i in (1, 2):
        conv(g + "conv_tran2highres/conv_tran%d/Conv2d_transpose" % i, 3, 64, 64, tr=True)
    conv(g + "output_stage/conv/Conv", 3, 64, 3, scale=damp)
    f = "fnet/autoencode_unit/"
    for name, cin, cout in (("encoder_1", 6, 32), ("encoder_2", 32, 64), ("encoder_3", 64, 128),
                            ("decoder_1", 128, 256), ("decoder_2", 256, 128), ("decoder_3", 128, 64)):
        conv(f + name + "/conv_1/Conv", 3, cin, cout)
        conv(f + name + "/conv_2/Conv", 3, cout, cout)
    conv(f + "output_stage/conv1/Conv", 3, 64, 32)
    conv(f + "output_stage/conv2/Conv", 3, 32, 2)
    if need_d:
        d = "tdiscriminator/discriminator_unit/"
        conv(d + "input_stage/conv/Conv", 3, 27, 64)
        for name, ci, co in (("disblock_1", 64, 64), ("disblock_3", 64, 64), ("disblock_5", 64, 128), ("disblock_7", 128, 256)):
            conv(d + name + "/conv1/Conv", 4, ci, co, bias=False)
            p[d + name + "/BatchNorm/beta"] = torch.zeros(co)
        p[d + "dense_layer_2/dense/kernel"] = _xav(gen, (256, 1), 256, 1)
        p[d + "dense_layer_2/dense/bias"] = torch.zeros(1)
    if need_vgg:
        for blk, reps, cin, cout in ((1, 2, 3, 64), (2, 2, 64, 128), (3, 4, 128, 256), (4, 4, 256, 512), (5, 4, 512, 512)):
            c = cin
            for j in range(1, reps + 1):
                name = "vgg_19/conv%d/conv%d_%d" % (blk, blk, j)
                p[name + "/weights"] = torch.randn((3, 3, c, cout), generator=gen) * math.sqrt(2.0 / (9 * c))
                p[name + "/biases"] = torch.zeros(cout)
                c = cout
    return p


def variable_shapes(num_resblock=16, need_d=False, need_vgg=False):
    """TF variable name -> shape for the graphs of this path, without allocating anything: what a
    checkpoint must provide (`tf.get_collection(MODEL_VARIABLES, scope=...)` in the reference, main.py:221-222,313,317,323)."""
    p = OrderedDict()

    def conv(name, k, cin, cout, tr=False, bias=True):
        p[name + "/weights"] = (k, k, cout, cin) if tr else (k, k, cin, cout)
        if bias:
            p[name + "/biases"] = (cout,)
    g = "generator/generator_unit/"
    conv(g + "input_stage/conv/Conv", 3, 51, 64)
    for i in range(1, num_resblock + 1):
        conv(g + "resblock_%d/conv_1/Conv" % i, 3, 64, 64)
        conv(g + "resblock_%d/conv_2/Conv" % i, 3, 64, 64)
    for i in (1, 2):
        conv(g + "conv_tran2highres/conv_tran%d/Conv2d_transpose" % i, 3, 64, 64, tr=True)
    conv(g + "output_stage/conv/Conv", 3, 64, 3)
    f = "fnet/autoencode_unit/"
    for name, cin, cout in (("encoder_1", 6, 32), ("encoder_2", 32, 64), ("encoder_3", 64, 128),
                            ("decoder_1", 128, 256), ("decoder_2", 256, 128), ("decoder_3", 128, 64)):
        conv(f + name + "/conv_1/Conv", 3, cin, cout)
        conv(f + name + "/conv_2/Conv", 3, cout, cout)
    conv(f + "output_stage/conv1/Conv", 3, 64, 32)
    conv(f + "output_stage/conv2/Conv", 3, 32, 2)
    if need_d:
        d = "tdiscriminator/discriminator_unit/"
        conv(d + "input_stage/conv/Conv", 3, 27, 64)
        for name, ci, co in (("disblock_1", 64, 64), ("disblock_3", 64, 64), ("disblock_5", 64, 128), ("disblock_7", 128, 256)):
            conv(d + name + "/conv1/Conv", 4, ci, co, bias=False)
            p[d + name + "/BatchNorm/beta"] = (co,)
        p[d + "dense_layer_2/dense/kernel"] = (256, 1)
        p[d + "dense_layer_2/dense/bias"] = (1,)
    if need_vgg:
        for blk, reps, cin, cout in ((1, 2, 3, 64), (2, 2, 64, 128), (3, 4, 128, 256), (4, 4, 256, 512), (5, 4, 512, 512)):
            c = cin
            for j in range(1, reps + 1):
                name = "vgg_19/conv%d/conv%d_%d" % (blk, blk, j)
                p[name + "/weights"] = (3, 3, c, cout)
                p[name + "/biases"] = (cout,)
                c = cout
    return p
