#!/usr/bin/env python
"""Entry point mirroring the reference's main.py: same flag names (main.py:30-105), `--mode inference` follows the
per-frame loop of main.py:253-268, `--mode train` follows main.py:273-430 (TecoGAN when --ratio > 0, else FRVSR).

Differences forced by the environment, stated once:
  * weights: TensorFlow checkpoints cannot be read here yet; --checkpoint takes a .pt file written by this
    program (name -> tensor, TF variable names) or `random:<seed>` for a seeded xavier initialisation;
  * training data: when --input_video_dir is missing, seeded synthetic HR clips stand in for the TF queue loader
    (lib/dataloader.py:52-273, out of scope); the device half (Gaussian down-sampling, crops) is the real one;
  * --precision {bf16,fp32} selects tcgen05 tensor-core or fp32 CUDA-core convolutions for inference;
  * under `torchrun` each rank trains on its own clip shard with one NCCL all-reduce per step.
"""
import argparse
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLAG_DEFS = [  # (name, type, default) -- reference main.py:32-103
    ('rand_seed', int, 1), ('input_dir_LR', str, None), ('input_dir_len', int, -1), ('input_dir_HR', str, None),
    ('mode', str, 'inference'), ('output_dir', str, None), ('output_pre', str, ''), ('output_name', str, 'output'),
    ('output_ext', str, 'jpg'), ('summary_dir', str, None), ('checkpoint', str, None), ('num_resblock', int, 16),
    ('pre_trained_model', bool, False), ('vgg_ckpt', str, None), ('cudaID', str, '0'), ('queue_thread', int, 6),
    ('name_video_queue_capacity', int, 512), ('video_queue_capacity', int, 256), ('video_queue_batch', int, 2),
    ('RNN_N', int, 10), ('batch_size', int, 4), ('flip', bool, True), ('random_crop', bool, True),
    ('movingFirstFrame', bool, True), ('crop_size', int, 32), ('input_video_dir', str, ''), ('input_video_pre', str, 'scene'),
    ('str_dir', int, 1000), ('end_dir', int, 2000), ('end_dir_val', int, 2050), ('max_frm', int, 119),
    ('vgg_scaling', float, -0.002), ('warp_scaling', float, 1.0), ('pingpang', bool, False), ('pp_scaling', float, 1.0),
    ('EPS', float, 1e-12), ('learning_rate', float, 0.0001), ('decay_step', int, 500000), ('decay_rate', float, 0.5),
    ('stair', bool, False), ('beta', float, 0.9), ('adameps', float, 1e-8), ('max_epoch', int, None), ('max_iter', int, 1000000),
    ('display_freq', int, 20), ('summary_freq', int, 100), ('save_freq', int, 10000), ('ratio', float, 0.01),
    ('Dt_mergeDs', bool, True), ('Dt_ratio_0', float, 1.0), ('Dt_ratio_add', float, 0.0), ('Dt_ratio_max', float, 1.0),
    ('Dbalance', float, 0.4), ('crop_dt', float, 0.75), ('D_LAYERLOSS', bool, True),
    ('precision', str, 'bf16'),   # extension
]


def parse_flags(argv=None):
    """tf.app.flags-style parsing: `--name value`, `--name=value`, booleans as `--name` / `--noname`."""
    ap = argparse.ArgumentParser(allow_abbrev=False)
    for name, typ, default in FLAG_DEFS:
        if typ is bool:
            ap.add_argument('--' + name, dest=name, nargs='?', const=True, default=default,
                            type=lambda s: s.lower() in ('1', 'true', 'yes'))
            ap.add_argument('--no' + name, dest=name, action='store_false')
        else:
            ap.add_argument('--' + name, type=typ, default=default)
    return ap.parse_args(argv)


def load_checkpoint(store, spec, num_resblock, need_d=False, need_vgg=False):
    import torch
    from tecogan_b200.init_params import xavier_params
    if spec.startswith('random:'):
        store.load(xavier_params(int(spec.split(':', 1)[1]), num_resblock, need_d, need_vgg))
    else:
        store.load(torch.load(spec, map_location='cpu'))


def inference(FLAGS):
    import numpy as np
    import torch
    from tecogan_b200 import config, variables as V
    from tecogan_b200.engine import InferenceEngine
    from tecogan_b200.lib.dataloader import inference_data_loader
    from tecogan_b200.lib.ops import save_img
    if FLAGS.checkpoint is None:
        raise ValueError('The checkpoint file is needed to performing the test.')
    inference_data = inference_data_loader(FLAGS)
    h, w = inference_data.inputs[0].shape[:2]
    print("input shape:", [1, h, w, 3])
    print("output shape:", [1, h * 4, w * 4, 3])
    config.set_precision(FLAGS.precision)
    store = V.set_default_store(V.VariableStore(seed=FLAGS.rand_seed))
    load_checkpoint(store, FLAGS.checkpoint, FLAGS.num_resblock)
    eng = InferenceEngine(h, w, FLAGS.num_resblock)
    print('Finish building the network')
    image_dir = FLAGS.output_dir if FLAGS.output_pre == "" else os.path.join(FLAGS.output_dir, FLAGS.output_pre)
    os.makedirs(image_dir, exist_ok=True)
    max_iter = len(inference_data.inputs)
    srtime = 0
    print('Frame evaluation starts!!')
    to_dev = lambda k: torch.from_numpy(np.array([inference_data.inputs[k]]).astype(np.float32)).cuda()
    nxt = to_dev(0)
    for i in range(max_iter):
        input_im, nxt = nxt, (to_dev(i + 1) if i + 1 < max_iter else None)
        t0 = time.time()
        out = eng.step(input_im, next_lr=nxt)     # the next frame's flow is estimated while this frame is generated
        torch.cuda.synchronize()
        srtime += time.time() - t0
        if i >= 5:
            name, _ = os.path.splitext(os.path.basename(str(inference_data.paths_LR[i])))
            filename = FLAGS.output_name + '_' + name
            print('saving image %s' % filename)
            save_img(os.path.join(image_dir, "%s.%s" % (filename, FLAGS.output_ext)), out[0])
        else:   # First 5 is a hard-coded symmetric frame padding, ignored but time added!
            print("Warming up %d" % (5 - i))
    print("total time " + str(srtime) + ", frame number " + str(max_iter))


def synthetic_hr_batch(FLAGS, step, rank, device):
    """Seeded smooth-noise HR clips [B,RNN_N,4crop+8,4crop+8,3] in [0,1] moving by a constant sub-pixel velocity."""
    import torch
    g = torch.Generator().manual_seed(1000003 * rank + step)
    B, T, S = FLAGS.batch_size, FLAGS.RNN_N, FLAGS.crop_size * 4 + 8
    base = torch.rand(B, 3, S // 8 + 8, S // 8 + 8, generator=g)
    big = torch.nn.functional.interpolate(base, scale_factor=8, mode='bicubic', align_corners=False).clamp(0, 1)
    out = torch.empty(B, T, S, S, 3)
    for t in range(T):
        oy, ox = 8 + 2 * t, 8 + t
        out[:, t] = big[:, :, oy:oy + S, ox:ox + S].permute(0, 2, 3, 1)
    return out.to(device)


def train(FLAGS):
    import torch
    import torch.distributed as dist
    from tecogan_b200 import variables as V
    from tecogan_b200.lib.dataloader import frvsr_gpu_data_loader
    from tecogan_b200.lib.Teco import FRVSR, TecoGAN
    rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from tecogan_b200 import config
    config.set_train_precision(FLAGS.precision)   # bf16: tcgen05 forward + input-gradient convs, fp32 master weights
    store = V.set_default_store(V.VariableStore(seed=FLAGS.rand_seed))   # same seed on every rank -> identical init
    gan = FLAGS.ratio > 0
    if FLAGS.checkpoint is not None:
        load_checkpoint(store, FLAGS.checkpoint, FLAGS.num_resblock, gan, FLAGS.vgg_scaling > 0)
    elif FLAGS.vgg_scaling > 0:
        print('[main] no vgg_19.ckpt reader yet: VGG19 uses seeded random weights (frozen)')
    dev = torch.device('cuda', local_rank)
    lr0, tg0 = frvsr_gpu_data_loader(synthetic_hr_batch(FLAGS, 0, rank, dev), FLAGS)
    Net = TecoGAN(lr0, tg0, FLAGS) if gan else FRVSR(lr0, tg0, FLAGS)
    print('Finish building the network.')
    frame_len = (FLAGS.RNN_N * 2 - 1) if FLAGS.pingpang else FLAGS.RNN_N
    max_iter, start = FLAGS.max_iter, time.time()
    try:
        for step in range(max_iter):
            lr_, tg_ = frvsr_gpu_data_loader(synthetic_hr_batch(FLAGS, step, rank, dev), FLAGS)
            res = Net.train(lr_, tg_)
            run_step = Net.global_step()
            if step == 0 and rank == 0:
                print('Optimization starts!!!(Ctrl+C to stop, will try saving the last model...)')
            if (run_step % FLAGS.display_freq) == 0 and rank == 0:
                rate = (step + 1) * FLAGS.batch_size * world / (time.time() - start)
                remaining = (max_iter - step) * FLAGS.batch_size * world / rate
                print("progress  step %d  image/sec %0.1fx%02d  remaining %dh%dm" %
                      (run_step, rate, frame_len, remaining // 3600, (remaining % 3600) // 60))
                print("global_step", run_step)
                print("learning_rate", res["lr"])
                for name, value in zip(Net.update_list_name, Net.update_list_avg()):
                    print(name, value)
            if (run_step % FLAGS.save_freq) == 0 and rank == 0:
                print('Save the checkpoint')
                torch.save({k: v.detach().cpu() for k, v in store.items()}, os.path.join(FLAGS.output_dir, 'model-%d.pt' % run_step))
    except KeyboardInterrupt:
        if rank == 0:
            print('main.py: KeyboardInterrupt->saving the checkpoint')
            torch.save({k: v.detach().cpu() for k, v in store.items()}, os.path.join(FLAGS.output_dir, 'model-%d.pt' % Net.global_step()))
        print('main.py: quit')
        sys.exit(0)
    if rank == 0:
        torch.save({k: v.detach().cpu() for k, v in store.items()}, os.path.join(FLAGS.output_dir, 'model-%d.pt' % Net.global_step()))
    print('Optimization done!!!!!!!!!!!!')
    if world > 1:
        dist.destroy_process_group()


def main(argv=None):
    FLAGS = parse_flags(argv)
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", FLAGS.cudaID) if 'LOCAL_RANK' not in os.environ else None
    if FLAGS.output_dir is None:
        raise ValueError('The output directory is needed')
    os.makedirs(FLAGS.output_dir, exist_ok=True)
    if FLAGS.summary_dir:
        os.makedirs(FLAGS.summary_dir, exist_ok=True)
    if FLAGS.mode == 'inference':
        inference(FLAGS)
    elif FLAGS.mode == 'train':
        train(FLAGS)
    else:
        raise ValueError("mode must be 'inference' or 'train'")


if __name__ == '__main__':
    main()
