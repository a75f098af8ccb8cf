"""CPU: TensorFlow checkpoint reader/writer without TensorFlow (tecogan_b200/tf_bundle.py).
PARITY UNPINNED: no real TF checkpoint exists in this environment, so the files read here are (a) assembled by hand in this
test from the published table / bundle layouts, independently of the module's writer, and (b) round trips of the writer."""
import os
import struct

import numpy as np
import pytest
import torch

from tecogan_b200 import tf_bundle as T


def _vi(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _crc_ref(data):                         # bitwise CRC32C (reflected 0x82F63B78), no tables
    c = 0xFFFFFFFF
    for byte in data:
        c ^= byte
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
    return c ^ 0xFFFFFFFF


def _masked(data):
    c = _crc_ref(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xa282ead8) & 0xFFFFFFFF


def _block(entries, restart_every=16):
    """LevelDB block with prefix compression, written independently of tf_bundle._TableWriter."""
    body, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_every == 0:
            restarts.append(len(body))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        body += _vi(shared) + _vi(len(k) - shared) + _vi(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    return bytes(body) + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))


def _table(blocks, ctype=0, compress=None):
    """blocks: list of entry lists -> file bytes (data blocks, empty metaindex, index, footer)."""
    out, handles = bytearray(), []

    def emit(contents, ct=0):
        off = len(out)
        out.extend(contents + bytes([ct]) + struct.pack("<I", _masked(contents + bytes([ct]))))
        return off, len(contents)
    for ent in blocks:
        raw = _block(ent, restart_every=2)
        if compress:
            off, size = emit(compress(raw), 1)
        else:
            off, size = emit(raw)
        handles.append((ent[-1][0], off, size))
    moff, msize = emit(_block([]))
    ioff, isize = emit(_block([(k, _vi(o) + _vi(s)) for k, o, s in handles]))
    footer = _vi(moff) + _vi(msize) + _vi(ioff) + _vi(isize)
    out.extend(footer + b"\x00" * (40 - len(footer)) + struct.pack("<II", 0x8b80fb57, 0xdb477524))
    return bytes(out)


def _pb(field, wt, payload):
    return _vi((field << 3) | wt) + payload


def _shape_pb(shape):
    return b"".join(_pb(2, 2, _vi(len(d)) + d) for d in (_pb(1, 0, _vi(s)) for s in shape))


def _hand_made_v2(tmp_path, compress=None):
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4) / 7
    b = np.array([1, -2, 3], dtype=np.int64)
    c = np.float32(2.5).reshape(())
    data = a.tobytes() + b.tobytes() + c.tobytes()

    def entry(dt, shape, off, size, raw):
        e = _pb(1, 0, _vi(dt)) + _pb(2, 2, _vi(len(_shape_pb(shape))) + _shape_pb(shape))
        if off:
            e += _pb(4, 0, _vi(off))
        return e + _pb(5, 0, _vi(size)) + _pb(6, 5, struct.pack("<I", _masked(raw)))
    header = _pb(1, 0, _vi(1)) + _pb(3, 2, _vi(2) + _pb(1, 0, _vi(1)))
    ents = [(b"", header),
            (b"generator/generator_unit/a/weights", entry(1, a.shape, 0, a.nbytes, a.tobytes())),
            (b"generator/generator_unit/b/step", entry(9, b.shape, a.nbytes, b.nbytes, b.tobytes())),
            (b"generator/generator_unit/c", entry(1, (), a.nbytes + b.nbytes, 4, c.tobytes()))]
    prefix = str(tmp_path / "model-7")
    open(prefix + ".index", "wb").write(_table([ents[:3], ents[3:]], compress=compress))
    open(prefix + ".data-00000-of-00001", "wb").write(data)
    return prefix, a, b, c


def test_crc32c_known_answer_and_mask():
    assert T.crc32c(b"123456789") == 0xE3069283            # the CRC-32C check value
    blob = os.urandom(70000)                                # > 4 KB: libteco's host routine when built
    assert T.crc32c(blob) == _crc_ref(blob)
    assert T.crc32c(blob[30000:], T.crc32c(blob[:30000])) == _crc_ref(blob)
    assert T.unmask_crc(T.mask_crc(0xDEADBEEF)) == 0xDEADBEEF
    assert T.mask_crc(_crc_ref(b"abc")) == _masked(b"abc")


def test_reads_a_hand_assembled_v2_bundle(tmp_path):
    prefix, a, b, c = _hand_made_v2(tmp_path)
    r = T.load_checkpoint(prefix)
    assert r.keys() == ["generator/generator_unit/a/weights", "generator/generator_unit/b/step", "generator/generator_unit/c"]
    assert r.has_tensor("generator/generator_unit/a/weights") and not r.has_tensor("nope")
    assert r.shape("generator/generator_unit/a/weights") == (2, 3, 4)
    np.testing.assert_array_equal(r.get_tensor("generator/generator_unit/a/weights"), a)
    np.testing.assert_array_equal(r.get_tensor("generator/generator_unit/b/step"), b)
    assert r.get_tensor("generator/generator_unit/c").shape == () and float(r.get_tensor("generator/generator_unit/c")) == 2.5
    with pytest.raises(KeyError):
        r.get_tensor("nope")
    r.close()


def _snappy_literal_and_copies(raw):
    """A valid snappy stream: literals, plus a 2-byte-offset copy for every repeated 8-byte run we can find cheaply."""
    out = bytearray(_vi(len(raw)))
    i = 0
    while i < len(raw):
        j = raw.find(raw[i:i + 8], max(0, i - 2000), i) if i + 8 <= len(raw) and i >= 8 else -1
        if j >= 0 and j + 8 <= i:
            out += bytes([((8 - 1) << 2) | 2]) + struct.pack("<H", i - j)
            i += 8
            continue
        n = min(40, len(raw) - i)
        out += bytes([(n - 1) << 2]) + raw[i:i + n]
        i += n
    return bytes(out)


def test_snappy_blocks_and_corruption_are_handled(tmp_path):
    raw = bytes(range(50)) * 7 + b"tail"
    assert T.snappy_decompress(_snappy_literal_and_copies(raw)) == raw
    prefix, a, _, _ = _hand_made_v2(tmp_path, compress=_snappy_literal_and_copies)
    np.testing.assert_array_equal(T.BundleReader(prefix).get_tensor("generator/generator_unit/a/weights"), a)
    # flip one byte of the tensor data -> tensor checksum; flip one byte of the index -> block checksum
    d = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    d[5] ^= 1
    open(prefix + ".data-00000-of-00001", "wb").write(d)
    with pytest.raises(ValueError, match="checksum"):
        T.BundleReader(prefix).get_tensor("generator/generator_unit/a/weights")
    assert T.BundleReader(prefix, verify=False).get_tensor("generator/generator_unit/a/weights").shape == (2, 3, 4)
    ix = bytearray(open(prefix + ".index", "rb").read())
    ix[3] ^= 0x40
    open(prefix + ".index", "wb").write(ix)
    with pytest.raises(ValueError):
        T.BundleReader(prefix)
    open(prefix + ".index", "wb").write(b"not a table file at all, but long enough to hold a footer........")
    with pytest.raises(ValueError, match="magic"):
        T.BundleReader(prefix)


def test_writer_round_trip_many_variables_multiple_blocks(tmp_path):
    rng = np.random.default_rng(0)
    tensors = {"scope_%03d/layer/Conv/weights" % i: rng.standard_normal((3, 3, i % 5 + 1, 4)).astype(np.float32) for i in range(300)}
    tensors["global_step"] = np.asarray(1234, dtype=np.int64)
    tensors["flags/u8"] = rng.integers(0, 255, (7,), dtype=np.uint8)
    tensors["empty"] = np.zeros((0, 3), dtype=np.float32)
    prefix = str(tmp_path / "ck" / "model-1234")
    T.write_bundle(prefix, tensors)
    assert os.path.getsize(prefix + ".index") > 3 * 4096          # several data blocks
    r = T.load_checkpoint(prefix)
    assert set(r.keys()) == set(tensors)
    for k, v in tensors.items():
        got = r.get_tensor(k)
        assert got.dtype == v.dtype and got.shape == v.shape
        np.testing.assert_array_equal(got, v)
    # the writer's index parses with the independent block walker of this test too: footer magic + first key is ""
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xdb4775248b80fb57
    with pytest.raises(ValueError):
        T.write_bundle(prefix, {"x": np.array(["a"])})


def _hand_made_v1(path):
    """slim-style V1 file: key "" -> SavedTensorSlices{meta}, one SavedSlice per variable, float_val packed."""
    w = (np.arange(18, dtype=np.float32).reshape(3, 3, 1, 2) - 4) / 3
    bvec = np.array([0.5, -1.5], dtype=np.float32)

    def meta(name, shape):
        return _pb(1, 2, _vi(len(name)) + name) + _pb(2, 2, _vi(len(_shape_pb(shape))) + _shape_pb(shape)) + _pb(3, 0, _vi(1))

    def data(name, arr):
        tp = _pb(1, 0, _vi(1)) + _pb(5, 2, _vi(arr.nbytes) + arr.tobytes())
        sl = _pb(1, 2, _vi(len(name)) + name) + _pb(2, 2, _vi(0)) + _pb(3, 2, _vi(len(tp)) + tp)
        return _pb(2, 2, _vi(len(sl)) + sl)
    n1, n2 = b"vgg_19/conv1/conv1_1/weights", b"vgg_19/conv1/conv1_1/biases"
    metas = b"".join(_pb(1, 2, _vi(len(m)) + m) for m in (meta(n2, bvec.shape), meta(n1, w.shape)))
    ents = [(b"", _pb(1, 2, _vi(len(metas)) + metas)), (b"\x00k1", data(n2, bvec)), (b"\x00k2", data(n1, w))]
    open(path, "wb").write(_table([ents]))
    return w, bvec


def test_reads_a_hand_assembled_v1_checkpoint(tmp_path):
    path = str(tmp_path / "vgg_19.ckpt")
    w, bvec = _hand_made_v1(path)
    r = T.load_checkpoint(path)
    assert isinstance(r, T.V1Reader) and r.keys() == ["vgg_19/conv1/conv1_1/biases", "vgg_19/conv1/conv1_1/weights"]
    np.testing.assert_array_equal(r.get_tensor("vgg_19/conv1/conv1_1/weights"), w)
    np.testing.assert_array_equal(r.get_tensor("vgg_19/conv1/conv1_1/biases"), bvec)
    with pytest.raises(ValueError):
        T.load_checkpoint(str(tmp_path / "missing"))


def test_main_load_checkpoint_follows_the_reference_rules(tmp_path):
    """Saver.restore strictness, --pre_trained_model zero fill (lib/ops.py:370-391), shape check message, save round trip."""
    import main as M
    from tecogan_b200.init_params import variable_shapes, xavier_params

    class Store(dict):
        def load(self, params):
            self.update({k: v.clone() for k, v in params.items()})
    params = xavier_params(3, num_resblock=2, need_d=True)
    assert {k: tuple(v.shape) for k, v in params.items()} == dict(variable_shapes(2, True, False))
    out = str(tmp_path / "run")
    os.makedirs(out)
    M.save_checkpoint(params, out, 50)
    assert os.path.isfile(os.path.join(out, "model-50.index")) and os.path.isfile(os.path.join(out, "model-50.pt"))
    st = Store()
    M.load_checkpoint(st, os.path.join(out, "model-50"), 2, need_d=True)
    assert set(st) == set(params) and all(torch.equal(st[k], params[k]) for k in params)
    # a deeper graph than the checkpoint: strict restore refuses, --pre_trained_model zero-fills generator variables
    with pytest.raises(ValueError, match="lacks"):
        M.load_checkpoint(Store(), os.path.join(out, "model-50"), 3)
    st = Store()
    M.load_checkpoint(st, os.path.join(out, "model-50"), 3, need_d=True, pre_trained_model=True)
    assert float(st["generator/generator_unit/resblock_3/conv_1/Conv/weights"].abs().sum()) == 0.0
    assert torch.equal(st["fnet/autoencode_unit/encoder_1/conv_1/Conv/weights"], params["fnet/autoencode_unit/encoder_1/conv_1/Conv/weights"])
    # wrong shape -> the reference's message
    bad = {k: v.numpy() for k, v in params.items()}
    bad["generator/generator_unit/output_stage/conv/Conv/weights"] = np.zeros((3, 3, 64, 4), dtype=np.float32)
    T.write_bundle(os.path.join(out, "bad-1"), bad)
    with pytest.raises(ValueError, match="Wrong shape in for generator/generator_unit/output_stage/conv/Conv/weights"):
        M.load_checkpoint(Store(), os.path.join(out, "bad-1"), 2)
    # VGG from a V1 file: only the first conv exists in this hand-made file -> names the missing ones
    vpath = str(tmp_path / "vgg_19.ckpt")
    _hand_made_v1(vpath)
    with pytest.raises(ValueError, match="lacks"):
        M.load_vgg_checkpoint(Store(), vpath)
