"""TensorFlow checkpoints without TensorFlow: reader for the "bundle V2" format (`<prefix>.index` +
`<prefix>.data-NNNNN-of-MMMMM`) and for the older single-file V1 format (e.g. slim's `vgg_19.ckpt`), plus a V2 writer.

Replaces, for this path, `tf.train.Saver.restore/save` and `tf.train.load_checkpoint` at the reference's call sites
main.py:224,245 (inference: generator + fnet), main.py:307,340-352,362-366,418-421 (training: resume / pre-trained
weights / VGG / periodic save) and lib/ops.py:370-391 (`get_existing_from_ckpt`: has_tensor / get_tensor / shape check).
Variable names are the TF names of SURVEY.md App. C, which is what `tecogan_b200.variables.VariableStore` uses as keys.

PARITY UNPINNED: TensorFlow is not installable in this environment and the reference ships no checkpoint, so these
routines follow the published on-disk formats (tensorflow/core/util/tensor_bundle/tensor_bundle.cc, .../lib/io/table*.cc
= LevelDB's table format, .../util/saved_tensor_slice.proto, .../framework/tensor.proto) and are tested against
hand-assembled files and their own writer only; the first real checkpoint read should be checked by eye (shapes/ranges).

Formats in one paragraph.  A *table* file is a sequence of blocks, each followed by a 5-byte trailer (compression type:
0 none / 1 snappy, masked CRC32C of block + type), and ends with a 48-byte footer (metaindex handle, index handle,
zero padding, magic 0xdb4775248b80fb57).  A block is a run of entries (varint shared-key-prefix length, varint
unshared length, varint value length, key suffix, value) and a trailer of restart offsets.  The index block maps
separator keys to data-block handles (varint offset, varint size).  V2: key "" -> BundleHeaderProto, key <variable
name> -> BundleEntryProto{dtype, shape, shard_id, offset, size, crc32c}; tensor bytes live little-endian in the data
shards.  V1: key "" -> SavedTensorSlices{meta}, other keys -> SavedTensorSlices{data: SavedSlice{name, slice,
TensorProto}} with the values inside the proto.
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xdb4775248b80fb57
_MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DT_BFLOAT16, DT_STRING = 14, 7
_DT_OF = {np.dtype(v): k for k, v in DTYPES.items()}


# ------------------------------------------------------------------ CRC32C (Castagnoli), masked as in crc32c.h
def _make_table():
    tab = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        tab.append(c)
    return tab


_CRC_TAB = _make_table()
_native = None


def _native_crc():
    """libteco's host-side teco_crc32c when the library is built (hundreds of MB/s); None otherwise."""
    global _native
    if _native is None:
        try:
            from . import _ffi
            _native = _ffi.lib().teco_crc32c
        except Exception:
            _native = False
    return _native or None


def crc32c(data, crc=0):
    """CRC32C of a bytes-like object, continuing from `crc`.  Inputs above 4 KB use libteco's host routine when the
    library is present; the pure-Python loop below is the definition (a few MB/s: pass verify=False to the readers to
    skip the checks on very large files if libteco is not built)."""
    data = bytes(data)
    if len(data) > 4096:
        fn = _native_crc()
        if fn is not None:
            r = fn(data, len(data), crc)
            if r >= 0:
                return int(r)
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TAB
    for b in data:
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c):
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + _MASK_DELTA) & 0xFFFFFFFF


def unmask_crc(m):
    r = (m - _MASK_DELTA) & 0xFFFFFFFF
    return ((r >> 17) | (r << 15)) & 0xFFFFFFFF


# ------------------------------------------------------------------ varints and a minimal protobuf wire decoder
def _get_varint(buf, pos):
    result, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise ValueError("tf_bundle: truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("tf_bundle: varint too long")


def _put_varint(v):
    if v < 0:
        v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def parse_proto(buf):
    """Wire-level decode: {field_number: [values]} with varints as ints, fixed32/64 as ints, length-delimited as bytes."""
    out, pos = {}, 0
    buf = bytes(buf)
    while pos < len(buf):
        tag, pos = _get_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _get_varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _get_varint(buf, pos)
            if pos + n > len(buf):
                raise ValueError("tf_bundle: truncated length-delimited field")
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("tf_bundle: unsupported protobuf wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


def _field(wire_type, number, payload):
    return _put_varint((number << 3) | wire_type) + payload


def _pb_varint(number, v):
    return _field(0, number, _put_varint(v))


def _pb_bytes(number, b):
    return _field(2, number, _put_varint(len(b)) + bytes(b))


def _parse_shape(buf):
    """TensorShapeProto: dim = 2 (repeated Dim{size = 1}), unknown_rank = 3."""
    msg = parse_proto(buf)
    if msg.get(3, [0])[0]:
        raise ValueError("tf_bundle: tensor of unknown rank")
    return tuple(_signed64(parse_proto(d).get(1, [0])[0]) for d in msg.get(2, []))


def _encode_shape(shape):
    return b"".join(_pb_bytes(2, _pb_varint(1, int(s))) for s in shape)


# ------------------------------------------------------------------ snappy (raw format) decompression
def snappy_decompress(buf):
    buf = bytes(buf)
    n, pos = _get_varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                   # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("tf_bundle: corrupt snappy stream")
        for _ in range(ln):                             # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("tf_bundle: snappy length mismatch (%d != %d)" % (len(out), n))
    return bytes(out)


# ------------------------------------------------------------------ table (SSTable) reader / writer
def _read_block(data, offset, size, verify):
    raw = data[offset:offset + size]
    if len(raw) != size or offset + size + 5 > len(data):
        raise ValueError("tf_bundle: block handle outside the file")
    ctype = data[offset + size]
    if verify:
        want = struct.unpack_from("<I", data, offset + size + 1)[0]
        if unmask_crc(want) != crc32c(data[offset:offset + size + 1]):
            raise ValueError("tf_bundle: block checksum mismatch at offset %d" % offset)
    if ctype == 1:
        raw = snappy_decompress(raw)
    elif ctype != 0:
        raise ValueError("tf_bundle: unknown block compression type %d" % ctype)
    return raw


def _block_entries(block):
    if len(block) < 4:
        raise ValueError("tf_bundle: block too small")
    nrestart = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestart
    if end < 0:
        raise ValueError("tf_bundle: bad restart array")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _get_varint(block, pos)
        unshared, pos = _get_varint(block, pos)
        vlen, pos = _get_varint(block, pos)
        if shared > len(key) or pos + unshared + vlen > end:
            raise ValueError("tf_bundle: corrupt block entry")
        key = key[:shared] + block[pos:pos + unshared]
        pos += unshared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(path, verify=True):
    """All (key, value) pairs of a table file, in key order."""
    with open(path, "rb") as f:
        data = f.read()
    if len(data) < 48:
        raise ValueError("tf_bundle: %s is too short to be a table file" % path)
    footer = data[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise ValueError("tf_bundle: %s is not a TensorFlow table file (bad magic number)" % path)
    _, p = _get_varint(footer, 0)          # metaindex handle (unused)
    _, p = _get_varint(footer, p)
    ioff, p = _get_varint(footer, p)
    isize, p = _get_varint(footer, p)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isize, verify)):
        boff, q = _get_varint(handle, 0)
        bsize, q = _get_varint(handle, q)
        out.extend(_block_entries(_read_block(data, boff, bsize, verify)))
    return out


class _TableWriter:
    """Uncompressed table with LevelDB's defaults (restart every 16 keys, ~4 KB blocks)."""

    def __init__(self, block_size=4096, restart_interval=16):
        self.buf = bytearray()
        self.block_size, self.restart_interval = block_size, restart_interval
        self.index = []                      # (last key of block, offset, size)
        self._reset()
        self.last_key = None

    def _reset(self):
        self.block, self.restarts, self.count, self.prev = bytearray(), [0], 0, b""

    def add(self, key, value):
        key, value = bytes(key), bytes(value)
        if self.last_key is not None and key <= self.last_key:
            raise ValueError("tf_bundle: table keys must be added in strictly increasing order")
        shared = 0
        if self.count and self.count % self.restart_interval == 0:
            self.restarts.append(len(self.block))
        elif self.count:
            m = min(len(self.prev), len(key))
            while shared < m and self.prev[shared] == key[shared]:
                shared += 1
        self.block += _put_varint(shared) + _put_varint(len(key) - shared) + _put_varint(len(value)) + key[shared:] + value
        self.prev = self.last_key = key
        self.count += 1
        if len(self.block) >= self.block_size:
            self._flush()

    def _emit(self, contents):
        off = len(self.buf)
        self.buf += contents + b"\x00" + struct.pack("<I", mask_crc(crc32c(bytes(contents) + b"\x00")))
        return off, len(contents)

    def _finish_block(self):
        blk = bytes(self.block) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))
        self._reset()
        return blk

    def _flush(self):
        if not self.count:
            return
        key = self.last_key
        off, size = self._emit(self._finish_block())
        self.index.append((key, off, size))

    def finish(self):
        self._flush()
        moff, msize = self._emit(self._finish_block())            # empty metaindex block
        for key, off, size in self.index:
            self.block += _put_varint(0) + _put_varint(len(key)) + _put_varint(len(_put_varint(off) + _put_varint(size)))
            self.block += key + _put_varint(off) + _put_varint(size)
            self.restarts = [0]
        # index block: one restart, no prefix compression (shared = 0 everywhere is valid)
        ioff, isize = self._emit(self._finish_block())
        footer = _put_varint(moff) + _put_varint(msize) + _put_varint(ioff) + _put_varint(isize)
        footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
        self.buf += footer
        return bytes(self.buf)


# ------------------------------------------------------------------ bundle V2
def _bf16_to_f32(raw):
    u = np.frombuffer(raw, dtype="<u2").astype(np.uint32) << 16
    return u.view(np.float32)


class BundleReader:
    """`tf.train.load_checkpoint(prefix)` stand-in: has_tensor / get_tensor / shape / dtype / keys."""

    def __init__(self, prefix, verify=True):
        self.prefix, self.verify = prefix, verify
        index = prefix if prefix.endswith(".index") else prefix + ".index"
        self.prefix = index[:-len(".index")]
        self.entries = {}
        self.num_shards = 1
        for key, value in read_table(index, verify):
            msg = parse_proto(value)
            if key == b"":
                self.num_shards = msg.get(1, [1])[0]
                if msg.get(2, [0])[0] != 0:
                    raise ValueError("tf_bundle: big-endian checkpoints are not supported")
                continue
            self.entries[key.decode("utf-8")] = {
                "dtype": msg.get(1, [0])[0], "shape": _parse_shape(msg[2][0]) if 2 in msg else (),
                "shard": msg.get(3, [0])[0], "offset": _signed64(msg.get(4, [0])[0]), "size": _signed64(msg.get(5, [0])[0]),
                "crc": msg.get(6, [None])[0], "sliced": 7 in msg}
        self._shards = {}

    def keys(self):
        return sorted(self.entries)

    def has_tensor(self, name):
        return name in self.entries

    def shape(self, name):
        return self.entries[name]["shape"]

    def dtype(self, name):
        return self.entries[name]["dtype"]

    def _shard(self, i):
        if i not in self._shards:
            self._shards[i] = open("%s.data-%05d-of-%05d" % (self.prefix, i, self.num_shards), "rb")
        return self._shards[i]

    def get_tensor(self, name):
        if name not in self.entries:
            raise KeyError("tf_bundle: tensor %r not found in checkpoint %s" % (name, self.prefix))
        e = self.entries[name]
        if e["sliced"]:
            raise ValueError("tf_bundle: %r is a partitioned variable (slices are not supported)" % name)
        if e["dtype"] == DT_STRING:
            raise ValueError("tf_bundle: %r is a string tensor (not supported)" % name)
        f = self._shard(e["shard"])
        f.seek(e["offset"])
        raw = f.read(e["size"])
        if len(raw) != e["size"]:
            raise ValueError("tf_bundle: data shard is truncated (tensor %r)" % name)
        if self.verify and e["crc"] is not None and unmask_crc(e["crc"]) != crc32c(raw):
            raise ValueError("tf_bundle: checksum mismatch for tensor %r" % name)
        if e["dtype"] == DT_BFLOAT16:
            arr = _bf16_to_f32(raw)
        elif e["dtype"] in DTYPES:
            arr = np.frombuffer(raw, dtype=np.dtype(DTYPES[e["dtype"]]).newbyteorder("<"))
        else:
            raise ValueError("tf_bundle: unsupported dtype enum %d for tensor %r" % (e["dtype"], name))
        n = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
        if arr.size != n:
            raise ValueError("tf_bundle: tensor %r has %d elements on disk, shape %s" % (name, arr.size, e["shape"]))
        return arr.reshape(e["shape"]).copy()

    def close(self):
        for f in self._shards.values():
            f.close()
        self._shards = {}


def write_bundle(prefix, tensors):
    """Write {name: array-like} as a single-shard V2 checkpoint `<prefix>.index` + `<prefix>.data-00000-of-00001`."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    table = _TableWriter()
    # BundleHeaderProto: num_shards = 1, endianness = 2 (LITTLE = 0, omitted), version = 3 {producer = 1}
    table.add(b"", _pb_varint(1, 1) + _pb_bytes(3, _pb_varint(1, 1)))
    offset = 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            arr = np.asarray(tensors[name])          # (ascontiguousarray would turn a scalar into shape (1,))
            if arr.dtype not in _DT_OF:
                raise ValueError("tf_bundle: cannot store dtype %s (tensor %r)" % (arr.dtype, name))
            raw = arr.astype(arr.dtype.newbyteorder("<"), copy=False).tobytes(order="C")
            f.write(raw)
            entry = _pb_varint(1, _DT_OF[arr.dtype]) + _pb_bytes(2, _encode_shape(arr.shape))
            if offset:
                entry += _pb_varint(4, offset)
            entry += _pb_varint(5, len(raw)) + _field(5, 6, struct.pack("<I", mask_crc(crc32c(raw))))
            table.add(name.encode("utf-8"), entry)
            offset += len(raw)
    with open(prefix + ".index", "wb") as f:
        f.write(table.finish())


# ------------------------------------------------------------------ checkpoint V1 (single table file, values inside protos)
def _tensor_proto_values(tp, name):
    """TensorProto (framework/tensor.proto): dtype = 1, tensor_shape = 2, tensor_content = 4, half_val = 13,
    float_val = 5, double_val = 6, int_val = 7, int64_val = 10, bool_val = 11 (packed or repeated)."""
    msg = parse_proto(tp)
    dt = msg.get(1, [0])[0]
    if 4 in msg and msg[4][0]:
        if dt == DT_BFLOAT16:
            return _bf16_to_f32(msg[4][0])
        return np.frombuffer(msg[4][0], dtype=np.dtype(DTYPES[dt]).newbyteorder("<"))

    def packed(field, fmt, size):
        vals = []
        for chunk in msg.get(field, []):
            if isinstance(chunk, bytes):
                vals.extend(struct.unpack("<%d%s" % (len(chunk) // size, fmt), chunk))
            else:                               # unpacked fixed-width element, already decoded as an unsigned int
                vals.append(struct.unpack("<" + fmt, struct.pack("<" + ("I" if size == 4 else "Q"), chunk))[0])
        return vals

    def varints(field):
        vals = []
        for chunk in msg.get(field, []):
            if isinstance(chunk, bytes):
                pos = 0
                while pos < len(chunk):
                    v, pos = _get_varint(chunk, pos)
                    vals.append(_signed64(v))
            else:
                vals.append(_signed64(chunk))
        return vals

    if dt == 1:
        return np.asarray(packed(5, "f", 4), dtype=np.float32)
    if dt == 2:
        return np.asarray(packed(6, "d", 8), dtype=np.float64)
    if dt in (3, 4, 5, 6, 17):
        return np.asarray(varints(7)).astype(DTYPES[dt])
    if dt == 9:
        return np.asarray(varints(10), dtype=np.int64)
    if dt == 10:
        return np.asarray(varints(11)).astype(np.bool_)
    raise ValueError("tf_bundle: V1 checkpoint tensor %r has unsupported dtype enum %d" % (name, dt))


class V1Reader:
    """Same interface for the pre-bundle format written by `tf.train.Saver(write_version=V1)` (slim model-zoo files).
    Only full (unpartitioned) slices are supported, which is all a Saver of plain variables writes."""

    def __init__(self, path, verify=True):
        self.path = path
        self.meta, self._data = {}, {}
        for key, value in read_table(path, verify):
            sts = parse_proto(value)             # SavedTensorSlices: meta = 1, data = 2
            if key == b"":
                for m in parse_proto(sts[1][0]).get(1, []) if 1 in sts else []:
                    sm = parse_proto(m)          # SavedSliceMeta: name = 1, shape = 2, type = 3, slice = 4
                    self.meta[sm[1][0].decode("utf-8")] = {"shape": _parse_shape(sm[2][0]) if 2 in sm else (),
                                                           "dtype": sm.get(3, [0])[0]}
            elif 2 in sts:
                sl = parse_proto(sts[2][0])      # SavedSlice: name = 1, slice = 2, data = 3
                self._data.setdefault(sl[1][0].decode("utf-8"), []).append(sl[3][0])

    def keys(self):
        return sorted(self.meta)

    def has_tensor(self, name):
        return name in self.meta

    def shape(self, name):
        return self.meta[name]["shape"]

    def dtype(self, name):
        return self.meta[name]["dtype"]

    def get_tensor(self, name):
        if name not in self.meta:
            raise KeyError("tf_bundle: tensor %r not found in checkpoint %s" % (name, self.path))
        chunks = self._data.get(name, [])
        if len(chunks) != 1:
            raise ValueError("tf_bundle: %r is stored as %d slices (only whole tensors are supported)" % (name, len(chunks)))
        shape = self.meta[name]["shape"]
        vals = _tensor_proto_values(chunks[0], name)
        n = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if vals.size != n:
            raise ValueError("tf_bundle: tensor %r has %d values, shape %s" % (name, vals.size, shape))
        return np.array(vals).reshape(shape)

    def close(self):
        pass


def is_tf_checkpoint(spec):
    return os.path.isfile(spec + ".index") or spec.endswith(".index") or (os.path.isfile(spec) and _has_table_magic(spec))


def _has_table_magic(path):
    try:
        with open(path, "rb") as f:
            f.seek(-8, os.SEEK_END)
            return struct.unpack("<Q", f.read(8))[0] == TABLE_MAGIC
    except (OSError, struct.error):
        return False


def load_checkpoint(spec, verify=True):
    """`tf.train.load_checkpoint` stand-in: a V2 prefix (`model-1000`, with `.index` next to it) or a V1 file."""
    if spec.endswith(".index") or os.path.isfile(spec + ".index"):
        return BundleReader(spec, verify)
    if os.path.isfile(spec) and _has_table_magic(spec):
        return V1Reader(spec, verify)
    raise ValueError("tf_bundle: %r is neither a V2 checkpoint prefix (no %s.index) nor a V1 checkpoint file" % (spec, spec))


def get_existing_from_ckpt(reader, wanted, rest_zero=False, print_level=1):
    """Reference lib/ops.py:370-391 on a name -> shape dict: returns {name: float32 array} for the variables present,
    zeros for the missing ones when rest_zero, and raises the same ValueError on a shape mismatch."""
    out = {}
    for name, shape in wanted.items():
        shape = tuple(int(s) for s in shape)
        if reader.has_tensor(name):
            val = reader.get_tensor(name)
            if tuple(val.shape) != shape:
                raise ValueError('Wrong shape in for {} in ckpt,expected {}, got {}.'.format(name, str(shape), str(tuple(val.shape))))
            out[name] = np.asarray(val, dtype=np.float32)
        else:
            if print_level >= 1:
                print("variable not found in ckpt: " + name)
            if rest_zero:
                if print_level >= 1:
                    print("Assign Zero of " + str(shape))
                out[name] = np.zeros(shape, dtype=np.float32)
    return out
