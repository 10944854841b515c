"""Reader for TensorFlow "V2" checkpoints (tensor bundles) without TensorFlow.

The reference saves and restores its models with ``tf.train.Saver`` (reference
vae/models.py:154,172-186 and ppo.py:184,202-216).  The shipped checkpoints
(``vae/models/*/checkpoints/model.ckpt-N.{index,data-00000-of-00001}`` and
``models/pretrained_agent/checkpoints/...``) are therefore TF tensor bundles:

* ``.index`` is a LevelDB-format sorted string table.  The last 48 bytes are the
  footer ``[metaindex handle][index handle][padding][magic]`` where a handle is
  ``varint64 offset, varint64 size`` and the magic is 0xdb4775248b80fb57 (LE).
  A block is a run of prefix-compressed entries ``varint shared | varint
  non_shared | varint value_len | key_delta | value`` followed by a
  ``uint32 restarts[n], uint32 n`` trailer; after the block come one compression
  byte and a 4-byte CRC that are not counted in the handle's size.
* The key ``""`` maps to a ``BundleHeaderProto``; every other key is a variable
  name mapping to a ``BundleEntryProto`` {1: dtype, 2: shape, 3: shard_id,
  4: offset, 5: size, 6: crc32c}.
* ``.data-00000-of-00001`` holds the raw little-endian tensors.

Only what the shipped fixtures need is implemented: uncompressed blocks, one
data shard, float32/int32 tensors.
"""
from __future__ import annotations

import os
import re
import struct
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8")}


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _block_entries(buf: bytes, offset: int, size: int) -> Iterator[Tuple[bytes, bytes]]:
    block = buf[offset:offset + size]
    if offset + size < len(buf) and buf[offset + size] != 0:
        raise ValueError("compressed index blocks are not supported")
    (n_restarts,) = struct.unpack_from("<I", block, len(block) - 4)
    end = len(block) - 4 - 4 * n_restarts
    pos = 0
    key = b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        value_len, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        value = block[pos:pos + value_len]
        pos += value_len
        yield key, value


def _parse_proto(buf: bytes) -> Dict[int, list]:
    """Minimal protobuf wire-format reader: field number -> list of raw values."""
    out: Dict[int, list] = {}
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            val, pos = _varint(buf, pos)
        elif wire == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wire == 2:
            length, pos = _varint(buf, pos)
            val = buf[pos:pos + length]
            pos += length
        elif wire == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        out.setdefault(field, []).append(val)
    return out


def _parse_shape(buf: bytes) -> Tuple[int, ...]:
    dims = []
    for dim in _parse_proto(buf).get(2, []):
        d = _parse_proto(dim)
        dims.append(int(d.get(1, [0])[0]))
    return tuple(dims)


class BundleReader:
    """Reads every tensor of a ``<prefix>.index`` / ``<prefix>.data-00000-of-00001`` pair."""

    def __init__(self, prefix: str):
        self.prefix = prefix
        with open(prefix + ".index", "rb") as f:
            idx = f.read()
        if len(idx) < 48 or struct.unpack_from("<Q", idx, len(idx) - 8)[0] != _MAGIC:
            raise ValueError("%s.index is not a TF tensor-bundle index" % prefix)
        footer = idx[-48:]
        pos = 0
        _, pos = _varint(footer, pos)      # metaindex offset
        _, pos = _varint(footer, pos)      # metaindex size
        index_off, pos = _varint(footer, pos)
        index_size, pos = _varint(footer, pos)
        self.entries: Dict[str, Tuple[np.dtype, Tuple[int, ...], int, int]] = {}
        for _, handle in _block_entries(idx, index_off, index_size):
            boff, p = _varint(handle, 0)
            bsize, p = _varint(handle, p)
            for key, value in _block_entries(idx, boff, bsize):
                if key == b"":
                    continue                # BundleHeaderProto
                e = _parse_proto(value)
                dtype = _DTYPES[int(e[1][0])]
                shape = _parse_shape(e[2][0]) if 2 in e else ()
                if int(e.get(3, [0])[0]) != 0:
                    raise ValueError("multi-shard bundles are not supported")
                offset = int(e.get(4, [0])[0])
                size = int(e[5][0])
                self.entries[key.decode()] = (dtype, shape, offset, size)
        self._data_path = prefix + ".data-00000-of-00001"

    def keys(self) -> List[str]:
        return sorted(self.entries)

    def __contains__(self, name: str) -> bool:
        return name in self.entries

    def get(self, name: str) -> np.ndarray:
        dtype, shape, offset, size = self.entries[name]
        with open(self._data_path, "rb") as f:
            f.seek(offset)
            raw = f.read(size)
        arr = np.frombuffer(raw, dtype=dtype)
        return arr.reshape(shape).copy()

    def all(self) -> Dict[str, np.ndarray]:
        with open(self._data_path, "rb") as f:
            blob = f.read()
        out = {}
        for name, (dtype, shape, offset, size) in self.entries.items():
            out[name] = np.frombuffer(blob[offset:offset + size], dtype=dtype).reshape(shape).copy()
        return out


def latest_checkpoint(checkpoint_dir: str) -> Optional[str]:
    """``tf.train.latest_checkpoint``: parse the text ``checkpoint`` state file."""
    state = os.path.join(checkpoint_dir, "checkpoint")
    if not os.path.isfile(state):
        return None
    with open(state) as f:
        m = re.search(r'^model_checkpoint_path:\s*"(.*)"', f.read(), re.M)
    if not m:
        return None
    path = m.group(1)
    if not os.path.isabs(path):
        path = os.path.join(checkpoint_dir, path)
    return path if os.path.isfile(path + ".index") else None


# ----------------------------------------------------------------------------- writer
_CRC_TABLE = None


def crc32c(data: bytes, crc: int = 0) -> int:
    """CRC-32C (Castagnoli), the checksum TF stores per tensor and per index block."""
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab.append(c)
        _CRC_TABLE = tab
    tab = _CRC_TABLE
    crc ^= 0xFFFFFFFF
    for b in memoryview(data).cast("B"):
        crc = tab[(crc ^ b) & 0xFF] ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    """The rotated + offset form LevelDB / TF store ("masked" CRC)."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _enc_varint(v: int) -> bytes:
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field(num: int, wire: int, payload: bytes) -> bytes:
    tag = _enc_varint((num << 3) | wire)
    if wire == 2:
        return tag + _enc_varint(len(payload)) + payload
    return tag + payload


def _block(entries) -> bytes:
    """One uncompressed table block: every entry is its own restart point (no prefix sharing)."""
    body = bytearray()
    restarts = []
    for key, value in entries:
        restarts.append(len(body))
        body += _enc_varint(0) + _enc_varint(len(key)) + _enc_varint(len(value)) + key + value
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    return bytes(body)


_DTYPE_ENUM = {np.dtype("<f4"): 1, np.dtype("<f8"): 2, np.dtype("<i4"): 3, np.dtype("<i8"): 9}


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    """Writes ``<prefix>.index`` + ``<prefix>.data-00000-of-00001`` in the TF-V2 tensor-bundle format described at the
    top of this file (one shard, one uncompressed data block, per-tensor crc32c), i.e. what ``tf.train.Saver.save``
    produces for the reference (vae/models.py:172-175, ppo.py:202-205) -- so a checkpoint written by this build can be
    restored by the reference's own ``saver.restore`` and vice versa."""
    names = sorted(tensors)
    data = bytearray()
    entries = []
    # key "" -> BundleHeaderProto{num_shards = 1, endianness = LITTLE (0, default), version{producer = 1}}
    header = _field(1, 0, _enc_varint(1)) + _field(3, 2, _field(1, 0, _enc_varint(1)))
    entries.append((b"", header))
    for name in names:
        arr = np.asarray(tensors[name])          # (ascontiguousarray would turn a 0-d variable into shape (1,))
        dt = arr.dtype.newbyteorder("<") if arr.dtype.byteorder == ">" else arr.dtype
        if np.dtype(dt) not in _DTYPE_ENUM:
            raise ValueError("%s: dtype %s cannot be stored" % (name, arr.dtype))
        raw = arr.astype(dt, copy=False).tobytes()
        shape = b"".join(_field(2, 2, _field(1, 0, _enc_varint(int(d)))) for d in arr.shape)
        entry = _field(1, 0, _enc_varint(_DTYPE_ENUM[np.dtype(dt)])) + _field(2, 2, shape)
        if len(data):
            entry += _field(4, 0, _enc_varint(len(data)))
        entry += _field(5, 0, _enc_varint(len(raw))) + _field(6, 5, struct.pack("<I", masked_crc32c(raw)))
        entries.append((name.encode(), entry))
        data += raw
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))

    def with_trailer(block: bytes) -> bytes:
        return block + b"\x00" + struct.pack("<I", masked_crc32c(block + b"\x00"))
    out = bytearray()
    data_block = _block(entries)
    data_handle = _enc_varint(0) + _enc_varint(len(data_block))
    out += with_trailer(data_block)
    meta_off = len(out)
    meta_block = _block([])
    out += with_trailer(meta_block)
    index_off = len(out)
    index_block = _block([(names[-1].encode() if names else b"", data_handle)])
    out += with_trailer(index_block)
    footer = _enc_varint(meta_off) + _enc_varint(len(meta_block)) + _enc_varint(index_off) + _enc_varint(len(index_block))
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    out += footer
    with open(prefix + ".index", "wb") as f:
        f.write(bytes(out))


def verify_bundle_crcs(prefix: str) -> int:
    """Re-computes the per-tensor crc32c of a bundle against the stored values; returns the number of tensors checked
    (raises on a mismatch).  Used on the reference's shipped checkpoints to pin this file's CRC / format code."""
    r = BundleReader(prefix)
    with open(prefix + ".index", "rb") as f:
        idx = f.read()
    footer = idx[-48:]
    pos = 0
    _, pos = _varint(footer, pos); _, pos = _varint(footer, pos)
    index_off, pos = _varint(footer, pos)
    index_size, pos = _varint(footer, pos)
    with open(prefix + ".data-00000-of-00001", "rb") as f:
        blob = f.read()
    n = 0
    for _, handle in _block_entries(idx, index_off, index_size):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        stored_block_crc = struct.unpack_from("<I", idx, boff + bsize + 1)[0]
        if masked_crc32c(idx[boff:boff + bsize + 1]) != stored_block_crc:
            raise ValueError("index block crc mismatch in %s" % prefix)
        for key, value in _block_entries(idx, boff, bsize):
            if key == b"":
                continue
            e = _parse_proto(value)
            offset = int(e.get(4, [0])[0]); size = int(e[5][0])
            stored = struct.unpack("<I", e[6][0])[0]
            if masked_crc32c(blob[offset:offset + size]) != stored:
                raise ValueError("crc mismatch for %s in %s" % (key.decode(), prefix))
            n += 1
    assert n == len(r.entries)
    return n
