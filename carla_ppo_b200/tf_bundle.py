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
