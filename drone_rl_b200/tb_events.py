"""Minimal TensorBoard event-file writer (scalars only) with no dependency on the ``tensorboard`` package -- the
reference logs through SB3's ``configure(tb_log_dir, ["stdout", "tensorboard"])`` (train.py:56-58) and is read with
``tensorboard --logdir ./tensorboard``; the package is not in this image, the file format is small:

  file      = sequence of TFRecords
  TFRecord  = uint64 length | uint32 masked_crc32c(length) | bytes data | uint32 masked_crc32c(data)     (little endian)
  data      = serialized ``tensorflow.Event`` protobuf: 1 wall_time (double) | 2 step (int64) |
              3 file_version (string, first record: "brain.Event:2") | 5 summary (message)
  Summary   = repeated 1 value { 1 tag (string) | 2 simple_value (float) | 4 image { 1 height | 2 width | 3 colorspace |
              4 encoded_image_string (PNG bytes) } }

``read_events`` parses the same subset back (used by the tests; CRCs verified).
"""
from __future__ import annotations

import os
import socket
import struct
import time

_POLY = 0x82F63B78          # CRC-32C (Castagnoli), reflected
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ (_POLY if _c & 1 else 0)
    _TABLE.append(_c)


def crc32c(data: bytes) -> int:
    c = 0xFFFFFFFF
    for b in data:
        c = _TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _masked(data: bytes) -> int:
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field_bytes(num: int, payload: bytes) -> bytes:
    return _varint((num << 3) | 2) + _varint(len(payload)) + payload


def _image_event(wall_time: float, step: int, tag: str, png: bytes, height: int, width: int) -> bytes:
    image = b"\x08" + _varint(height) + b"\x10" + _varint(width) + b"\x18" + _varint(3) + _field_bytes(4, png)
    value = _field_bytes(1, tag.encode()) + _field_bytes(4, image)
    return b"\x09" + struct.pack("<d", wall_time) + b"\x10" + _varint(step) + _field_bytes(5, _field_bytes(1, value))


def _event(wall_time: float, step: int = 0, file_version: str | None = None, scalars: dict | None = None) -> bytes:
    ev = b"\x09" + struct.pack("<d", wall_time) + b"\x10" + _varint(step)
    if file_version is not None:
        ev += _field_bytes(3, file_version.encode())
    if scalars:
        summary = b""
        for tag, val in scalars.items():
            value = _field_bytes(1, tag.encode()) + b"\x15" + struct.pack("<f", float(val))
            summary += _field_bytes(1, value)
        ev += _field_bytes(5, summary)
    return ev


class EventFileWriter:
    def __init__(self, logdir: str):
        os.makedirs(logdir, exist_ok=True)
        self.path = os.path.join(logdir, f"events.out.tfevents.{int(time.time())}.{socket.gethostname()}.dronecu")
        self._f = open(self.path, "wb")
        self._record(_event(time.time(), 0, file_version="brain.Event:2"))

    def _record(self, data: bytes):
        head = struct.pack("<Q", len(data))
        self._f.write(head + struct.pack("<I", _masked(head)) + data + struct.pack("<I", _masked(data)))

    def add_scalars(self, scalars: dict, step: int):
        nums = {k: v for k, v in scalars.items() if isinstance(v, (int, float)) and not isinstance(v, bool)}
        if nums:
            self._record(_event(time.time(), int(step), scalars=nums))
            self._f.flush()

    def add_image(self, tag: str, pil_image, step: int):
        """What SB3's ``writer.add_figure(tag, fig, step)`` stores (traj_tb.py:66): a PNG under an image summary."""
        import io
        b = io.BytesIO()
        pil_image.convert("RGB").save(b, format="PNG")
        self._record(_image_event(time.time(), int(step), tag, b.getvalue(), pil_image.height, pil_image.width))
        self._f.flush()

    def close(self):
        self._f.close()


def _read_varint(buf: bytes, i: int):
    n, shift = 0, 0
    while True:
        b = buf[i]
        i += 1
        n |= (b & 0x7F) << shift
        shift += 7
        if not b & 0x80:
            return n, i


def _parse(buf: bytes) -> dict:
    out, i = {}, 0
    while i < len(buf):
        key, i = _read_varint(buf, i)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _read_varint(buf, i)
        elif wt == 1:
            v, i = buf[i:i + 8], i + 8
        elif wt == 5:
            v, i = buf[i:i + 4], i + 4
        elif wt == 2:
            n, i = _read_varint(buf, i)
            v, i = buf[i:i + n], i + n
        else:
            raise ValueError(f"wire type {wt}")
        out.setdefault(num, []).append(v)
    return out


def read_events(path: str) -> list:
    """-> [{"wall_time", "step", "file_version"?, "scalars": {tag: value}}], CRCs checked."""
    data, i, events = open(path, "rb").read(), 0, []
    while i < len(data):
        head = data[i:i + 8]
        (n,) = struct.unpack("<Q", head)
        assert struct.unpack("<I", data[i + 8:i + 12])[0] == _masked(head), "length CRC"
        rec = data[i + 12:i + 12 + n]
        assert struct.unpack("<I", data[i + 12 + n:i + 16 + n])[0] == _masked(rec), "data CRC"
        i += 16 + n
        f = _parse(rec)
        ev = {"wall_time": struct.unpack("<d", f[1][0])[0], "step": f.get(2, [0])[0], "scalars": {}}
        if 3 in f:
            ev["file_version"] = f[3][0].decode()
        for summary in f.get(5, []):
            for value in _parse(summary).get(1, []):
                v = _parse(value)
                if 2 in v:
                    ev["scalars"][v[1][0].decode()] = struct.unpack("<f", v[2][0])[0]
                if 4 in v:
                    im = _parse(v[4][0])
                    ev.setdefault("images", {})[v[1][0].decode()] = {"height": im[1][0], "width": im[2][0], "png": im[4][0]}
        events.append(ev)
    return events
