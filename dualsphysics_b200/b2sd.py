"""Reader/writer of the B2SD dump container used by the parity fixtures.

Layout (written by oracle/ref_probe.cpp and by tools/make_golden.py): ``B2SDUMP1``, u32 nentries, then
per entry ``char name[24]; u32 dtype (0=u8,1=u16,2=u32,3=i32,4=f32,5=f64); u64 count; raw data``.
"""
from __future__ import annotations

import struct

import numpy as np

_DT = [np.uint8, np.uint16, np.uint32, np.int32, np.float32, np.float64]
_CODE = {np.dtype(d): i for i, d in enumerate(_DT)}


def load(path: str) -> dict:
    with open(path, "rb") as f:
        buf = f.read()
    if buf[:8] != b"B2SDUMP1":
        raise ValueError(f"{path}: not a B2SD file")
    n = struct.unpack_from("<I", buf, 8)[0]
    p = 12
    out = {}
    for _ in range(n):
        name = buf[p:p + 24].split(b"\0", 1)[0].decode()
        dtype, count = struct.unpack_from("<IQ", buf, p + 24)
        p += 36
        dt = np.dtype(_DT[dtype])
        a = np.frombuffer(buf, dtype=dt, count=count, offset=p).copy()
        p += count * dt.itemsize
        out[name] = a[0] if count == 1 else a
    return out


def save(path: str, entries: dict) -> None:
    with open(path, "wb") as f:
        f.write(b"B2SDUMP1")
        f.write(struct.pack("<I", len(entries)))
        for name, v in entries.items():
            a = np.ascontiguousarray(v)
            if a.dtype == np.int64:
                a = a.astype(np.uint32)
            a = a.reshape(-1)
            f.write(name.encode()[:23].ljust(24, b"\0"))
            f.write(struct.pack("<IQ", _CODE[a.dtype], a.size))
            f.write(a.tobytes())
