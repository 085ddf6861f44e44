"""Reader/writer for the DualSPHysics ``.bi4`` binary container (JBinaryData).

Format contract restated from the reference (nothing is copied; all file:line
citations are relative to /root/reference/src/source):

* 64-byte file header ``"#FileJBD <filecode>"`` padded with spaces to 58 bytes,
  ``'\\n'`` at [58], 0 at [59], byteorder byte (0 = little endian, +10 when
  64-bit sizes are used), si64 byte, two unused bytes
  (JBinaryData.h:186-192, JBinaryData.cpp:1283-1299).
* One root ITEM, written recursively (JBinaryData.cpp:916-966):
  ``uint sizeitemdef`` + itemdef{str "\\nITEM\\n", str name, bool hide,
  bool hidevalues, str fmtfloat, str fmtdouble, uint narrays, uint nitems,
  uint sizevalues} + VALUES block {str "\\nVALUES", uint n, n x (str name,
  int type, payload)} (JBinaryData.cpp:792-816,1095-1101) + arrays + child items.
* ARRAY: ``uint sizearraydef`` + {str "\\nARRAY", str name, bool hide, int type,
  uint|ullong count, uint|ullong sizedata} + raw data (JBinaryData.cpp:850-910).
* Strings are ``uint len`` + bytes (JBinaryData.cpp:779-786); bools are int32
  (JBinaryData.h:252); type ids JBinaryData.h:83-87.

This module is host-side tooling (case writer, PART reader for the tests); the
solver's own loader/saver is the C++ code in ``csrc/host``.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Any

import numpy as np

# JBinaryDataDef::TpData (JBinaryData.h:83-87)
DAT_TEXT, DAT_BOOL, DAT_CHAR, DAT_UCHAR, DAT_SHORT, DAT_USHORT = 1, 2, 3, 4, 5, 6
DAT_INT, DAT_UINT, DAT_LLONG, DAT_ULLONG, DAT_FLOAT, DAT_DOUBLE = 7, 8, 9, 10, 11, 12
DAT_INT3, DAT_UINT3, DAT_FLOAT3, DAT_DOUBLE3 = 20, 21, 22, 23

_SCALAR_FMT = {
    DAT_CHAR: "<b", DAT_UCHAR: "<B", DAT_SHORT: "<h", DAT_USHORT: "<H",
    DAT_INT: "<i", DAT_UINT: "<I", DAT_LLONG: "<q", DAT_ULLONG: "<Q",
    DAT_FLOAT: "<f", DAT_DOUBLE: "<d",
    DAT_INT3: "<3i", DAT_UINT3: "<3I", DAT_FLOAT3: "<3f", DAT_DOUBLE3: "<3d",
}
_NP_TYPE = {
    DAT_CHAR: (np.int8, 1), DAT_UCHAR: (np.uint8, 1), DAT_SHORT: (np.int16, 1),
    DAT_USHORT: (np.uint16, 1), DAT_INT: (np.int32, 1), DAT_UINT: (np.uint32, 1),
    DAT_LLONG: (np.int64, 1), DAT_ULLONG: (np.uint64, 1), DAT_FLOAT: (np.float32, 1),
    DAT_DOUBLE: (np.float64, 1), DAT_INT3: (np.int32, 3), DAT_UINT3: (np.uint32, 3),
    DAT_FLOAT3: (np.float32, 3), DAT_DOUBLE3: (np.float64, 3),
}
MAX_SIZE_SI32 = 4228102018  # JBinaryData.h: MaxSizeSi32


@dataclass
class Item:
    """One JBinaryData item: ordered values, named arrays, child items."""
    name: str
    values: list = field(default_factory=list)   # [(name, type, value)]
    arrays: list = field(default_factory=list)   # [(name, type, ndarray)]
    items: list = field(default_factory=list)
    hide: bool = False
    hidevalues: bool = False
    fmtfloat: str = "%.7E"
    fmtdouble: str = "%.15E"

    # -- convenience -------------------------------------------------------
    def set(self, name: str, typ: int, value: Any) -> None:
        for i, (n, _, _) in enumerate(self.values):
            if n == name:
                self.values[i] = (name, typ, value)
                return
        self.values.append((name, typ, value))

    def get(self, name: str, default: Any = None) -> Any:
        for n, _, v in self.values:
            if n == name:
                return v
        return default

    def array(self, name: str):
        for n, _, a in self.arrays:
            if n == name:
                return a
        return None

    def item(self, name: str):
        for it in self.items:
            if it.name == name:
                return it
        return None

    def add_array(self, name: str, typ: int, data) -> None:
        dt, k = _NP_TYPE[typ]
        a = np.ascontiguousarray(data, dtype=dt)
        if k == 3:
            a = a.reshape(-1, 3)
        self.arrays.append((name, typ, a))


def _s(b: str) -> bytes:
    e = b.encode("latin-1")
    return struct.pack("<I", len(e)) + e


def _value_bytes(name: str, typ: int, v: Any) -> bytes:
    out = _s(name) + struct.pack("<i", typ)
    if typ == DAT_TEXT:
        return out + _s(v)
    if typ == DAT_BOOL:
        return out + struct.pack("<i", 1 if v else 0)
    fmt = _SCALAR_FMT[typ]
    if typ >= DAT_INT3:
        return out + struct.pack(fmt, *v)
    return out + struct.pack(fmt, v)


def _values_block(it: Item) -> bytes:
    b = _s("\nVALUES") + struct.pack("<I", len(it.values))
    for n, t, v in it.values:
        b += _value_bytes(n, t, v)
    return b


def _array_bytes(name: str, typ: int, a: np.ndarray, si64: bool) -> list:
    count = a.shape[0]
    raw = a.tobytes()
    base = _s("\nARRAY") + _s(name) + struct.pack("<i", 0) + struct.pack("<i", typ)
    if si64:
        base += struct.pack("<QQ", count, len(raw))
    else:
        base += struct.pack("<II", count, len(raw))
    return [struct.pack("<I", len(base)), base, raw]


def _item_chunks(it: Item, si64: bool) -> list:
    vals = _values_block(it)
    base = (_s("\nITEM\n") + _s(it.name) + struct.pack("<i", int(it.hide))
            + struct.pack("<i", int(it.hidevalues)) + _s(it.fmtfloat) + _s(it.fmtdouble)
            + struct.pack("<III", len(it.arrays), len(it.items), len(vals)))
    chunks = [struct.pack("<I", len(base)), base, vals]
    for n, t, a in it.arrays:
        chunks += _array_bytes(n, t, a, si64)
    for ch in it.items:
        chunks += _item_chunks(ch, si64)
    return chunks


def save(path: str, root: Item, filecode: str | None = None) -> None:
    """Write ``root`` as a bi4 file (JBinaryData::SaveFile, JBinaryData.cpp:1671)."""
    filecode = root.name if filecode is None else filecode
    chunks = _item_chunks(root, False)
    if sum(len(c) for c in chunks) >= MAX_SIZE_SI32:
        chunks = _item_chunks(root, True)
        si64 = True
    else:
        si64 = False
    titu = ("#FileJBD " + filecode)[:58].ljust(58, " ").encode("latin-1") + b"\n\x00"
    head = titu + bytes([10 if si64 else 0, 1 if si64 else 0, 0, 0])
    assert len(head) == 64
    with open(path, "wb") as f:
        f.write(head)
        for c in chunks:
            f.write(c)


class _Reader:
    def __init__(self, buf: bytes, si64: bool):
        self.b, self.p, self.si64 = buf, 0, si64

    def u32(self) -> int:
        v = struct.unpack_from("<I", self.b, self.p)[0]
        self.p += 4
        return v

    def i32(self) -> int:
        v = struct.unpack_from("<i", self.b, self.p)[0]
        self.p += 4
        return v

    def u64(self) -> int:
        v = struct.unpack_from("<Q", self.b, self.p)[0]
        self.p += 8
        return v

    def s(self) -> str:
        n = self.u32()
        v = self.b[self.p:self.p + n].decode("latin-1")
        self.p += n
        return v

    def value(self):
        name = self.s()
        typ = self.i32()
        if typ == DAT_TEXT:
            return name, typ, self.s()
        if typ == DAT_BOOL:
            return name, typ, self.i32() != 0
        fmt = _SCALAR_FMT[typ]
        v = struct.unpack_from(fmt, self.b, self.p)
        self.p += struct.calcsize(fmt)
        return name, typ, (v if typ >= DAT_INT3 else v[0])

    def item(self) -> Item:
        self.u32()  # sizeitemdef
        code = self.s()
        if code != "\nITEM\n":
            raise ValueError("bi4: bad ITEM code")
        it = Item(self.s())
        it.hide = self.i32() != 0
        it.hidevalues = self.i32() != 0
        it.fmtfloat = self.s()
        it.fmtdouble = self.s()
        narrays, nitems, sizevalues = self.u32(), self.u32(), self.u32()
        if sizevalues:
            end = self.p + sizevalues
            if self.s() != "\nVALUES":
                raise ValueError("bi4: bad VALUES code")
            for _ in range(self.u32()):
                it.values.append(self.value())
            self.p = end
        for _ in range(narrays):
            self.u32()  # sizearraydef
            if self.s() != "\nARRAY":
                raise ValueError("bi4: bad ARRAY code")
            name = self.s()
            self.i32()  # hide
            typ = self.i32()
            if self.si64:
                count, size = self.u64(), self.u64()
            else:
                count, size = self.u32(), self.u32()
            dt, k = _NP_TYPE[typ]
            a = np.frombuffer(self.b, dtype=dt, count=count * k, offset=self.p).copy()
            if k == 3:
                a = a.reshape(-1, 3)
            self.p += size
            it.arrays.append((name, typ, a))
        for _ in range(nitems):
            it.items.append(self.item())
        return it


def load(path: str) -> Item:
    """Read a bi4 file into an :class:`Item` tree (JBinaryData::LoadFile)."""
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 64 or not buf.startswith(b"#FileJBD "):
        raise ValueError(f"{path}: not a JBinaryData file")
    si64 = buf[61] != 0
    bo = buf[60] - (10 if si64 else 0)
    if bo != 0:
        raise ValueError(f"{path}: byte order not supported")
    return _Reader(buf[64:], si64).item()


def load_list(path: str) -> list:
    """Read a list file (JBinaryData::SaveFileListApp / LoadFileListApp, JBinaryData.cpp:1736-1810): the head item followed
    by the items appended to it, one per saved PART (PartFloatInfo.ibi4, PartOut_000.obi4, PartMotionRef.ibi4, PartInfo.ibi4)."""
    with open(path, "rb") as f:
        buf = f.read()
    if len(buf) < 64 or not buf.startswith(b"#FileJBD "):
        raise ValueError(f"{path}: not a JBinaryData file")
    si64 = buf[61] != 0
    rd = _Reader(buf[64:], si64)
    out = []
    while rd.p < len(rd.b):
        out.append(rd.item())
    return out


# ---------------------------------------------------------------------------
# JPartDataBi4 helpers (JPartDataBi4.cpp:190-390)
# ---------------------------------------------------------------------------
def make_case_item(casename: str, *, np_total: int, nfixed: int, nmoving: int, nfloat: int,
                   nfluid: int, posmin, posmax, dp: float, h: float, b: float, rhop0: float,
                   gamma: float, massbound: float, massfluid: float, data2d: bool = False,
                   data2dposy: float = 0.0, appname: str = "b200sph-casegen") -> Item:
    """Root item of a ``Case.bi4`` with the head values the reference writes."""
    d = Item("JPartDataBi4")
    d.set("Piece", DAT_UINT, 0)
    d.set("Npiece", DAT_UINT, 1)
    d.set("RunCode", DAT_TEXT, "00000000")
    d.set("Date", DAT_TEXT, "???")
    d.set("AppName", DAT_TEXT, appname)
    d.set("CaseName", DAT_TEXT, casename)
    d.set("Data2d", DAT_BOOL, data2d)
    d.set("Data2dPosY", DAT_DOUBLE, data2dposy)
    d.set("MapPosMin", DAT_DOUBLE3, (0.0, 0.0, 0.0))
    d.set("MapPosMax", DAT_DOUBLE3, (0.0, 0.0, 0.0))
    d.set("PeriMode", DAT_INT, 96)  # PERI_Unknown (JPeriodicDef.h:42)
    d.set("PeriXinc", DAT_DOUBLE3, (0.0, 0.0, 0.0))
    d.set("PeriYinc", DAT_DOUBLE3, (0.0, 0.0, 0.0))
    d.set("PeriZinc", DAT_DOUBLE3, (0.0, 0.0, 0.0))
    d.set("AxisDiv", DAT_INT, 99)  # DIV_Unknown (JPartDataBi4.h:84)
    d.set("CaseNp", DAT_ULLONG, np_total)
    d.set("CaseNfixed", DAT_ULLONG, nfixed)
    d.set("CaseNmoving", DAT_ULLONG, nmoving)
    d.set("CaseNfloat", DAT_ULLONG, nfloat)
    d.set("CaseNfluid", DAT_ULLONG, nfluid)
    d.set("CasePosMin", DAT_DOUBLE3, tuple(float(x) for x in posmin))
    d.set("CasePosMax", DAT_DOUBLE3, tuple(float(x) for x in posmax))
    d.set("NpDynamic", DAT_BOOL, False)
    d.set("ReuseIds", DAT_BOOL, False)
    d.set("Dp", DAT_DOUBLE, dp)
    d.set("H", DAT_DOUBLE, h)
    d.set("B", DAT_DOUBLE, b)
    d.set("Rhop0", DAT_DOUBLE, rhop0)
    d.set("Gamma", DAT_DOUBLE, gamma)
    d.set("MassBound", DAT_DOUBLE, massbound)
    d.set("MassFluid", DAT_DOUBLE, massfluid)
    d.set("Symmetry", DAT_BOOL, False)
    d.set("Splitting", DAT_BOOL, False)
    return d


def add_part(root: Item, cpart: int, *, timestep: float, idp, pos, vel, rhop, nout: int = 0,
             step: int = 0, posdouble: bool = True, boundnor=None, domainmin=(0, 0, 0),
             domainmax=(0, 0, 0), extra_values=()) -> Item:
    """Append a ``PART_%04u`` child with the arrays of JPartDataBi4::AddPartData."""
    idp = np.asarray(idp, dtype=np.uint32)
    n = int(idp.shape[0])
    p = Item("PART_%04u" % cpart)
    p.set("Cpart", DAT_UINT, cpart)
    p.set("TimeStep", DAT_DOUBLE, float(timestep))
    p.set("Npok", DAT_UINT, n)
    p.set("Nout", DAT_UINT, nout)
    p.set("Step", DAT_UINT, step)
    p.set("RunTime", DAT_DOUBLE, 0.0)
    p.set("DomainMin", DAT_DOUBLE3, tuple(float(x) for x in domainmin))
    p.set("DomainMax", DAT_DOUBLE3, tuple(float(x) for x in domainmax))
    for nm, t, v in extra_values:
        p.set(nm, t, v)
    p.add_array("Idp", DAT_UINT, idp)
    if posdouble:
        p.add_array("Posd", DAT_DOUBLE3, pos)
    else:
        p.add_array("Pos", DAT_FLOAT3, pos)
    p.add_array("Vel", DAT_FLOAT3, vel)
    p.add_array("Rhop", DAT_FLOAT, rhop)
    if boundnor is not None:
        p.add_array("BoundNor", DAT_FLOAT3, boundnor)
    root.items.append(p)
    return p


def load_part(path: str) -> dict:
    """Load a Case.bi4 / Part_XXXX.bi4 into a dict of numpy arrays + head values."""
    root = load(path)
    part = next((it for it in root.items if it.name.startswith("PART_")), None)
    if part is None:
        raise ValueError(f"{path}: no PART item")
    out = {"head": {n: v for n, _, v in root.values}, "part": {n: v for n, _, v in part.values}}
    for n, _, a in part.arrays:
        out[n] = a
    if "Posd" in out:
        out["pos"] = out["Posd"]
    elif "Pos" in out:
        out["pos"] = out["Pos"].astype(np.float64)
    return out
