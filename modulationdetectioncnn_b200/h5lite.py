"""Minimal read-only HDF5 parser for Keras 2.x ``*.wts.h5`` checkpoints.

The reference loads its checkpoints with ``model.load_weights(filepath)``
(/root/reference/cnn.py:147, CNN.ipynb cell 8) through h5py/libhdf5.  Neither is
available in this image, so this module reads the subset of the HDF5 file
format those five checkpoints use:

* superblock v0, 8-byte offsets and lengths;
* object headers v1 (+ continuation messages);
* "old style" groups: symbol-table message -> B-tree v1 + local heap -> SNOD;
* datasets: dataspace v1/v2, datatype classes 0 (int), 1 (float), 3 (fixed
  string), 9 (variable-length string), contiguous or compact layout;
* attributes v1..v3, variable-length strings resolved through global heaps.

Anything else (chunking, filters, new-style groups, ...) raises
``H5FormatError`` loudly rather than guessing.

Public surface::

    f = H5File(path)
    f.attrs("/")                     -> dict
    f.listdir("/model_weights")      -> [names]
    f.dataset("/model_weights/conv2d_3/conv2d_3/kernel:0") -> np.ndarray
    f.visit()                        -> iterator of (path, kind)
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Tuple

import numpy as np

__all__ = ["H5File", "H5FormatError"]

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    """The file uses an HDF5 feature this reader does not implement."""


@dataclass
class _Datatype:
    cls: int
    size: int
    np_dtype: Optional[np.dtype] = None
    # class 9 (vlen): True if vlen *string*
    vlen_string: bool = False
    base: Optional["_Datatype"] = None
    strpad: int = 0


@dataclass
class _Object:
    addr: int
    messages: List[Tuple[int, int, bytes]] = field(default_factory=list)  # (type, flags, body)


def _pad8(n: int) -> int:
    return (n + 7) & ~7


class H5File:
    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self._d = fh.read()
        self.path = path
        d = self._d
        if d[:8] != _SIG:
            raise H5FormatError(f"{path}: not an HDF5 file (bad signature)")
        ver = d[8]
        if ver not in (0, 1):
            raise H5FormatError(f"{path}: superblock version {ver} not supported (need 0/1)")
        self._osz, self._lsz = d[13], d[14]
        if (self._osz, self._lsz) != (8, 8):
            raise H5FormatError(f"{path}: offset/length sizes {self._osz}/{self._lsz} != 8/8")
        off = 24 if ver == 0 else 28
        self._base = struct.unpack_from("<Q", d, off)[0]
        # root symbol table entry follows base, freespace, eof, driver (4 x 8 bytes)
        ste = off + 32
        _link, self._root_addr, cache_type, _ = struct.unpack_from("<QQII", d, ste)
        self._gheap_cache: Dict[int, Dict[int, bytes]] = {}
        self._obj_cache: Dict[int, _Object] = {}

    # ---------------------------------------------------------------- low level
    def _u(self, fmt: str, off: int):
        return struct.unpack_from("<" + fmt, self._d, off)

    def _read_object(self, addr: int) -> _Object:
        if addr in self._obj_cache:
            return self._obj_cache[addr]
        d = self._d
        a = addr + self._base
        version = d[a]
        if version != 1:
            raise H5FormatError(f"object header v{version} at {addr} not supported (need v1)")
        nmsg = self._u("H", a + 2)[0]
        hsize = self._u("I", a + 8)[0]
        obj = _Object(addr)
        blocks = [(a + 16, hsize)]
        count = 0
        while blocks and count < nmsg:
            start, size = blocks.pop(0)
            p, end = start, start + size
            while p + 8 <= end and count < nmsg:
                mtype, msize, mflags = self._u("HHB", p)
                body = d[p + 8 : p + 8 + msize]
                p += 8 + msize
                count += 1
                if mtype == 0x10:  # continuation
                    coff, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((coff + self._base, clen))
                else:
                    obj.messages.append((mtype, mflags, body))
        self._obj_cache[addr] = obj
        return obj

    def _local_heap_data(self, heap_addr: int) -> int:
        a = heap_addr + self._base
        if self._d[a : a + 4] != b"HEAP":
            raise H5FormatError(f"bad local heap signature at {heap_addr}")
        return self._u("Q", a + 24)[0] + self._base

    def _cstr(self, off: int) -> str:
        end = self._d.index(b"\x00", off)
        return self._d[off:end].decode("utf-8")

    def _btree_group_entries(self, tree_addr: int, heap_data: int, out: Dict[str, int]):
        a = tree_addr + self._base
        d = self._d
        if d[a : a + 4] == b"SNOD":
            nsym = self._u("H", a + 6)[0]
            p = a + 8
            for _ in range(nsym):
                name_off, ohdr = self._u("QQ", p)
                out[self._cstr(heap_data + name_off)] = ohdr
                p += 40
            return
        if d[a : a + 4] != b"TREE":
            raise H5FormatError(f"bad B-tree node signature at {tree_addr}")
        ntype, level, used = self._u("BBH", a + 4)
        if ntype != 0:
            raise H5FormatError("B-tree node type %d (chunked data) not supported" % ntype)
        p = a + 24  # after sig(4) type(1) level(1) used(2) left(8) right(8)
        for i in range(used):
            p += 8  # key i
            child = self._u("Q", p)[0]
            p += 8
            self._btree_group_entries(child, heap_data, out)

    def _children(self, obj: _Object) -> Optional[Dict[str, int]]:
        for mtype, _f, body in obj.messages:
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", body, 0)
                out: Dict[str, int] = {}
                self._btree_group_entries(btree, self._local_heap_data(heap), out)
                return out
            if mtype in (0x02, 0x06):
                raise H5FormatError("new-style (link message) groups not supported")
        return None

    def _resolve(self, path: str) -> _Object:
        obj = self._read_object(self._root_addr)
        for part in [p for p in path.split("/") if p]:
            kids = self._children(obj)
            if kids is None or part not in kids:
                raise KeyError(f"{self.path}: no object {path!r} (missing {part!r})")
            obj = self._read_object(kids[part])
        return obj

    # ---------------------------------------------------------------- datatypes
    def _parse_datatype(self, b: bytes, off: int = 0) -> Tuple[_Datatype, int]:
        cv = b[off]
        cls, ver = cv & 0x0F, cv >> 4
        bits0, bits1, bits2 = b[off + 1], b[off + 2], b[off + 3]
        size = struct.unpack_from("<I", b, off + 4)[0]
        p = off + 8
        if cls == 0:  # fixed point
            order = ">" if bits0 & 1 else "<"
            signed = bool(bits0 & 8)
            dt = np.dtype(f"{order}{'i' if signed else 'u'}{size}")
            return _Datatype(cls, size, dt), p + 4
        if cls == 1:  # float
            order = ">" if bits0 & 1 else "<"
            if size not in (2, 4, 8):
                raise H5FormatError(f"float size {size}")
            return _Datatype(cls, size, np.dtype(f"{order}f{size}")), p + 12
        if cls == 3:  # fixed-length string
            return _Datatype(cls, size, np.dtype(f"S{size}"), strpad=bits0 & 0x0F), p
        if cls == 9:  # variable length
            vtype = bits0 & 0x0F  # 0 sequence, 1 string
            base, q = self._parse_datatype(b, p)
            return _Datatype(cls, size, None, vlen_string=(vtype == 1), base=base), q
        raise H5FormatError(f"datatype class {cls} not supported")

    @staticmethod
    def _parse_dataspace(b: bytes) -> Tuple[int, ...]:
        ver = b[0]
        rank = b[1]
        if ver == 1:
            p = 8
        elif ver == 2:
            if b[3] == 2:  # null dataspace
                return (0,)
            p = 4
        else:
            raise H5FormatError(f"dataspace version {ver}")
        return tuple(struct.unpack_from("<Q", b, p + 8 * i)[0] for i in range(rank))

    def _global_heap_object(self, coll_addr: int, index: int) -> bytes:
        if coll_addr not in self._gheap_cache:
            a = coll_addr + self._base
            d = self._d
            if d[a : a + 4] != b"GCOL":
                raise H5FormatError(f"bad global heap signature at {coll_addr}")
            csize = self._u("Q", a + 8)[0]
            objs: Dict[int, bytes] = {}
            p, end = a + 16, a + csize
            while p + 16 <= end:
                idx, _ref, _r, osize = self._u("HHIQ", p)
                if idx == 0:
                    break
                objs[idx] = d[p + 16 : p + 16 + osize]
                p += 16 + _pad8(osize)
            self._gheap_cache[coll_addr] = objs
        return self._gheap_cache[coll_addr][index]

    def _decode(self, dt: _Datatype, shape: Tuple[int, ...], raw: bytes):
        n = int(np.prod(shape)) if shape else 1
        if dt.cls == 9:
            if not dt.vlen_string:
                raise H5FormatError("vlen sequences not supported")
            vals = []
            for i in range(n):
                length, coll, idx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append(self._global_heap_object(coll, idx)[:length].decode("utf-8") if length else "")
            return vals[0] if shape == () else np.array(vals, dtype=object).reshape(shape)
        arr = np.frombuffer(raw, dtype=dt.np_dtype, count=n).reshape(shape)
        if dt.cls == 3:
            arr = np.array([s.rstrip(b"\x00 ").decode("utf-8") for s in arr.ravel()], dtype=object).reshape(shape)
            return arr[()] if shape == () else arr
        arr = arr.astype(dt.np_dtype.newbyteorder("="), copy=True)
        return arr[()] if shape == () else arr

    # ---------------------------------------------------------------- public
    def listdir(self, path: str = "/") -> List[str]:
        kids = self._children(self._resolve(path))
        if kids is None:
            raise KeyError(f"{path!r} is not a group")
        return sorted(kids)

    def is_group(self, path: str) -> bool:
        return self._children(self._resolve(path)) is not None

    def attrs(self, path: str = "/") -> Dict[str, object]:
        out: Dict[str, object] = {}
        for mtype, _f, body in self._resolve(path).messages:
            if mtype != 0x0C:
                continue
            ver = body[0]
            if ver == 1:
                nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
                p = 8
                name = body[p : p + nsz].split(b"\x00")[0].decode()
                p += _pad8(nsz)
                dt, _ = self._parse_datatype(body, p)
                p += _pad8(tsz)
                shape = self._parse_dataspace(body[p : p + ssz]) if ssz else ()
                p += _pad8(ssz)
            elif ver in (2, 3):
                nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
                p = 8 + (1 if ver == 3 else 0)
                name = body[p : p + nsz].split(b"\x00")[0].decode()
                p += nsz
                dt, _ = self._parse_datatype(body, p)
                p += tsz
                shape = self._parse_dataspace(body[p : p + ssz]) if ssz else ()
                p += ssz
            else:
                raise H5FormatError(f"attribute message version {ver}")
            out[name] = self._decode(dt, shape, body[p:])
        return out

    def dataset(self, path: str) -> np.ndarray:
        obj = self._resolve(path)
        dt = shape = None
        raw = None
        for mtype, _f, body in obj.messages:
            if mtype == 0x01:
                shape = self._parse_dataspace(body)
            elif mtype == 0x03:
                dt, _ = self._parse_datatype(body)
            elif mtype == 0x0B:
                raise H5FormatError(f"{path}: filtered (compressed) datasets not supported")
            elif mtype == 0x08:
                ver = body[0]
                if ver != 3:
                    raise H5FormatError(f"{path}: data layout version {ver} not supported (need 3)")
                lclass = body[1]
                if lclass == 1:
                    addr, size = struct.unpack_from("<QQ", body, 2)
                    raw = b"" if addr == _UNDEF else self._d[addr + self._base : addr + self._base + size]
                elif lclass == 0:
                    size = struct.unpack_from("<H", body, 2)[0]
                    raw = body[4 : 4 + size]
                else:
                    raise H5FormatError(f"{path}: chunked layout not supported")
        if dt is None or shape is None or raw is None:
            raise KeyError(f"{path!r} is not a dataset")
        return self._decode(dt, shape, raw)

    def visit(self, path: str = "/") -> Iterator[Tuple[str, str]]:
        kids = self._children(self._resolve(path))
        if kids is None:
            return
        for name in sorted(kids):
            full = path.rstrip("/") + "/" + name
            if self.is_group(full):
                yield full, "group"
                yield from self.visit(full)
            else:
                yield full, "dataset"
