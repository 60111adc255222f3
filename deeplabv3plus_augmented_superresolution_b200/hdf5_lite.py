"""Minimal HDF5 reader/writer for the augmented-copies hand-off file.

The reference writes one file per image with h5py (augmentation_utils.py:123-136) and reads it back
in load_SR_data (superres_utils.py:170-208).  h5py / libhdf5 are not available in this image, so this
module implements exactly the subset of the HDF5 file format that layout needs, following the public
"HDF5 File Format Specification Version 2.0/3.0" as h5py's defaults (libver='earliest') use it:

  superblock v0 . v1 object headers (+ continuation blocks) . old-style root group (symbol-table
  message -> v1 group B-tree -> SNOD nodes -> local heap) . contiguous (and compact) dataset layout .
  dataspace v1/v2 . fixed-point / IEEE float / fixed string / variable-length string datatypes .
  attribute messages v1-v3 . global heap collections for variable-length strings.

API = the h5py subset the reference touches: File(path, "r"|"w"), file[name].shape, file[name][:n],
iteration over dataset names, file.attrs[...], create_dataset(name, data=...), close().
Files written here use the same structures h5py writes, so h5py (where installed) reads them and
returns `str` for the two string attributes, which load_SR_data's `mode != "slice"` test relies on.
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class Hdf5FormatError(Exception):
    pass


# =====================================================================================================
# datatype / dataspace messages
# =====================================================================================================
def _decode_datatype(buf: bytes, off: int = 0):
    """-> (kind, numpy dtype | None, element size).  kind in {"num", "str", "vlen_str"}."""
    cv, b0, b1, _b2, size = struct.unpack_from("<BBBBI", buf, off)
    cls = cv & 0x0F
    if cls == 0:      # fixed point
        order = ">" if (b0 & 1) else "<"
        signed = bool(b0 & 0x08)
        return "num", np.dtype(f"{order}{'i' if signed else 'u'}{size}"), size
    if cls == 1:      # floating point
        order = ">" if (b0 & 1) else "<"
        return "num", np.dtype(f"{order}f{size}"), size
    if cls == 3:      # fixed-length string
        return "str", np.dtype(f"S{size}"), size
    if cls == 9:      # variable length
        if (b0 & 0x0F) != 1:
            raise Hdf5FormatError("variable-length sequences are not supported (only strings)")
        return "vlen_str", None, size
    raise Hdf5FormatError(f"unsupported datatype class {cls}")


def _encode_datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.kind == "f":
        size = dt.itemsize
        sign, exp_loc, exp_sz, man_sz, bias = {4: (31, 23, 8, 23, 127), 8: (63, 52, 11, 52, 1023)}[size]
        return struct.pack("<BBBBI", 0x11, 0x20, sign, 0, size) + struct.pack("<HHBBBBI", 0, size * 8, exp_loc, exp_sz, 0, man_sz, bias)
    if dt.kind in "iu":
        size = dt.itemsize
        return struct.pack("<BBBBI", 0x10, 0x08 if dt.kind == "i" else 0x00, 0, 0, size) + struct.pack("<HH", 0, size * 8)
    raise Hdf5FormatError(f"cannot encode dtype {dt}")


def _encode_vlen_str_datatype() -> bytes:
    # class 9 (variable length), type = string, null-terminated, UTF-8; base type = 1-byte unsigned integer
    base = struct.pack("<BBBBI", 0x10, 0x00, 0, 0, 1) + struct.pack("<HH", 0, 8)
    return struct.pack("<BBBBI", 0x19, 0x01, 0x01, 0, 16) + base


def _decode_dataspace(buf: bytes, off: int = 0):
    ver = buf[off]
    if ver == 1:
        rank, flags = buf[off + 1], buf[off + 2]
        p = off + 8
    elif ver == 2:
        rank, flags = buf[off + 1], buf[off + 2]
        if buf[off + 3] == 2:   # null dataspace
            return None
        p = off + 4
    else:
        raise Hdf5FormatError(f"dataspace version {ver}")
    return tuple(struct.unpack_from(f"<{rank}Q", buf, p)) if rank else ()


def _encode_dataspace(shape) -> bytes:
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", int(d)) for d in shape)


def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


# =====================================================================================================
# reader
# =====================================================================================================
class _Reader:
    def __init__(self, data: bytes):
        self.d = data
        if data[:8] != SIGNATURE:
            raise Hdf5FormatError("not an HDF5 file (bad signature)")
        ver = data[8]
        if ver not in (0, 1):
            raise Hdf5FormatError(f"superblock version {ver} not supported (h5py libver='earliest' writes 0)")
        if data[13] != 8 or data[14] != 8:
            raise Hdf5FormatError("only 8-byte offsets/lengths are supported")
        p = 24 + (4 if ver == 1 else 0)
        self.base, _fs, _eof, _drv = struct.unpack_from("<4Q", data, p)
        p += 32
        _name_off, self.root_header, cache, _r = struct.unpack_from("<QQII", data, p)
        self.root_scratch = struct.unpack_from("<QQ", data, p + 24) if cache == 1 else None

    # ---- object headers ------------------------------------------------------------------------------
    def messages(self, addr: int):
        """Yield (type, flags, body bytes) of a version-1 object header, following continuations."""
        d = self.d
        addr += self.base
        if d[addr:addr + 4] == b"OHDR":
            raise Hdf5FormatError("version-2 object headers (libver='latest') are not supported")
        ver, _res, nmsg, _ref, hsize = struct.unpack_from("<BBHII", d, addr)
        if ver != 1:
            raise Hdf5FormatError(f"object header version {ver}")
        blocks = [(addr + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and seen < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", d, p)
                body = d[p + 8:p + 8 + msize]
                p += 8 + msize
                seen += 1
                if mtype == 0x0010:      # continuation
                    coff, clen = struct.unpack_from("<QQ", body)
                    blocks.append((coff + self.base, clen))
                else:
                    yield mtype, mflags, body

    # ---- old-style groups ------------------------------------------------------------------------------
    def _heap_data(self, heap_addr: int):
        d = self.d
        a = heap_addr + self.base
        if d[a:a + 4] != b"HEAP":
            raise Hdf5FormatError("bad local heap signature")
        size, _free, seg = struct.unpack_from("<QQQ", d, a + 8)
        return d[seg + self.base: seg + self.base + size]

    def _walk_btree(self, addr: int, heap: bytes, out: "OrderedDict[str, int]"):
        d = self.d
        a = addr + self.base
        sig = d[a:a + 4]
        if sig == b"TREE":
            _ntype, level, used = struct.unpack_from("<BBH", d, a + 4)
            p = a + 24
            for i in range(used):
                child = struct.unpack_from("<Q", d, p + 8 + 16 * i)[0]
                self._walk_btree(child, heap, out)
        elif sig == b"SNOD":
            nsym = struct.unpack_from("<H", d, a + 6)[0]
            for i in range(nsym):
                noff, ohdr = struct.unpack_from("<QQ", d, a + 8 + 40 * i)
                name = heap[noff: heap.index(b"\0", noff)].decode("utf-8")
                out[name] = ohdr
        else:
            raise Hdf5FormatError(f"bad group node signature {sig!r}")

    def links(self, header_addr: int, scratch=None) -> "OrderedDict[str, int]":
        out: "OrderedDict[str, int]" = OrderedDict()
        btree = heap = None
        for mtype, _f, body in self.messages(header_addr):
            if mtype == 0x0011:
                btree, heap = struct.unpack_from("<QQ", body)
            elif mtype in (0x0002, 0x0006):
                raise Hdf5FormatError("new-style (link-message) groups are not supported")
        if btree is None and scratch:
            btree, heap = scratch
        if btree is not None:
            self._walk_btree(btree, self._heap_data(heap), out)
        return out

    # ---- values ---------------------------------------------------------------------------------------
    def _global_heap_object(self, coll_addr: int, index: int) -> bytes:
        d = self.d
        a = coll_addr + self.base
        if d[a:a + 4] != b"GCOL":
            raise Hdf5FormatError("bad global heap signature")
        csize = struct.unpack_from("<Q", d, a + 8)[0]
        p, end = a + 16, a + csize
        while p + 16 <= end:
            idx, _ref, _res, osize = struct.unpack_from("<HHIQ", d, p)
            if idx == index:
                return d[p + 16:p + 16 + osize]
            if idx == 0:
                break
            p += 16 + osize + (-osize % 8)
        raise Hdf5FormatError(f"global heap object {index} not found")

    def decode_values(self, kind, dt, esize, shape, raw: bytes):
        n = int(np.prod(shape)) if shape else 1
        if kind == "vlen_str":
            vals = []
            for i in range(n):
                _length, coll, idx = struct.unpack_from("<IQI", raw, 16 * i)
                vals.append("" if coll in (0, UNDEF) else self._global_heap_object(coll, idx).split(b"\0")[0].decode("utf-8"))
            return vals[0] if shape == () else np.array(vals, dtype=object).reshape(shape)
        arr = np.frombuffer(raw, dtype=dt, count=n)
        if kind == "str":
            arr = np.array([s.split(b"\0")[0] for s in arr.tolist()], dtype=dt)
        if shape == ():
            v = arr[0]
            return v.item() if kind == "num" else bytes(v)
        return arr.reshape(shape)

    def attributes(self, header_addr: int) -> dict:
        out = {}
        for mtype, _f, body in self.messages(header_addr):
            if mtype != 0x000C:
                continue
            ver = body[0]
            if ver == 1:
                nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
                p = 8
                name = body[p:p + nsz].split(b"\0")[0].decode("utf-8"); p += nsz + (-nsz % 8)
                tbuf = body[p:p + tsz]; p += tsz + (-tsz % 8)
                sbuf = body[p:p + ssz]; p += ssz + (-ssz % 8)
            elif ver in (2, 3):
                if body[1] & 3:
                    raise Hdf5FormatError("shared attribute datatypes are not supported")
                nsz, tsz, ssz = struct.unpack_from("<HHH", body, 2)
                p = 8 + (1 if ver == 3 else 0)
                name = body[p:p + nsz].split(b"\0")[0].decode("utf-8"); p += nsz
                tbuf = body[p:p + tsz]; p += tsz
                sbuf = body[p:p + ssz]; p += ssz
            else:
                raise Hdf5FormatError(f"attribute message version {ver}")
            kind, dt, esize = _decode_datatype(tbuf)
            shape = _decode_dataspace(sbuf)
            out[name] = None if shape is None else self.decode_values(kind, dt, esize, shape, body[p:])
        return out

    def dataset(self, header_addr: int):
        kind = dt = esize = shape = None
        layout = None
        for mtype, _f, body in self.messages(header_addr):
            if mtype == 0x0001:
                shape = _decode_dataspace(body)
            elif mtype == 0x0003:
                kind, dt, esize = _decode_datatype(body)
            elif mtype == 0x0008:
                ver = body[0]
                if ver == 3:
                    cls = body[1]
                    if cls == 1:
                        addr, size = struct.unpack_from("<QQ", body, 2)
                        layout = ("contiguous", addr, size)
                    elif cls == 0:
                        size = struct.unpack_from("<H", body, 2)[0]
                        layout = ("compact", body[4:4 + size], size)
                    else:
                        raise Hdf5FormatError("chunked datasets are not supported (the reference writes contiguous data)")
                elif ver in (1, 2):
                    rank, cls = body[1], body[2]
                    if cls != 1:
                        raise Hdf5FormatError("only contiguous layout is supported")
                    addr = struct.unpack_from("<Q", body, 8)[0]
                    layout = ("contiguous", addr, None)
                else:
                    raise Hdf5FormatError(f"data layout version {ver}")
            elif mtype == 0x000B:
                raise Hdf5FormatError("filtered (compressed) datasets are not supported")
        if shape is None or kind is None or layout is None:
            raise Hdf5FormatError("object is not a simple dataset")
        return kind, dt, esize, shape, layout


class Dataset:
    def __init__(self, rd: _Reader, header_addr: int):
        self._rd = rd
        self._kind, self.dtype, self._esize, self.shape, self._layout = rd.dataset(header_addr)
        self._attr_addr = header_addr

    def __len__(self):
        return self.shape[0]

    def _read(self, first_rows: int | None):
        shape = self.shape
        if first_rows is not None and shape:
            shape = (max(0, min(first_rows, shape[0])),) + tuple(shape[1:])
        n = int(np.prod(shape)) if shape else 1
        if self._layout[0] == "compact":
            raw = self._layout[1][: n * self._esize]
        else:
            addr = self._layout[1]
            if addr == UNDEF:
                raw = b"\0" * (n * self._esize)
            else:
                a = addr + self._rd.base
                raw = self._rd.d[a:a + n * self._esize]
        return self._rd.decode_values(self._kind, self.dtype, self._esize, shape, raw)

    def __getitem__(self, key):
        # the reference only ever does file[name][:num_aug]; read just those rows, then apply the key
        if isinstance(key, slice) and key.start in (None, 0) and key.step in (None, 1) and self.shape:
            stop = self.shape[0] if key.stop is None else (key.stop if key.stop >= 0 else self.shape[0] + key.stop)
            return np.array(self._read(stop))
        return np.array(self._read(None))[key]

    @property
    def attrs(self):
        return self._rd.attributes(self._attr_addr)


# =====================================================================================================
# writer
# =====================================================================================================
def _message(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages: list) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body


class _Writer:
    def __init__(self):
        self.chunks = []          # list of bytes, laid out back to back
        self.pos = 0

    def alloc(self, nbytes: int, align: int = 8) -> int:
        pad = -self.pos % align
        if pad:
            self.chunks.append(b"\0" * pad)
            self.pos += pad
        addr = self.pos
        self.chunks.append(None)   # placeholder
        self.pos += nbytes
        return addr

    def put(self, addr_index: int, data: bytes):
        self.chunks[addr_index] = data


def write_file(path: str, datasets: "OrderedDict[str, np.ndarray]", attrs: "OrderedDict[str, object]"):
    """One root group (symbol-table style) holding contiguous datasets and scalar root attributes."""
    if len(datasets) > 8:
        raise Hdf5FormatError("this writer emits a single symbol-table node: at most 8 datasets")
    names = sorted(datasets)                       # SNOD entries are ordered by name
    # ---- layout pass: superblock | root header | btree | heap | snod | gcol | dataset headers | raw data ----
    SB = 96
    # attribute messages (need the global-heap address for strings -> build the heap first)
    strings = [(k, v) for k, v in attrs.items() if isinstance(v, str)]
    gcol_objs = []
    for i, (_k, v) in enumerate(strings):
        gcol_objs.append(v.encode("utf-8"))
    # root object header size depends only on message sizes, which are known without addresses
    def attr_message(name: str, value, gcol_addr: int, str_index: dict) -> bytes:
        nm = name.encode("utf-8") + b"\0"
        if isinstance(value, str):
            tbuf = _encode_vlen_str_datatype()
            data = struct.pack("<IQI", len(value.encode("utf-8")), gcol_addr, str_index[name])
        elif isinstance(value, (bool, np.bool_)):
            raise Hdf5FormatError("boolean attributes are not supported")
        elif isinstance(value, (int, np.integer)):
            tbuf = _encode_datatype(np.int64)
            data = struct.pack("<q", int(value))
        elif isinstance(value, (float, np.floating)):
            tbuf = _encode_datatype(np.float64)
            data = struct.pack("<d", float(value))
        else:
            raise Hdf5FormatError(f"unsupported attribute type {type(value)}")
        sbuf = struct.pack("<BBBB4x", 1, 0, 0, 0)   # scalar dataspace
        body = struct.pack("<BBHHH", 1, 0, len(nm), len(tbuf), len(sbuf)) + _pad8(nm) + _pad8(tbuf) + _pad8(sbuf) + data
        return _message(0x000C, body)

    str_index = {k: i + 1 for i, (k, _v) in enumerate(strings)}
    dummy_attrs = [attr_message(k, v, 0, str_index) for k, v in attrs.items()]
    root_hdr_size = 16 + len(_message(0x0011, b"\0" * 16)) + sum(len(m) for m in dummy_attrs)

    BTREE_SIZE = 24 + (2 * 16 + 1) * 8 + 2 * 16 * 8
    SNOD_SIZE = 8 + 8 * 40
    heap_names = b"\0" * 8                         # offset 0: empty name of the root group
    name_off = {}
    for n in names:
        name_off[n] = len(heap_names)
        heap_names += _pad8(n.encode("utf-8") + b"\0")
    heap_data_size = max(88, len(heap_names) + 16)
    heap_data_size += -heap_data_size % 8

    pos = SB
    root_hdr = pos; pos += root_hdr_size; pos += -pos % 8
    btree = pos; pos += BTREE_SIZE
    heap = pos; pos += 32
    heap_seg = pos; pos += heap_data_size
    snod = pos; pos += SNOD_SIZE
    gcol = pos
    gcol_size = 0
    if strings:
        need = 16 + sum(16 + len(o) + (-len(o) % 8) for o in gcol_objs) + 16
        gcol_size = max(4096, need + (-need % 8))
        pos += gcol_size
    ds_hdr = {}
    ds_msgs = {}
    for n in names:
        arr = datasets[n]
        msgs = [
            _message(0x0001, _encode_dataspace(arr.shape)),
            _message(0x0003, _encode_datatype(arr.dtype), flags=1),
            _message(0x0005, struct.pack("<BBBB", 2, 2, 2, 0)),            # fill value v2: late alloc, write if set, undefined
            _message(0x0008, struct.pack("<BBQQ", 3, 1, 0, 0)),            # patched below
        ]
        ds_msgs[n] = msgs
        ds_hdr[n] = pos
        pos += 16 + sum(len(m) for m in msgs)
        pos += -pos % 8
    ds_data = {}
    for n in names:
        pos += -pos % 8
        ds_data[n] = pos
        pos += datasets[n].nbytes
    eof = pos

    # ---- emit ---------------------------------------------------------------------------------------------
    out = bytearray(eof)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", btree, heap)
    assert len(sb) == SB
    out[0:SB] = sb

    root_msgs = [_message(0x0011, struct.pack("<QQ", btree, heap))] + \
                [attr_message(k, v, gcol, str_index) for k, v in attrs.items()]
    rh = _object_header(root_msgs)
    assert len(rh) == root_hdr_size
    out[root_hdr:root_hdr + len(rh)] = rh

    bt = b"TREE" + struct.pack("<BBH", 0, 0, 1) + struct.pack("<QQ", UNDEF, UNDEF)
    bt += struct.pack("<QQQ", 0, snod, name_off[names[-1]] if names else 0)   # key0 (""), child0, key1 (largest name)
    out[btree:btree + len(bt)] = bt

    free_off = len(heap_names)
    hp = b"HEAP" + struct.pack("<B3x", 0) + struct.pack("<QQQ", heap_data_size, free_off, heap_seg)
    out[heap:heap + 32] = hp
    seg = bytearray(heap_data_size)
    seg[:len(heap_names)] = heap_names
    seg[free_off:free_off + 16] = struct.pack("<QQ", 1, heap_data_size - free_off)   # single free block, next = H5HL_FREE_NULL
    out[heap_seg:heap_seg + heap_data_size] = seg

    sn = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
    for n in names:
        sn += struct.pack("<QQII16x", name_off[n], ds_hdr[n], 0, 0)
    out[snod:snod + len(sn)] = sn

    if strings:
        g = b"GCOL" + struct.pack("<B3xQ", 1, gcol_size)
        for i, o in enumerate(gcol_objs):
            g += struct.pack("<HHIQ", i + 1, 1, 0, len(o)) + _pad8(o)
        g += struct.pack("<HHIQ", 0, 0, 0, gcol_size - len(g))    # free-space object spans the rest
        out[gcol:gcol + len(g)] = g

    for n in names:
        arr = np.ascontiguousarray(datasets[n])
        msgs = ds_msgs[n]
        msgs[3] = _message(0x0008, struct.pack("<BBQQ", 3, 1, ds_data[n], arr.nbytes))
        oh = _object_header(msgs)
        out[ds_hdr[n]:ds_hdr[n] + len(oh)] = oh
        out[ds_data[n]:ds_data[n] + arr.nbytes] = arr.tobytes()

    with open(path, "wb") as f:
        f.write(out)


# =====================================================================================================
# h5py-like front end
# =====================================================================================================
class _AttrsProxy(dict):
    pass


class File:
    """h5py.File look-alike for the subset the reference uses."""

    def __init__(self, path, mode="r"):
        self.filename = str(path)
        self.mode = mode
        if mode == "r":
            with open(self.filename, "rb") as f:
                self._rd = _Reader(f.read())
            self._links = self._rd.links(self._rd.root_header, self._rd.root_scratch)
            self.attrs = _AttrsProxy(self._rd.attributes(self._rd.root_header))
        elif mode == "w":
            self._pending = OrderedDict()
            self.attrs = _AttrsProxy()
        else:
            raise ValueError("mode must be 'r' or 'w'")
        self._open = True

    def __iter__(self):
        return iter(self._links if self.mode == "r" else self._pending)

    def keys(self):
        return list(iter(self))

    def __contains__(self, name):
        return name in (self._links if self.mode == "r" else self._pending)

    def __getitem__(self, name):
        if self.mode != "r":
            return self._pending[name]
        if name not in self._links:
            raise KeyError(f"Unable to open object (object '{name}' doesn't exist)")
        return Dataset(self._rd, self._links[name])

    def create_dataset(self, name, data=None, **_kw):
        if self.mode != "w":
            raise IOError("file is not open for writing")
        arr = np.asarray(data)
        if arr.dtype == np.float64 and isinstance(data, (list, tuple)) and len(data) and hasattr(data[0], "dtype"):
            arr = arr.astype(data[0].dtype)
        self._pending[name] = np.ascontiguousarray(arr)
        return self._pending[name]

    def close(self):
        if self._open and self.mode == "w":
            write_file(self.filename, self._pending, OrderedDict(self.attrs))
        self._open = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
