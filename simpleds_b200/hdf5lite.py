"""A small pure-Python HDF5 reader and writer: the file side of the DelaySpectrum path.

h5py / pyuvdata do not exist in this image, and the two file formats on either side of the hot path
are plain HDF5: UVH5 visibility files going in (delay_spectrum.py:861-1131 reads them
through pyuvdata) and the DelaySpectrum ``/Header`` + ``/Data`` layout going out
(delay_spectrum.py:2545-3194).  This module implements the slice of the HDF5 File Format
Specification (version 2.0 of the document) those files use:

* reading: superblock version 0, old-style groups (symbol table message: v1 B-tree + local heap +
  SNOD leaves), version-1 object headers with continuation blocks, dataspace / datatype / layout
  (compact, contiguous, chunked through a v1 chunk B-tree) / filter-pipeline / attribute messages,
  the LZF filter h5py registers as id 32000, datatypes: fixed point, floating point, fixed strings,
  the ``{r, i}`` compound h5py uses for complex numbers and the int8 enum it uses for bool;
* writing: the same structures with CONTIGUOUS dataset layouts (no filters), which is what makes
  ``write_partial`` a seek and a write: ``initialize_save_file`` lays the datasets out at their
  full size, later calls fill in hyperslabs in place.

Pinned by the reference's bundled ``simpleDS/data/*.uvh5`` files, all written by h5py (the reader
reproduces the literals the reference's tests assert on them, tests/test_uvh5_golden.py), and by
round trips through the writer (tests/test_hdf5lite.py).  Files written here follow the
specification's structures for a version-0 superblock; opening them with h5py itself could not be
tried in this image.
"""
import mmap
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
_SIG = b"\x89HDF\r\n\x1a\n"


# --------------------------------------------------------------------------------------- LZF
def lzf_decompress(src, out_len):
    out = bytearray(out_len)
    ip, op, n = 0, 0, len(src)
    while ip < n:
        ctrl = src[ip]
        ip += 1
        if ctrl < 32:
            ln = ctrl + 1
            out[op:op + ln] = src[ip:ip + ln]
            ip += ln
            op += ln
        else:
            ln = ctrl >> 5
            ref = op - ((ctrl & 0x1F) << 8) - 1
            if ln == 7:
                ln += src[ip]
                ip += 1
            ref -= src[ip]
            ip += 1
            ln += 2
            for _ in range(ln):  # may overlap
                out[op] = out[ref]
                op += 1
                ref += 1
    assert op == out_len, (op, out_len)
    return bytes(out)


# --------------------------------------------------------------------------------------- datatypes
def _encode_dtype(dt):
    """Datatype message body (version 1) of a numpy dtype."""
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits = 0x08 if dt.kind == "i" else 0x00  # little endian, zero padding, signed flag in bit 3
        return struct.pack("<BBBBI", 0x10 | 0, bits, 0, 0, dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f":
        size = dt.itemsize
        # class 1 (floating point) bit field: byte order LE, mantissa normalisation = 2 (implied msb), sign location
        sign = 8 * size - 1
        if size == 4:
            props = struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
        elif size == 8:
            props = struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        else:
            raise NotImplementedError(dt)
        return struct.pack("<BBBBI", 0x10 | 1, 0x20, sign, 0, size) + props
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x10 | 3, 0x00, 0, 0, dt.itemsize)  # null-terminated, ASCII
    if dt.kind == "c":  # h5py's complex: compound {r, i} of two floats
        half = np.dtype("<f" + str(dt.itemsize // 2))
        body = b""
        for name, off in ((b"r", 0), (b"i", dt.itemsize // 2)):
            body += name.ljust(8, b"\0") + struct.pack("<IB3xII4I", off, 0, 0, 0, 0, 0, 0, 0) + _encode_dtype(half)
        return struct.pack("<BBBBI", 0x10 | 6, 2, 0, 0, dt.itemsize) + body
    if dt.kind == "b":  # h5py's bool: enum over int8 with FALSE = 0, TRUE = 1
        base = _encode_dtype(np.int8)
        names = b"FALSE".ljust(8, b"\0") + b"TRUE".ljust(8, b"\0")
        return struct.pack("<BBBBI", 0x10 | 8, 2, 0, 0, 1) + base + names + bytes([0, 1])
    raise NotImplementedError(f"dtype {dt}")


def _decode_dtype(data):
    cls = data[0] & 0xF
    bits0 = data[1]
    size = struct.unpack_from("<I", data, 4)[0]
    if cls == 0:
        return np.dtype(("<" if not bits0 & 1 else ">") + ("i" if bits0 & 8 else "u") + str(size))
    if cls == 1:
        return np.dtype(("<" if not bits0 & 1 else ">") + "f" + str(size))
    if cls == 3:
        return np.dtype("S" + str(size))
    if cls == 6:  # compound {r, i} of two equal floats (the only compound on this path)
        assert size in (8, 16)
        return np.dtype("<c" + str(size))
    if cls == 8:  # enum over int8 (h5py bool)
        assert size == 1
        return np.dtype("bool")
    raise NotImplementedError(f"datatype class {cls}")


def _dtype_message_size(data):
    """Length of a version-1 datatype message body (needed to step through attribute messages)."""
    return len(data)


# --------------------------------------------------------------------------------------- reader
class File:
    """Read-only view of an HDF5 file (see the module docstring for what is understood)."""

    def __init__(self, path):
        with open(path, "rb") as f:  # mapped, not read: initialised save files are mostly unwritten extents
            self.b = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
        assert self.b[:8] == _SIG and self.b[8] == 0, "superblock v0 expected"
        assert self.b[13] == 8 and self.b[14] == 8
        # root symbol table entry at 56: name offset, header address, cache type, scratch
        _, hdr, _ = struct.unpack_from("<QQI", self.b, 56)
        self.root = hdr

    # ---- object headers ------------------------------------------------------------
    def messages(self, addr):
        b = self.b
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        assert ver == 1, "object header v1 expected"
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _ = struct.unpack_from("<HHB", b, pos)
                data = b[pos + 8:pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x10:
                    off, ln = struct.unpack_from("<QQ", data, 0)
                    blocks.append((off, ln))
                out.append((mtype, data))
        return out

    # ---- groups ----------------------------------------------------------------------
    def _heap_string(self, heap_addr, off):
        assert self.b[heap_addr:heap_addr + 4] == b"HEAP"
        data_addr = struct.unpack_from("<Q", self.b, heap_addr + 24)[0]
        s = data_addr + off
        return self.b[s:self.b.find(b"\0", s)].decode()

    def _walk_group_btree(self, addr, heap, out):
        b = self.b
        assert b[addr:addr + 4] == b"TREE"
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        assert ntype == 0
        pos = addr + 24
        for _ in range(used):
            child = struct.unpack_from("<Q", b, pos + 8)[0]  # key i (8), child i (8)
            pos += 16
            if level > 0:
                self._walk_group_btree(child, heap, out)
            else:
                assert b[child:child + 4] == b"SNOD"
                nsym = struct.unpack_from("<H", b, child + 6)[0]
                for k in range(nsym):
                    e = child + 8 + 40 * k
                    name_off, hdr = struct.unpack_from("<QQ", b, e)
                    out[self._heap_string(heap, name_off)] = hdr

    def children(self, addr):
        for mtype, data in self.messages(addr):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", data, 0)
                out = {}
                self._walk_group_btree(btree, heap, out)
                return out
        return {}

    def lookup(self, path):
        addr = self.root
        for part in [p for p in path.split("/") if p]:
            addr = self.children(addr)[part]
        return addr

    def keys(self, path="/"):
        return sorted(self.children(self.lookup(path)))

    def __contains__(self, path):
        try:
            self.lookup(path)
            return True
        except KeyError:
            return False

    # ---- datasets --------------------------------------------------------------------
    def _read_chunks(self, addr, rank, out, chunk_dims, dt, lzf):
        b = self.b
        assert b[addr:addr + 4] == b"TREE"
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        assert ntype == 1
        key_size = 8 + 8 * (rank + 1)
        pos = addr + 24
        for _ in range(used):
            csize, fmask = struct.unpack_from("<II", b, pos)
            offs = struct.unpack_from("<" + "Q" * (rank + 1), b, pos + 8)[:rank]
            child = struct.unpack_from("<Q", b, pos + key_size)[0]
            pos += key_size + 8
            if level > 0:
                self._read_chunks(child, rank, out, chunk_dims, dt, lzf)
                continue
            raw = b[child:child + csize]
            nbytes = int(np.prod(chunk_dims)) * dt.itemsize
            if lzf and not (fmask & 1):
                raw = lzf_decompress(raw, nbytes)
            chunk = np.frombuffer(raw[:nbytes], dtype=dt).reshape(chunk_dims)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk_dims, out.shape))
            out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]

    def _describe(self, path):
        msgs = self.messages(self.lookup(path))
        shape, dt, layout, lzf = (), None, None, False
        for mtype, data in msgs:
            if mtype == 0x1:
                ver, rank = data[0], data[1]
                off = 8 if ver == 1 else 4
                shape = struct.unpack_from("<" + "Q" * rank, data, off) if rank else ()
            elif mtype == 0x3:
                dt = _decode_dtype(data)
            elif mtype == 0x8:
                layout = data
            elif mtype == 0xB:
                nf = data[1]
                pos = 8
                for _ in range(nf):
                    fid, nlen, _, ncd = struct.unpack_from("<HHHH", data, pos)
                    pos += 8 + ((nlen + 7) // 8) * 8 + 4 * (ncd + (ncd & 1))
                    lzf = lzf or fid == 32000
        assert layout is not None and layout[0] == 3, "layout message v3 expected"
        return tuple(int(s) for s in shape), dt, layout, lzf, msgs

    def shape(self, path):
        return self._describe(path)[0]

    def read(self, path):
        shape, dt, layout, lzf, _ = self._describe(path)
        raw_dt = np.dtype("i1") if dt == np.dtype("bool") else dt
        cls = layout[1]
        n = int(np.prod(shape)) if shape else 1
        if cls == 0:
            size = struct.unpack_from("<H", layout, 2)[0]
            arr = np.frombuffer(bytes(layout[4:4 + size]), dtype=raw_dt, count=n)
        elif cls == 1:
            addr, size = struct.unpack_from("<QQ", layout, 2)
            arr = np.frombuffer(self.b, dtype=raw_dt, count=n, offset=addr) if addr != UNDEF else np.zeros(n, raw_dt)
        else:
            rank1 = layout[2]
            btree = struct.unpack_from("<Q", layout, 3)[0]
            dims = struct.unpack_from("<" + "I" * rank1, layout, 11)
            out = np.zeros(shape, dtype=raw_dt)
            if btree != UNDEF:
                self._read_chunks(btree, rank1 - 1, out, dims[:-1], raw_dt, lzf)
            return out.astype(dt) if dt != raw_dt else out
        arr = arr.reshape(shape) if shape else arr.reshape(())
        return arr.astype(dt) if dt != raw_dt else arr.copy()

    def contiguous_extent(self, path):
        """(file offset, nbytes) of a contiguous dataset (what ``write_partial`` seeks into)."""
        _, _, layout, _, _ = self._describe(path)
        assert layout[1] == 1, "contiguous layout expected"
        return struct.unpack_from("<QQ", layout, 2)

    def attrs(self, path):
        """Attributes of an object (version-1 attribute messages): name -> numpy value."""
        out = {}
        for mtype, data in self.messages(self.lookup(path)):
            if mtype != 0xC:
                continue
            ver, _, nlen, tlen, slen = struct.unpack_from("<BBHHH", data, 0)
            assert ver == 1, "attribute message v1 expected"
            pos = 8
            name = data[pos:pos + nlen].split(b"\0")[0].decode()
            pos += (nlen + 7) // 8 * 8
            dt = _decode_dtype(data[pos:pos + tlen])
            pos += (tlen + 7) // 8 * 8
            sp = data[pos:pos + slen]
            rank = sp[1]
            shape = struct.unpack_from("<" + "Q" * rank, sp, 8 if sp[0] == 1 else 4) if rank else ()
            pos += (slen + 7) // 8 * 8
            n = int(np.prod(shape)) if shape else 1
            val = np.frombuffer(bytes(data), dtype=dt, count=n, offset=pos)
            val = val.reshape(shape) if shape else val.reshape(())[()]
            if isinstance(val, bytes):  # fixed strings come back as str
                val = val.split(b"\0")[0].decode()
            out[name] = val
        return out


# --------------------------------------------------------------------------------------- writer
def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _dataspace(shape):
    shape = tuple(int(s) for s in shape)
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _message(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _attr_message(name, value):
    if isinstance(value, str):
        value = np.array(value.encode() + b"\0")
    value = np.asarray(value)
    if value.dtype.kind == "U":
        value = np.char.encode(value) if value.ndim else np.array(str(value).encode() + b"\0")
    nm = name.encode() + b"\0"
    dt, sp = _encode_dtype(value.dtype), _dataspace(value.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(sp)) + _pad8(nm) + _pad8(dt) + _pad8(sp) + value.tobytes()
    return _message(0xC, body)


class Writer:
    """Create an HDF5 file of groups and CONTIGUOUS datasets.

        w = Writer(path)
        w.create_dataset("Header/Ntimes", np.int64(3))
        w.create_dataset("Data/visdata", shape=(..), dtype=np.complex128, attrs={"unit": "Jy"})   # laid out, zero-filled
        w.close()
        write_hyperslab(path, "Data/visdata", (slice(..), ...), block)                           # later, in place

    Everything is kept in memory until ``close`` except the bulk data of datasets created from a shape
    (those are extents of zeros, written sparse)."""

    LEAF_K = 64  # symbol-table leaf nodes hold up to 2 K entries: one leaf per group on this path

    def __init__(self, path):
        self.path = path
        self.groups = {"": {}}   # group path -> {name: ("group", path) | ("dataset", record)}
        self.order = []

    def _group(self, gpath):
        if gpath not in self.groups:
            parent, _, name = gpath.rpartition("/")
            self._group(parent)[name] = ("group", gpath)
            self.groups[gpath] = {}
        return self.groups[gpath]

    def create_group(self, gpath):
        self._group(gpath.strip("/"))

    def create_dataset(self, path, data=None, shape=None, dtype=None, attrs=None):
        gpath, _, name = path.strip("/").rpartition("/")
        if data is not None:
            if isinstance(data, (str, bytes)):
                raw = data.encode() if isinstance(data, str) else data
                data = np.array(raw)  # fixed-length string, scalar dataspace (h5py: header[...] = b"...")
            data = np.ascontiguousarray(data) if np.ndim(data) else np.asarray(data)
            if data.dtype.kind == "U":
                data = np.char.encode(data)
            shape, dtype = data.shape, data.dtype
        rec = dict(shape=tuple(int(s) for s in shape), dtype=np.dtype(dtype), data=data, attrs=dict(attrs or {}))
        if name in self._group(gpath):
            raise ValueError(f"{path} exists")
        self._group(gpath)[name] = ("dataset", rec)

    # ---- layout --------------------------------------------------------------------------
    def close(self):
        buf = bytearray(b"\0" * 96)  # superblock (56 bytes) + root symbol table entry (40 bytes)
        tail = []                    # (offset, nbytes) of zero-filled extents past the in-memory image

        def alloc(nbytes, align=8):
            pos = (len(buf) + align - 1) // align * align
            buf.extend(b"\0" * (pos + nbytes - len(buf)))
            return pos

        def write_dataset(rec):
            nbytes = int(np.prod(rec["shape"], dtype=np.int64)) * rec["dtype"].itemsize if rec["shape"] else rec["dtype"].itemsize
            msgs = _message(0x1, _dataspace(rec["shape"])) + _message(0x3, _encode_dtype(rec["dtype"]), flags=1)
            msgs += _message(0x5, struct.pack("<BBBB", 2, 1, 0, 0))  # fill value v2: early allocation, undefined
            layout_pos = len(msgs) + 8 + 2  # where the data address sits inside the header block
            msgs += _message(0x8, struct.pack("<BBQQ", 3, 1, 0, nbytes))
            for k, v in rec["attrs"].items():
                msgs += _attr_message(k, v)
            hdr = alloc(16 + len(msgs))
            nmsg = 4 + len(rec["attrs"])
            struct.pack_into("<BBHII", buf, hdr, 1, 0, nmsg, 1, len(msgs))
            buf[hdr + 16:hdr + 16 + len(msgs)] = msgs
            rec["_addr_field"] = hdr + 16 + layout_pos
            rec["_nbytes"] = nbytes
            return hdr

        def write_group(gpath):
            members = self.groups[gpath]
            entries = []
            for name in sorted(members, key=lambda s: s.encode()):
                kind, val = members[name]
                entries.append((name, write_group(val) if kind == "group" else write_dataset(val)))
            if len(entries) > 2 * self.LEAF_K:
                raise NotImplementedError("more members than one symbol-table leaf holds")
            # local heap: offset 0 is the empty string; names null-terminated, 8-byte aligned
            heap_data = bytearray(b"\0" * 8)
            name_off = {}
            for name, _ in entries:
                name_off[name] = len(heap_data)
                heap_data += _pad8(name.encode() + b"\0")
            free_off = len(heap_data)
            heap_data += struct.pack("<QQ", 1, 16)  # one free block closing the heap (next = 1: last)
            heap_data_addr = alloc(len(heap_data))
            buf[heap_data_addr:heap_data_addr + len(heap_data)] = heap_data
            heap = alloc(32)
            buf[heap:heap + 32] = b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), free_off, heap_data_addr)
            snod = alloc(8 + 40 * 2 * self.LEAF_K)
            buf[snod:snod + 8] = b"SNOD" + struct.pack("<BBH", 1, 0, len(entries))
            for k, (name, addr) in enumerate(entries):
                struct.pack_into("<QQII16x", buf, snod + 8 + 40 * k, name_off[name], addr, 0, 0)
            tree = alloc(24 + 8 * (2 * 16 + 1) + 8 * 2 * 16)
            buf[tree:tree + 24] = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1 if entries else 0, UNDEF, UNDEF)
            if entries:
                struct.pack_into("<QQQ", buf, tree + 24, 0, snod, name_off[entries[-1][0]])
            msgs = _message(0x11, struct.pack("<QQ", tree, heap))
            hdr = alloc(16 + len(msgs))
            struct.pack_into("<BBHII", buf, hdr, 1, 0, 1, 1, len(msgs))
            buf[hdr + 16:hdr + 16 + len(msgs)] = msgs
            self._group_info[gpath] = (tree, heap)
            return hdr

        self._group_info = {}
        root = write_group("")
        # raw data after the metadata: in-memory arrays first, zero extents last (left sparse)
        recs = [v for g in self.groups.values() for k, (kind, v) in g.items() if kind == "dataset"]
        for rec in recs:
            if rec["data"] is not None:
                raw = rec["data"].astype(np.int8).tobytes() if rec["dtype"].kind == "b" else rec["data"].tobytes()
                pos = alloc(len(raw)) if raw else UNDEF
                if raw:
                    buf[pos:pos + len(raw)] = raw
                struct.pack_into("<Q", buf, rec["_addr_field"], pos)
        eof = len(buf)
        for rec in recs:
            if rec["data"] is None:
                pos = (eof + 7) // 8 * 8
                struct.pack_into("<Q", buf, rec["_addr_field"], pos if rec["_nbytes"] else UNDEF)
                eof = pos + rec["_nbytes"]
        # superblock v0
        tree, heap = self._group_info[""]
        sb = _SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree, heap)  # root entry, cached symbol table
        buf[:len(sb)] = sb
        with open(self.path, "wb") as f:
            f.write(buf)
            if eof > len(buf):
                f.truncate(eof)
        self.groups = None


def write_hyperslab(path, dataset, index, block):
    """Write ``block`` into ``dataset[index]`` of a file whose datasets are contiguous, in place.
    ``index``: a tuple of slices / integers (one per axis, leading axes first; missing axes are full)."""
    f = File(path)
    shape, dt, _, _, _ = f._describe(dataset)
    addr, _ = f.contiguous_extent(dataset)
    del f
    raw_dt = np.dtype("i1") if dt == np.dtype("bool") else dt
    block = np.asarray(block).astype(raw_dt, copy=False)
    index = tuple(index) + (slice(None),) * (len(shape) - len(index))
    starts, counts = [], []
    for ix, n in zip(index, shape):
        if isinstance(ix, slice):
            a, b, step = ix.indices(n)
            assert step == 1, "unit-stride hyperslabs only"
            starts.append(a)
            counts.append(b - a)
        else:
            starts.append(int(ix))
            counts.append(1)
    block = block.reshape(counts)
    # contiguous runs: all trailing axes that are taken whole, times the innermost partial axis
    k = len(shape)
    while k > 0 and counts[k - 1] == shape[k - 1]:
        k -= 1
    run_axes = max(k - 1, 0)  # axes [run_axes, ...) form one contiguous run per outer index
    run = int(np.prod(counts[run_axes:], dtype=np.int64)) if shape else 1
    strides = [int(np.prod(shape[a + 1:], dtype=np.int64)) for a in range(len(shape))]
    flat = np.ascontiguousarray(block).reshape(-1, run) if run else block.reshape(0, 0)
    with open(path, "r+b") as fh:
        for n, outer in enumerate(np.ndindex(*counts[:run_axes])):
            off = sum((s + o) * st for s, o, st in zip(starts[:run_axes], outer, strides[:run_axes]))
            off += sum(s * st for s, st in zip(starts[run_axes:], strides[run_axes:]))
            fh.seek(addr + off * raw_dt.itemsize)
            fh.write(flat[n].tobytes())
