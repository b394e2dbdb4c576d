"""A throw-away pure-Python reader for the old-style HDF5 files bundled with the reference
(simpleDS/data/*.uvh5): superblock v0, symbol-table groups (B-tree v1 + local heap), v1
object headers, contiguous / chunked (B-tree v1) layouts, the LZF filter (h5py id 32000).
h5py does not exist in this image.  Used only by make_uvh5_golden.py to turn two bundled
files into npz fixtures; not part of the product and not needed on the GPU box."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


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


class File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        assert self.b[:8] == b"\x89HDF\r\n\x1a\n" and self.b[8] == 0, "superblock v0 expected"
        assert self.b[13] == 8 and self.b[14] == 8
        # root symbol table entry at 56: name offset, header address, cache type, scratch
        _, hdr, ctype = struct.unpack_from("<QQI", self.b, 56)
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
        return self.b[s:self.b.index(b"\0", s)].decode()

    def _walk_group_btree(self, addr, heap, out):
        b = self.b
        assert b[addr:addr + 4] == b"TREE"
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        assert ntype == 0
        pos = addr + 24
        for i in range(used):
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

    # ---- datasets --------------------------------------------------------------------
    @staticmethod
    def _dtype(data):
        cls, ver = data[0] & 0xF, data[0] >> 4
        bits0 = data[1]
        size = struct.unpack_from("<I", data, 4)[0]
        if cls == 0:
            return np.dtype(("<" if not bits0 & 1 else ">") + ("i" if bits0 & 8 else "u") + str(size))
        if cls == 1:
            return np.dtype(("<" if not bits0 & 1 else ">") + "f" + str(size))
        if cls == 3:
            return np.dtype("S" + str(size))
        if cls == 6:  # compound {r, i} of two equal floats (the only compound in these files)
            assert size in (8, 16)
            return np.dtype("<c" + str(size))
        if cls == 8:  # enum over int8 (h5py bool)
            assert size == 1
            return np.dtype("i1")
        raise NotImplementedError(f"datatype class {cls}")

    def _read_chunks(self, addr, rank, out, chunk_dims, dt, lzf):
        b = self.b
        assert b[addr:addr + 4] == b"TREE"
        ntype, level, used = struct.unpack_from("<BBH", b, addr + 4)
        assert ntype == 1
        key_size = 8 + 8 * (rank + 1)
        pos = addr + 24
        for i in range(used):
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

    def read(self, path):
        msgs = self.messages(self.lookup(path))
        shape, dt, layout, lzf = (), None, None, False
        for mtype, data in msgs:
            if mtype == 0x1:
                ver, rank, flags = data[0], data[1], data[2]
                off = 8 if ver == 1 else 4
                shape = struct.unpack_from("<" + "Q" * rank, data, off) if rank else ()
            elif mtype == 0x3:
                dt = self._dtype(data)
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
        cls = layout[1]
        n = int(np.prod(shape)) if shape else 1
        if cls == 0:
            size = struct.unpack_from("<H", layout, 2)[0]
            arr = np.frombuffer(layout[4:4 + size], dtype=dt, count=n)
        elif cls == 1:
            addr, size = struct.unpack_from("<QQ", layout, 2)
            arr = np.frombuffer(self.b, dtype=dt, count=n, offset=addr) if addr != UNDEF else np.zeros(n, dt)
        else:
            rank1 = layout[2]
            btree = struct.unpack_from("<Q", layout, 3)[0]
            dims = struct.unpack_from("<" + "I" * rank1, layout, 11)
            out = np.zeros(shape, dtype=dt)
            if btree != UNDEF:
                self._read_chunks(btree, rank1 - 1, out, dims[:-1], dt, lzf)
            return out
        arr = arr.reshape(shape) if shape else arr.reshape(())
        return arr.copy()
