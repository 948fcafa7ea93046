"""Minimal pure-Python HDF5 reader of the host mirror: JuliaGrid case files (`powerSystem("case.h5")`,
src/powerSystem/load.jl:141-289) without h5py, which this image does not have. In the Julia drop-in the files are read
by JuliaGrid itself; the oracle keeps its own copy of this reader so that it never depends on the product.
This reader understands exactly what those files use (written by HDF5.jl / libhdf5 1.12+):

* superblock v0, 8-byte offsets/lengths
* v1 object headers (+ continuation blocks)
* old-style groups (symbol-table message -> v1 B-tree + local heap + SNOD)
* new-style groups: compact link messages (0x06) and dense storage
  (link-info message 0x02 -> fractal heap, direct or indirect root block)
* datasets: contiguous / compact layout v3, fixed-point and IEEE float types, scalar or simple dataspaces
* attributes (message 0x0C, v1-v3) with the same numeric types

Variable-length strings (labels) are not decoded; such datasets return None.
The reference's own reader for this format is `src/powerSystem/load.jl:141-289, 1360-1368`
(HDF5.jl does the parsing there); SURVEY.md Appendix C records the layout.
"""
from __future__ import annotations

import struct
import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path: str):
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        ver = b[8]
        if ver not in (0, 1):
            raise ValueError(f"unsupported superblock version {ver}")
        if b[13] != 8 or b[14] != 8:
            raise ValueError("only 8-byte offsets/lengths supported")
        off = 24 if ver == 0 else 28
        self.base = struct.unpack_from("<Q", b, off)[0]
        ste = off + 32
        self.root_addr = struct.unpack_from("<Q", b, ste + 8)[0]
        self._cache: dict[int, dict] = {}

    # ------------------------------------------------------------------ object headers
    def _messages(self, addr: int):
        b = self.buf
        ver = b[addr]
        if ver != 1:
            if b[addr:addr + 4] == b"OHDR":
                return self._messages_v2(addr)
            raise ValueError(f"unsupported object header version {ver} at {addr}")
        nmsg = struct.unpack_from("<H", b, addr + 2)[0]
        hsize = struct.unpack_from("<I", b, addr + 8)[0]
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                body = p + 8
                if mtype == 0x10:
                    caddr, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((caddr, clen))
                out.append((mtype, body, msize, mflags))
                p = body + msize
        return out

    def _messages_v2(self, addr: int):
        b = self.buf
        flags = b[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        szf = 1 << (flags & 3)
        size0 = int.from_bytes(b[p:p + szf], "little")
        p += szf
        track = bool(flags & 0x04)
        blocks = [(p, size0)]
        out = []
        while blocks:
            p, size = blocks.pop(0)
            end = p + size
            while p + 4 <= end:
                mtype = b[p]
                msize = struct.unpack_from("<H", b, p + 1)[0]
                mflags = b[p + 3]
                body = p + 4 + (2 if track else 0)
                if mtype == 0x10:
                    caddr, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((caddr + 4, clen - 8))  # skip OCHK, drop checksum
                if mtype != 0:
                    out.append((mtype, body, msize, mflags))
                p = body + msize
        return out

    # ------------------------------------------------------------------ groups
    def _links(self, addr: int) -> dict[str, int]:
        if addr in self._cache:
            return self._cache[addr]
        b = self.buf
        links: dict[str, int] = {}
        for mtype, body, msize, _ in self._messages(addr):
            if mtype == 0x11:  # symbol table
                btree, heap = struct.unpack_from("<QQ", b, body)
                self._walk_btree(btree, heap, links)
            elif mtype == 0x06:
                name, oaddr = self._parse_link(body)
                if name is not None:
                    links[name] = oaddr
            elif mtype == 0x02:  # link info
                flags = b[body + 1]
                p = body + 2 + (8 if flags & 1 else 0)
                fheap = struct.unpack_from("<Q", b, p)[0]
                if fheap != UNDEF:
                    self._walk_fractal_heap(fheap, links)
        self._cache[addr] = links
        return links

    def _parse_link(self, p: int):
        b = self.buf
        flags = b[p + 1]
        q = p + 2
        ltype = 0
        if flags & 0x08:
            ltype = b[q]
            q += 1
        if flags & 0x04:
            q += 8
        if flags & 0x10:
            q += 1
        lsz = 1 << (flags & 3)
        nlen = int.from_bytes(b[q:q + lsz], "little")
        q += lsz
        name = b[q:q + nlen].decode("utf-8", "replace")
        q += nlen
        if ltype != 0:
            return None, None
        return name, struct.unpack_from("<Q", b, q)[0]

    def _walk_btree(self, addr: int, heap: int, links: dict):
        b = self.buf
        if b[addr:addr + 4] == b"SNOD":
            nsym = struct.unpack_from("<H", b, addr + 6)[0]
            data_addr = struct.unpack_from("<Q", b, heap + 24)[0]
            for i in range(nsym):
                e = addr + 8 + 40 * i
                noff, oaddr = struct.unpack_from("<QQ", b, e)
                s = data_addr + noff
                t = b.index(b"\x00", s)
                links[b[s:t].decode()] = oaddr
            return
        if b[addr:addr + 4] != b"TREE":
            raise ValueError("bad group B-tree node")
        nent = struct.unpack_from("<H", b, addr + 6)[0]
        p = addr + 24
        for i in range(nent):
            child = struct.unpack_from("<Q", b, p + 8)[0]
            self._walk_btree(child, heap, links)
            p += 16

    def _walk_fractal_heap(self, addr: int, links: dict):
        b = self.buf
        if b[addr:addr + 4] != b"FRHP":
            raise ValueError("bad fractal heap")
        hflags = b[addr + 9]
        (table_width,) = struct.unpack_from("<H", b, addr + 110)
        start_bs, max_dbs = struct.unpack_from("<QQ", b, addr + 112)
        (max_heap_bits,) = struct.unpack_from("<H", b, addr + 128)
        root_addr = struct.unpack_from("<Q", b, addr + 132)[0]
        (cur_rows,) = struct.unpack_from("<H", b, addr + 140)
        io_filter_len = struct.unpack_from("<H", b, addr + 7)[0]
        if io_filter_len:
            raise ValueError("filtered fractal heap unsupported")
        off_bytes = (max_heap_bits + 7) // 8
        cks = 4 if (hflags & 0x02) else 0

        def scan_direct(daddr: int, size: int):
            if b[daddr:daddr + 4] != b"FHDB":
                return
            p = daddr + 5 + 8 + off_bytes + cks
            end = daddr + size
            # objects are link messages stored back to back; free space is zero-filled
            while p < end - 10:
                if b[p] == 0:  # free space left by a deleted/relocated object
                    p += 1
                    continue
                if b[p] != 1:  # link message version
                    break
                try:
                    flags = b[p + 1]
                    q = p + 2
                    if flags & 0x08:
                        q += 1
                    if flags & 0x04:
                        q += 8
                    if flags & 0x10:
                        q += 1
                    lsz = 1 << (flags & 3)
                    nlen = int.from_bytes(b[q:q + lsz], "little")
                    name, oaddr = self._parse_link(p)
                    if name is not None:
                        links[name] = oaddr
                    p = q + lsz + nlen + 8
                except Exception:
                    break

        if root_addr == UNDEF:
            return
        if cur_rows == 0:
            scan_direct(root_addr, start_bs)
            return
        # indirect root block: rows of `table_width` direct blocks, sizes double from row 2
        if b[root_addr:root_addr + 4] != b"FHIB":
            raise ValueError("bad fractal heap indirect block")
        p = root_addr + 5 + 8 + off_bytes
        max_direct_rows = (int(max_dbs).bit_length() - int(start_bs).bit_length()) + 2
        for row in range(cur_rows):
            bs = start_bs if row < 2 else start_bs << (row - 1)
            for _ in range(table_width):
                child = struct.unpack_from("<Q", b, p)[0]
                p += 8
                if row >= max_direct_rows:
                    continue  # nested indirect blocks: not needed for the reference files
                if child != UNDEF:
                    scan_direct(child, bs)

    # ------------------------------------------------------------------ datasets
    def _dtype(self, body: int):
        b = self.buf
        cls = b[body] & 0x0F
        bits0 = b[body + 1]
        size = struct.unpack_from("<I", b, body + 4)[0]
        if cls == 0:
            signed = bool(bits0 & 0x08)
            return np.dtype(("<i" if signed else "<u") + str(size))
        if cls == 1:
            return np.dtype("<f" + str(size))
        if cls == 4:          # bitfield: how HDF5.jl stores Julia Bool vectors (measurement layouts)
            return np.dtype("<u" + str(size))
        return None

    def _dims(self, body: int):
        b = self.buf
        ver = b[body]
        rank = b[body + 1]
        if ver == 1:
            p = body + 8
        else:
            if b[body + 3] == 2:  # null dataspace
                return None
            p = body + 4
        return tuple(struct.unpack_from("<Q", b, p + 8 * i)[0] for i in range(rank))

    def read_object(self, addr: int):
        b = self.buf
        dims = dtype = None
        data = None
        for mtype, body, msize, _ in self._messages(addr):
            if mtype == 0x01:
                dims = self._dims(body)
            elif mtype == 0x03:
                dtype = self._dtype(body)
            elif mtype == 0x08:
                ver = b[body]
                if ver != 3:
                    raise ValueError(f"layout version {ver} unsupported")
                lcls = b[body + 1]
                if lcls == 1:
                    daddr, dsize = struct.unpack_from("<QQ", b, body + 2)
                    data = (daddr, dsize)
                elif lcls == 0:
                    dsize = struct.unpack_from("<H", b, body + 2)[0]
                    data = (body + 4, dsize)
                else:
                    raise ValueError("chunked datasets unsupported")
        if dtype is None or data is None or dims is None:
            return None
        n = int(np.prod(dims)) if dims else 1
        daddr, dsize = data
        if daddr == UNDEF:
            return np.zeros(dims, dtype=dtype)
        arr = np.frombuffer(b, dtype=dtype, count=n, offset=daddr).copy()
        # HDF5.jl writes Julia (column-major) arrays with reversed dims
        return arr.reshape(dims) if dims else arr[0]

    def attrs(self, path: str = "/") -> dict:
        addr = self._resolve(path)
        b = self.buf
        out = {}
        for mtype, body, msize, _ in self._messages(addr):
            if mtype != 0x0C:
                continue
            ver = b[body]
            nsz, tsz, ssz = struct.unpack_from("<HHH", b, body + 2)
            p = body + 8 + (1 if ver == 3 else 0)
            pad = (lambda x: (x + 7) & ~7) if ver == 1 else (lambda x: x)
            name = b[p:p + nsz].split(b"\x00")[0].decode()
            p += pad(nsz)
            dtype = self._dtype(p)
            p += pad(tsz)
            dims = self._dims(p)
            p += pad(ssz)
            if dtype is None:
                continue
            n = int(np.prod(dims)) if dims else 1
            val = np.frombuffer(b, dtype=dtype, count=n, offset=p)
            out[name] = val[0] if not dims else val.copy()
        return out

    # ------------------------------------------------------------------ public API
    def _resolve(self, path: str) -> int:
        addr = self.root_addr
        for part in [p for p in path.split("/") if p]:
            links = self._links(addr)
            if part not in links:
                raise KeyError(path)
            addr = links[part]
        return addr

    def keys(self, path: str = "/"):
        return sorted(self._links(self._resolve(path)).keys())

    def is_group(self, path: str) -> bool:
        addr = self._resolve(path)
        return any(m[0] in (0x11, 0x02, 0x06) for m in self._messages(addr))

    def __contains__(self, path: str) -> bool:
        try:
            self._resolve(path)
            return True
        except KeyError:
            return False

    def __getitem__(self, path: str):
        return self.read_object(self._resolve(path))

    def tree(self, path: str = "/", depth: int = 0, maxdepth: int = 6):
        out = []
        for k in self.keys(path):
            full = path.rstrip("/") + "/" + k
            if self.is_group(full) and depth < maxdepth:
                out.append("  " * depth + k + "/")
                out.extend(self.tree(full, depth + 1, maxdepth))
            else:
                v = self[full]
                shp = None if v is None else (v.shape if hasattr(v, "shape") else ())
                out.append("  " * depth + f"{k} {shp}")
        return out
