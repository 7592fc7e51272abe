"""Minimal read-only HDF5 reader for the reference's gain-schedule files.

Build-time tool only (never imported by the product path).  It understands just
the subset of HDF5 the reference's `save_lookup_data` produces
(lib/FlightPhysics/src/control.jl:879-893 in the reference): superblock v0,
version-1 object headers with continuation blocks, groups described by Link
messages (0x0006) or old-style symbol tables, and contiguous little-endian
float64 datasets.  No third-party HDF5 library is available in this image.
"""
import struct
import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        with open(path, "rb") as f:
            self.d = f.read()
        d = self.d
        assert d[:8] == b"\x89HDF\r\n\x1a\n", "not an HDF5 file"
        assert d[8] == 0, "only superblock v0 supported"
        assert d[13] == 8 and d[14] == 8, "only 8-byte offsets/lengths supported"
        # root symbol table entry starts at byte 56
        self.root_addr = struct.unpack_from("<Q", d, 56 + 8)[0]

    # -- object headers -------------------------------------------------
    def _messages(self, addr):
        d = self.d
        ver, _, nmsg, _refc, hsize = struct.unpack_from("<BBHII", d, addr)
        assert ver == 1, "only v1 object headers supported"
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", d, pos)
                body = d[pos + 8: pos + 8 + msize]
                pos += 8 + msize
                if mtype == 0x0010:  # continuation
                    off, ln = struct.unpack_from("<QQ", body, 0)
                    blocks.append((off, ln))
                out.append((mtype, body))
        return out

    def _links(self, msgs):
        links = {}
        for mtype, body in msgs:
            if mtype == 0x0006:
                ver, flags = body[0], body[1]
                assert ver == 1
                p = 2
                ltype = 0
                if flags & 0x08:
                    ltype = body[p]; p += 1
                if flags & 0x04:
                    p += 8
                if flags & 0x10:
                    p += 1
                lsz = 1 << (flags & 3)
                nlen = int.from_bytes(body[p:p + lsz], "little"); p += lsz
                name = body[p:p + nlen].decode("utf-8"); p += nlen
                assert ltype == 0, "only hard links supported"
                links[name] = struct.unpack_from("<Q", body, p)[0]
            elif mtype == 0x0011:  # symbol table message
                btree, heap = struct.unpack_from("<QQ", body, 0)
                links.update(self._symtab(btree, heap))
        return links

    def _symtab(self, btree, heap):
        d = self.d
        out = {}
        assert d[heap:heap + 4] == b"HEAP"
        heap_data = struct.unpack_from("<Q", d, heap + 24)[0]

        def name_at(off):
            s = heap_data + off
            e = d.index(b"\0", s)
            return d[s:e].decode("utf-8")

        def walk(node):
            assert d[node:node + 4] == b"TREE"
            ntype, level, nused = struct.unpack_from("<BBH", d, node + 4)
            p = node + 24
            for i in range(nused):
                child = struct.unpack_from("<Q", d, p + 8)[0]
                p += 16
                if level > 0:
                    walk(child)
                else:
                    assert d[child:child + 4] == b"SNOD"
                    nsym = struct.unpack_from("<H", d, child + 6)[0]
                    q = child + 8
                    for _ in range(nsym):
                        noff, oaddr = struct.unpack_from("<QQ", d, q)
                        out[name_at(noff)] = oaddr
                        q += 40
        if btree != UNDEF:
            walk(btree)
        return out

    def _dataset(self, msgs):
        dims = None
        addr = None
        size = None
        dt_ok = False
        for mtype, body in msgs:
            if mtype == 0x0001:
                ver, rank, flags = body[0], body[1], body[2]
                p = 8 if ver == 1 else 4
                dims = struct.unpack_from("<%dQ" % rank, body, p) if rank else ()
            elif mtype == 0x0003:
                cls = body[0] & 0x0F
                sz = struct.unpack_from("<I", body, 4)[0]
                dt_ok = (cls == 1 and sz == 8 and (body[1] & 1) == 0)
            elif mtype == 0x0008:
                ver, lclass = body[0], body[1]
                assert ver == 3 and lclass == 1, "only contiguous v3 layout supported"
                addr, size = struct.unpack_from("<QQ", body, 2)
        if dims is None or addr is None:
            return None
        assert dt_ok, "only little-endian float64 datasets supported"
        n = int(np.prod(dims)) if dims else 1
        assert size == 8 * n
        arr = np.frombuffer(self.d, dtype="<f8", count=n, offset=addr).reshape(dims)
        return arr.copy()

    def read_tree(self, addr=None):
        """Return nested dict {name: ndarray | dict} in on-disk (C) dim order."""
        if addr is None:
            addr = self.root_addr
        msgs = self._messages(addr)
        ds = self._dataset(msgs)
        if ds is not None:
            return ds
        return {k: self.read_tree(a) for k, a in self._links(msgs).items()}


def read(path):
    return H5File(path).read_tree()


if __name__ == "__main__":
    import sys
    t = read(sys.argv[1])

    def show(t, ind=0):
        for k, v in t.items():
            if isinstance(v, dict):
                print(" " * ind + k + "/"); show(v, ind + 2)
            else:
                print(" " * ind + f"{k}: shape={v.shape}")
    show(t)
