"""Independent Python model of the BAM the host driver writes (test infrastructure): a fabricated batch of compact results
and the uncompressed BAM bytes it must turn into (record fields as AlignRead builds them, src/graph/alignment.go:114-156,
296-315; header as setupBAM, src/pipeline/boss.go:45-105), plus a strict BGZF block checker."""
import struct
import zlib

import numpy as np


def _reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14: return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17: return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20: return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23: return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26: return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


_NT16 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
_COMP = {ord("A"): ord("T"), ord("T"): ord("A"), ord("C"): ord("G"), ord("G"): ord("C"), ord("N"): ord("N")}


class Batch:
    """A fabricated batch + its expected uncompressed BAM record bytes."""

    def __init__(self, rng, n_reads, len_lo, len_hi, max_recs, path_bytes=1, n_nodes=50, n_graphs=7, zero_frac=0.1, quals="random"):
        self.path_bytes = path_bytes
        max_paths = 200 if path_bytes == 1 else 700
        self.graph_paths = rng.integers(1, max_paths, n_graphs)
        self.graph_ref_base = np.concatenate([[0], np.cumsum(self.graph_paths)]).astype(np.uint32)
        self.refs = [("ref|%d|%d" % (g, p), int(rng.integers(100, 5000))) for g in range(n_graphs) for p in range(self.graph_paths[g])]
        self.nodes = []
        for _ in range(n_nodes):
            g = int(rng.integers(0, n_graphs))
            k = int(rng.integers(1, self.graph_paths[g] + 1))
            ids = np.sort(rng.choice(self.graph_paths[g], k, replace=False)).astype(np.uint32)
            self.nodes.append((g, ids, rng.integers(0, 3000, k).astype(np.int32)))
        self.ids, self.seqs, self.quals = [], [], []
        for r in range(n_reads):
            L = int(rng.integers(len_lo, len_hi + 1))
            self.ids.append(b"@" + (b"read_%07d" % r) + bytes(rng.integers(33, 127, int(rng.integers(0, 12))).astype(np.uint8)))
            self.seqs.append(bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), L, p=[0.24, 0.24, 0.24, 0.24, 0.04])))
            if quals == "random":
                self.quals.append(bytes(rng.integers(33, 75, L).astype(np.uint8)))
            else:                                                           # binned qualities with long runs, like recent Illumina output
                self.quals.append(bytes(np.repeat(rng.choice(np.frombuffer(b"F:,#", dtype=np.uint8), L // 7 + 1, p=[0.8, 0.1, 0.07, 0.03]), 7)[:L]))
        self.cpairs, self.rec_path = [], []
        for r in range(n_reads):
            for _ in range(int(rng.integers(0, 3))):
                node = int(rng.integers(0, n_nodes))
                g, ids, _ = self.nodes[node]
                cnt = 0 if rng.random() < zero_frac else int(rng.integers(1, max_recs + 1))
                L = len(self.seqs[r])
                flags = int(rng.integers(0, 1 << 20))
                if rng.random() < 0.5: flags |= 0x10000000
                if L >= 3 and rng.random() < 0.2: flags |= 0x20000000
                if L >= 3 and rng.random() < 0.2: flags |= 0x40000000
                self.cpairs.append((r, node, flags, cnt))
                # mostly paths through the node, a few that are not (position 0 then)
                self.rec_path += [int(ids[rng.integers(0, len(ids))]) if rng.random() < 0.9 else int(rng.integers(0, self.graph_paths[g])) for _ in range(cnt)]

    def write(self, path):
        off = lambda xs: np.concatenate([[0], np.cumsum([len(x) for x in xs])]).astype(np.uint64)
        with open(path, "wb") as f:
            f.write(b"BAMT" + struct.pack("<IIQIIII", len(self.ids), len(self.cpairs), len(self.rec_path), self.path_bytes, len(self.graph_paths), len(self.refs), len(self.nodes)))
            for xs in (self.ids, self.seqs, self.quals): f.write(off(xs).tobytes())
            for xs in (self.ids, self.seqs, self.quals): f.write(b"".join(xs))
            f.write(np.array(self.cpairs, dtype=np.uint32).reshape(-1, 4).tobytes())
            f.write(np.array(self.rec_path, dtype=np.uint8 if self.path_bytes == 1 else np.uint16).tobytes())
            f.write(self.graph_ref_base.tobytes())
            for name, ln in self.refs: f.write(struct.pack("<I", len(name)) + name.encode() + struct.pack("<i", ln))
            for g, ids, pos in self.nodes: f.write(struct.pack("<II", g, len(ids)) + ids.tobytes() + pos.tobytes())

    def expected(self):
        text = b"@HD\tVN:1.5\tSO:unknown\n"
        out = [b"BAM\x01" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(self.refs))]
        for name, ln in self.refs: out.append(struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", ln))
        self.header_bytes = sum(len(x) for x in out)
        at = 0
        for r, node, flags, cnt in self.cpairs:
            g, ids, pos = self.nodes[node]
            seq, qual, name = self.seqs[r], self.quals[r], self.ids[r][1:]
            rev, cs, ce, offset = bool(flags & 0x10000000), 1 if flags & 0x20000000 else 0, 1 if flags & 0x40000000 else 0, flags & 0x0fffffff
            if rev:
                seq = bytes(_COMP.get(c, 0) for c in reversed(seq)); qual = qual[::-1]
            match = len(seq) - cs - ce
            cig = ([(cs << 4) | 5] if cs else []) + [match << 4] + ([(ce << 4) | 5] if ce else [])
            codes = [_NT16.get(chr(c).upper(), 15) for c in seq[:match]] + [0]
            packed = bytes((codes[i] << 4) | codes[i + 1] for i in range(0, match, 2))
            for j in range(cnt):
                path = self.rec_path[at + j]
                k = np.searchsorted(ids, path)
                p = (int(pos[k]) if k < len(ids) and ids[k] == path else 0) + offset
                flag = (0x100 if cnt > 1 and j else 0) | (0x10 if rev else 0)
                body = struct.pack("<iiBBHHHiiii", int(self.graph_ref_base[g]) + path, p, len(name) + 1, 30, _reg2bin(p, p + max(1, match)), len(cig), flag, match, -1, -1, 0)
                body += name + b"\0" + b"".join(struct.pack("<I", c) for c in cig) + packed + qual[:match]
                out.append(struct.pack("<i", len(body)) + body)
            at += cnt
        return b"".join(out)


def check_blocks(raw):
    """Every BGZF block: gzip member with the BC extra field, BSIZE = its size - 1, deflate data that inflates to ISIZE <= 0xff00
    bytes with the right CRC; the last one the 28-byte EOF marker. Returns the concatenated payload."""
    at, out = 0, []
    while at < len(raw):
        assert raw[at:at + 4] == b"\x1f\x8b\x08\x04" and raw[at + 10:at + 16] == b"\x06\x00BC\x02\x00"
        bsize = struct.unpack_from("<H", raw, at + 16)[0] + 1
        crc, isize = struct.unpack_from("<II", raw, at + bsize - 8)
        d = zlib.decompressobj(-15)
        data = d.decompress(raw[at + 18:at + bsize - 8])
        assert d.eof and not d.unused_data and len(data) == isize <= 0xff00 and zlib.crc32(data) == crc
        out.append(data)
        at += bsize
    assert at == len(raw) and out[-1] == b"" and raw[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    return b"".join(out)
