"""TEST INFRASTRUCTURE — a restatement of Go's encoding/gob ENCODER, just large enough to write the reference's index
files: groot.gg (pipeline.Info, src/pipeline/runtime.go:15-27,64-73) and groot.lshe (lshe.ContainmentIndex,
src/lshe/lshe.go:37-49,72-92). There is no Go toolchain in this image, so the library's gob reader
(groot_b200/csrc/host/gob_reader.cpp) is exercised with streams written here, from the published format description:

  message      := uvarint(byte count) body
  body         := varint(-id) wireType-value          (definition of type id; sent before first use)
                | varint(id) [0x00 if not a struct] value
  uvarint      := one byte < 128, or a byte holding the negated byte count followed by big-endian bytes
  varint       := uvarint(i << 1) for i >= 0, uvarint((~i << 1) | 1) otherwise
  float        := uvarint(byte-reversed IEEE-754 bits)
  struct       := (uvarint(field delta) value)* 0x00          zero-valued fields are omitted
  slice / map  := uvarint(count) elements                      elements are always sent
Maps are written in shuffled order (Go's map iteration order is random).
"""
import random
import struct

BOOL, INT, UINT, FLOAT, BYTES, STRING = 1, 2, 3, 4, 5, 6
T_WIRE, T_ARRAY, T_COMMON, T_SLICE, T_STRUCT, T_FIELD, T_FIELDSLICE, T_MAP = 16, 17, 18, 19, 20, 21, 22, 23


def uvarint(u):
    if u < 128:
        return bytes([u])
    b = u.to_bytes((u.bit_length() + 7) // 8, "big")
    return bytes([256 - len(b)]) + b


def varint(i):
    return uvarint((i << 1) if i >= 0 else ((~i) << 1) | 1)


def gfloat(f):
    return uvarint(int.from_bytes(struct.pack("<d", f), "big"))


class Struct:
    def __init__(self, name, fields):
        self.name, self.fields = name, fields          # fields: [(name, type)]


class Slice:
    def __init__(self, elem, name=""):
        self.elem, self.name = elem, name


class Map:
    def __init__(self, key, elem, name=""):
        self.key, self.elem, self.name = key, elem, name


class Encoder:
    def __init__(self, seed=0):
        self.ids, self.out, self.next_id = {}, bytearray(), 65
        self.rng = random.Random(seed) if seed is not None else None      # None: maps in insertion order (what the C++ writer emits)

    # ---- type definitions ----
    def type_id(self, t):
        if isinstance(t, int):
            return t
        if id(t) in self.ids:
            return self.ids[id(t)]
        tid = self.next_id
        self.next_id += 1
        self.ids[id(t)] = tid
        # components first (Go sends a type's definition, then the types it refers to as they are met; any order with
        # definitions before use decodes)
        if isinstance(t, Struct):
            kids = [self.type_id(ft) for _, ft in t.fields]
            fields = b"".join(uvarint(1) + uvarint(len(n)) + n.encode() + uvarint(1) + varint(k) + b"\x00" for (n, _), k in zip(t.fields, kids))
            body = (uvarint(3) +                                                   # wireType.StructT (field 2)
                    uvarint(1) + self._common(t.name, tid) +                       #   CommonType
                    uvarint(1) + uvarint(len(t.fields)) + fields + b"\x00" +       #   Field []*fieldType
                    b"\x00")
        elif isinstance(t, Slice):
            k = self.type_id(t.elem)
            body = uvarint(2) + uvarint(1) + self._common(t.name, tid) + uvarint(1) + varint(k) + b"\x00" + b"\x00"          # SliceT (field 1)
        else:
            kk, ke = self.type_id(t.key), self.type_id(t.elem)
            body = uvarint(4) + uvarint(1) + self._common(t.name, tid) + uvarint(1) + varint(kk) + uvarint(1) + varint(ke) + b"\x00" + b"\x00"   # MapT (field 3)
        self._message(varint(-tid) + body)
        return tid

    @staticmethod
    def _common(name, tid):
        out = b""
        if name:
            out += uvarint(1) + uvarint(len(name)) + name.encode() + uvarint(1) + varint(tid)
        else:
            out += uvarint(2) + varint(tid)
        return out + b"\x00"

    def _message(self, body):
        self.out += uvarint(len(body)) + body

    # ---- values ----
    def _is_zero(self, t, v):
        if isinstance(t, int):
            return v in (0, 0.0, False, b"", "", None)
        if isinstance(t, Slice):
            return v is None or len(v) == 0
        if isinstance(t, Map):
            return v is None
        return False

    def _value(self, t, v):
        if t == BOOL:
            return uvarint(1 if v else 0)
        if t == INT:
            return varint(int(v))
        if t == UINT:
            return uvarint(int(v))
        if t == FLOAT:
            return gfloat(float(v))
        if t in (BYTES, STRING):
            b = v.encode() if isinstance(v, str) else bytes(v)
            return uvarint(len(b)) + b
        if isinstance(t, Struct):
            out, prev = bytearray(), -1
            for i, (name, ft) in enumerate(t.fields):
                fv = v.get(name)
                if fv is None or self._is_zero(ft, fv):
                    continue
                out += uvarint(i - prev) + self._value(ft, fv)
                prev = i
            return bytes(out) + b"\x00"
        if isinstance(t, Slice):
            return uvarint(len(v)) + b"".join(self._value(t.elem, x) for x in v)
        items = list(v.items())
        if self.rng is not None:
            self.rng.shuffle(items)
        return uvarint(len(items)) + b"".join(self._value(t.key, k) + self._value(t.elem, x) for k, x in items)

    def encode(self, t, v):
        tid = self.type_id(t)
        body = varint(tid) + (b"" if isinstance(t, Struct) else b"\x00") + self._value(t, v)
        self._message(body)
        return bytes(self.out)


# ---- the reference's types ---------------------------------------------------------------------------------------------
NODE = Struct("GrootGraphNode", [("SegmentID", UINT), ("SegmentLength", FLOAT), ("Sequence", BYTES), ("OutEdges", Slice(UINT, "Nodes")),
                                 ("PathIDs", Slice(UINT)), ("Position", Map(INT, INT)), ("KmerFreq", FLOAT), ("Marked", BOOL)])
GRAPH = Struct("GrootGraph", [("GrootVersion", STRING), ("GraphID", UINT), ("SortedNodes", Slice(NODE)), ("Paths", Map(UINT, BYTES)),
                              ("Lengths", Map(UINT, INT)), ("NodeLookup", Map(UINT, INT)), ("Masked", BOOL), ("KmerTotal", UINT), ("EMiterations", INT)])
ALIGNCMD = Struct("AlignCmd", [("Fasta", BOOL), ("BloomFilter", BOOL), ("MinKmerCoverage", FLOAT), ("BAMout", STRING), ("NoExactAlign", BOOL)])
HAPLOCMD = Struct("HaploCmd", [("Cutoff", FLOAT), ("MinIterations", INT), ("MaxIterations", INT), ("TotalKmers", INT), ("HaploDir", STRING)])
INFO = Struct("Info", [("Version", STRING), ("NumProc", INT), ("Profiling", BOOL), ("KmerSize", INT), ("SketchSize", INT), ("WindowSize", INT),
                       ("NumPart", INT), ("MaxK", INT), ("MaxSketchSpan", INT), ("ContainmentThreshold", FLOAT), ("IndexDir", STRING),
                       ("Store", Map(UINT, GRAPH, "Store")), ("Sketch", ALIGNCMD), ("Haplotype", HAPLOCMD)])
KEY = Struct("Key", [("GraphID", UINT), ("Node", UINT), ("OffSet", UINT), ("ContainedNodes", Map(UINT, FLOAT)), ("Ref", Slice(UINT)), ("RC", BOOL),
                     ("Sketch", Slice(UINT)), ("Freq", FLOAT), ("MergeSpan", UINT), ("WindowSize", UINT)])
CINDEX = Struct("ContainmentIndex", [("NumPart", INT), ("MaxK", INT), ("NumWindowKmers", INT), ("SketchSize", INT), ("WindowLookup", Map(STRING, KEY))])


def parse_dump(path):
    """The canonical index dump (grootgpu_index_dump_file / the oracle's dump_file) -> (params, graphs, windows)."""
    params, graphs, windows = {}, [], []
    for line in open(path):
        f = line.rstrip("\n").split(" ")
        if f[0] == "I":
            params = {k: int(v) for k, v in (x.split("=") for x in f[1:])}
        elif f[0] == "G":
            kv = {k: int(v) for k, v in (x.split("=") for x in f[2:])}
            graphs.append({"id": int(f[1]), "masked": kv["masked"], "paths": [], "nodes": []})
        elif f[0] == "P":
            graphs[-1]["paths"].append((int(f[1]), int(f[2]), " ".join(f[3:])))
        elif f[0] == "N":
            e, p = f.index("E"), f.index("P")
            graphs[-1]["nodes"].append({"seg": int(f[1]), "seq": f[2], "edges": [int(x) for x in f[e + 1:p]],
                                        "paths": [tuple(int(y) for y in x.split(":")) for x in f[p + 1:]]})
        elif f[0] == "W":
            s, c = f.index("S"), f.index("C")
            windows.append({"graph": int(f[1]), "seg": int(f[2]), "off": int(f[3]), "span": int(f[4].split("=")[1]), "w": int(f[5].split("=")[1]),
                            "sketch": [int(x, 16) for x in f[s + 1:c]], "cn": [tuple(int(y) for y in x.split(":")) for x in f[c + 1:]]})
    return params, graphs, windows


def write_reference_index(dump_path, gg_path, lshe_path, seed=0, kmer_freq=None):
    """Writes groot.gg / groot.lshe as `groot index` would for the index described by a canonical dump."""
    params, graphs, windows = parse_dump(dump_path)
    store, node_i = {}, 0
    for g in graphs:
        nodes = []
        for n in g["nodes"]:
            nodes.append({"SegmentID": n["seg"], "SegmentLength": float(len(n["seq"])), "Sequence": n["seq"].encode(), "OutEdges": n["edges"],
                          "PathIDs": [p for p, _ in n["paths"]], "Position": {p: pos for p, pos in n["paths"]},
                          "KmerFreq": float(kmer_freq[node_i]) if kmer_freq is not None else 0.0})
            node_i += 1
        store[g["id"]] = {"GrootVersion": "1.1.2", "GraphID": g["id"], "SortedNodes": nodes, "Paths": {p: name.encode() for p, _, name in g["paths"]},
                          "Lengths": {p: ln for p, ln, _ in g["paths"]}, "NodeLookup": {n["seg"]: i for i, n in enumerate(g["nodes"])},
                          "Masked": bool(g["masked"])}
    info = {"Version": "1.1.2", "NumProc": 1, "KmerSize": params["k"], "SketchSize": params["S"], "WindowSize": params["w"], "NumPart": params["numPart"],
            "MaxK": params["maxK"], "MaxSketchSpan": 30, "ContainmentThreshold": 0.99, "IndexDir": "index", "Store": store,
            "Sketch": {"MinKmerCoverage": 1.0}, "Haplotype": {}}
    open(gg_path, "wb").write(Encoder(seed).encode(INFO, info))
    lookup, counter = {}, {}
    for w in windows:                       # arrival order within one (graph, node, offset) == dump order
        base = "g%dn%do%d" % (w["graph"], w["seg"], w["off"])
        i = counter.get(base, 0)
        counter[base] = i + 1
        lookup["%s-%d" % (base, i)] = {"GraphID": w["graph"], "Node": w["seg"], "OffSet": w["off"], "ContainedNodes": {s: float(c) for s, c in w["cn"]},
                                       "Ref": [0], "Sketch": w["sketch"], "MergeSpan": w["span"], "WindowSize": w["w"]}
    ci = {"NumPart": params["numPart"], "MaxK": params["maxK"], "NumWindowKmers": params["w"] - params["k"] + 1, "SketchSize": params["S"], "WindowLookup": lookup}
    open(lshe_path, "wb").write(Encoder(seed + 1 if seed is not None else None).encode(CINDEX, ci))
