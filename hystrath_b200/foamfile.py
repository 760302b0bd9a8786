"""Reader / writer for the ASCII and binary OpenFOAM file formats an unchanged dsmcFoam+ case directory uses:
dictionaries (`FoamFile` header, `{}` sub-dictionaries, `( )` lists, `//` and `/* */` comments),
polyMesh files, lagrangian cloud files (BASIC/IOPosition/IOPosition.C:65-150,
DSMC/parcels/dsmcParcelIO.C:133-450) and volScalarField internalField blocks.

Used by hystrath_b200.case (Python face of the case reader), by tests, and by
tests/golden/make_golden.py.  The C++ driver has its own parser (csrc/foam_dict.cpp).
"""
from __future__ import annotations

import re

import numpy as np

_COMMENT = re.compile(r"//[^\n]*|/\*.*?\*/", re.S)


def strip_comments(text: str) -> str:
    return _COMMENT.sub(" ", text)


def _tokenise(text: str):
    # punctuation tokens: { } ( ) ; ; strings in quotes kept whole
    return re.findall(r'"[^"]*"|[{}();]|[^\s{}();]+', text)


class FoamDict(dict):
    """Ordered dictionary with OpenFOAM-style lookups."""

    def lookup(self, key):
        if key not in self:
            raise KeyError(f"keyword {key} is undefined in dictionary")
        return self[key]

    def lookup_or_default(self, key, default):
        return self.get(key, default)

    def sub_dict(self, key):
        v = self.lookup(key)
        if not isinstance(v, FoamDict):
            raise KeyError(f"{key} is not a dictionary")
        return v


def _convert(tok: str):
    if tok.startswith('"'):
        return tok[1:-1]
    try:
        return int(tok)
    except ValueError:
        pass
    try:
        return float(tok)
    except ValueError:
        return tok


def _parse_list(toks, i):
    """toks[i] == '(' ; returns (list, next index).  Entries may be words, numbers, lists or dicts."""
    assert toks[i] == "("
    i += 1
    out = []
    while toks[i] != ")":
        if toks[i] == "(":
            v, i = _parse_list(toks, i)
            out.append(v)
        elif toks[i] == "{":
            v, i = _parse_dict_body(toks, i + 1)
            out.append(v)
        elif i + 1 < len(toks) and toks[i + 1] == "{" and not _is_number(toks[i]):
            v, j = _parse_dict_body(toks, i + 2)
            out.append((toks[i], v))
            i = j
        elif _is_number(toks[i]) and i + 1 < len(toks) and toks[i + 1] == "(":
            v, i = _parse_list(toks, i + 1)  # sized list  N ( ... )
            out.append(v)
        else:
            out.append(_convert(toks[i]))
            i += 1
    return out, i + 1


def _is_number(tok):
    try:
        float(tok)
        return True
    except ValueError:
        return False


def _parse_dict_body(toks, i):
    """Parse entries until the matching '}' (or end of tokens); returns (FoamDict, next index)."""
    d = FoamDict()
    n = len(toks)
    while i < n and toks[i] != "}":
        key = toks[i]
        i += 1
        if i >= n:
            break
        if toks[i] == "{":
            v, i = _parse_dict_body(toks, i + 1)
            d[_convert(key) if key.startswith('"') else key] = v
            continue
        vals = []
        while i < n and toks[i] != ";":
            if toks[i] == "(":
                v, i = _parse_list(toks, i)
                vals.append(v)
            elif toks[i] == "{":
                v, i = _parse_dict_body(toks, i + 1)
                vals.append(v)
            else:
                vals.append(_convert(toks[i]))
                i += 1
        i += 1  # ';'
        # "N ( ... )" sized list -> the list
        if len(vals) == 2 and isinstance(vals[0], int) and isinstance(vals[1], list):
            vals = [vals[1]]
        d[key] = vals[0] if len(vals) == 1 else (vals if vals else None)
    return d, i + 1


def parse_dict(text: str) -> FoamDict:
    """Parse a whole dictionary file.  Top-level bare lists `name ( ... );` are supported (boundariesDict)."""
    toks = _tokenise(strip_comments(text))
    d, _ = _parse_dict_body(toks, 0)
    return d


def read_dict(path) -> FoamDict:
    with open(path) as f:
        return parse_dict(f.read())


# ---------------------------------------------------------------------------------------------
# bulk data files
# ---------------------------------------------------------------------------------------------
def _body(text: str) -> str:
    """File content after the FoamFile header dictionary, comments stripped."""
    text = strip_comments(text)
    m = re.search(r"FoamFile\s*\{[^}]*\}", text)
    return text[m.end():] if m else text


class _Binary:
    """Cursor over the bytes of a `format binary;` file behind its header (IOstream::BINARY: a contiguous list is its size followed by
    the raw bytes in round brackets, OSstream::write(const char*, streamsize); BASIC/particle/particleIO.C:121-143)."""

    def __init__(self, raw: bytes, label):
        self.b, self.i, self.label = raw, 0, label

    def skip(self):
        b = self.b
        while self.i < len(b):
            if b[self.i:self.i + 1].isspace():
                self.i += 1
            elif b[self.i:self.i + 2] == b"//":
                e = b.find(b"\n", self.i)
                self.i = len(b) if e < 0 else e
            elif b[self.i:self.i + 2] == b"/*":
                self.i = b.index(b"*/", self.i) + 2
            else:
                break

    def seek(self, pattern: bytes):
        m = re.compile(pattern).search(self.b, self.i)
        if not m:
            raise ValueError(f"{pattern!r} not found")
        self.i = m.end()
        return m

    def integer(self) -> int:
        self.skip()
        m = re.compile(rb"\d+").match(self.b, self.i)
        if not m:
            raise ValueError("list size expected")
        self.i = m.end()
        return int(m.group(0))

    def expect(self, c: bytes):
        self.skip()
        assert self.b[self.i:self.i + 1] == c, (c, self.b[self.i:self.i + 8])
        self.i += 1

    def block(self, dtype, count):
        """`(` count items of dtype `)`; an empty list has no block at all."""
        if count == 0:
            return np.zeros(0, dtype)
        self.expect(b"(")
        n = np.dtype(dtype).itemsize * count
        a = np.frombuffer(self.b, dtype=dtype, count=count, offset=self.i).copy()
        self.i += n
        assert self.b[self.i:self.i + 1] == b")"
        self.i += 1
        return a

    def labels(self, count):
        return self.block(self.label, count).astype(np.int32)


def _binary(path):
    """_Binary over the file's payload when its header says `format binary;`, else None."""
    raw = open(path, "rb").read()
    m = re.search(rb"FoamFile\s*\{[^}]*\}", raw)
    if not m or not re.search(rb"format\s+binary\s*;", m.group(0)):
        return None
    return _Binary(raw[m.end():], np.int64 if b"label=64" in m.group(0) else np.int32)


def _sized_block(body: str, start=0):
    """Locate `N (` or `N {v}`; return (N, inner text or None, uniform value text or None, end index)."""
    m = re.compile(r"(\d+)\s*([({])").search(body, start)
    if not m:
        raise ValueError("no sized list found")
    n = int(m.group(1))
    if m.group(2) == "{":
        e = body.index("}", m.end())
        return n, None, body[m.end():e].strip(), e + 1
    depth, i = 1, m.end()
    while depth:
        c = body[i]
        depth += c == "("
        depth -= c == ")"
        i += 1
    return n, body[m.end():i - 1], None, i


def read_scalar_list(path, dtype=np.float64):
    """`N ( a b c ... )` or `N{v}` (labelList / scalarList / IOField<scalar>)."""
    b = _binary(path)
    if b is not None:
        n = b.integer()
        return b.block(np.float64, n).astype(dtype) if np.issubdtype(dtype, np.floating) else b.labels(n).astype(dtype)
    n, inner, uni, _ = _sized_block(_body(open(path).read()))
    if inner is None:
        return np.full(n, dtype(float(uni)))
    a = np.array(inner.split(), dtype=np.float64).astype(dtype)
    assert len(a) == n, (path, len(a), n)
    return a


def read_vector_list(path):
    """`N ( (x y z) ... )` (pointField / vectorField)."""
    b = _binary(path)
    if b is not None:
        n = b.integer()
        return b.block(np.float64, 3 * n).reshape(n, 3)
    n, inner, uni, _ = _sized_block(_body(open(path).read()))
    if inner is None:
        v = np.array(uni.replace("(", " ").replace(")", " ").split(), dtype=np.float64)
        return np.tile(v, (n, 1))
    a = np.array(inner.replace("(", " ").replace(")", " ").split(), dtype=np.float64)
    return a.reshape(n, 3)


def read_faces(path):
    """faceList `N ( 4(a b c d) 3(a b c) ... )` -> (offsets[N+1], labels); in binary a faceCompactList: offsets, then all labels."""
    b = _binary(path)
    if b is not None:
        offs = b.labels(b.integer())
        return offs, b.labels(b.integer())
    n, inner, _, _ = _sized_block(_body(open(path).read()))
    offs = np.zeros(n + 1, np.int32)
    labels = []
    k = 0
    for m in re.finditer(r"(\d+)\s*\(([^)]*)\)", inner):
        pts = m.group(2).split()
        assert int(m.group(1)) == len(pts)
        labels.extend(pts)
        k += 1
        offs[k] = len(labels)
    assert k == n
    return offs, np.array(labels, dtype=np.int32)


def read_positions(path):
    """Cloud `positions`: `N ( (x y z) cell ... )` -> (xyz[N,3], cell[N]); in binary one block of position, cellI, faceI,
    stepFraction per particle."""
    b = _binary(path)
    if b is not None:
        n = b.integer()
        b.expect(b"(")
        rec = np.dtype([("x", np.float64, 3), ("cell", b.label), ("face", b.label), ("stepFraction", np.float64)])
        xyz, cell = np.zeros((n, 3)), np.zeros(n, np.int32)
        for i in range(n):
            r = b.block(rec, 1)[0]
            xyz[i], cell[i] = r["x"], r["cell"]
        return xyz, cell
    n, inner, _, _ = _sized_block(_body(open(path).read()))
    a = np.array(inner.replace("(", " ").replace(")", " ").split(), dtype=np.float64).reshape(n, 4)
    return a[:, :3].copy(), a[:, 3].astype(np.int32)


def read_label_list_list(path):
    """List<labelList> such as vibLevel: `N ( 1(i) 1(j) ... )` -> int array [N, width] (ragged rows padded with 0)."""
    b = _binary(path)
    if b is not None:
        n = b.integer()
        b.expect(b"(")
        rows = [b.labels(b.integer()) for _ in range(n)]
        w = max((len(r) for r in rows), default=0)
        out = np.zeros((n, w), np.int32)
        for i, r in enumerate(rows):
            out[i, :len(r)] = r
        return out
    n, inner, uni, _ = _sized_block(_body(open(path).read()))
    if inner is None:
        m = re.match(r"(\d+)\s*\(([^)]*)\)", uni)
        row = [int(x) for x in m.group(2).split()] if m else []
        return np.tile(np.array(row, np.int32), (n, 1)) if row else np.zeros((n, 0), np.int32)
    rows = [[int(x) for x in m.group(2).split()] for m in re.finditer(r"(\d+)\s*\(([^)]*)\)", inner)]
    assert len(rows) == n
    w = max((len(r) for r in rows), default=0)
    out = np.zeros((n, w), np.int32)
    for i, r in enumerate(rows):
        out[i, :len(r)] = r
    return out


def read_boundary(path):
    """constant/polyMesh/boundary -> list of dicts (name, type, nFaces, startFace, ...)."""
    body = _body(open(path).read())
    n, inner, _, _ = _sized_block(body)
    toks = _tokenise(inner)
    out, i = [], 0
    while i < len(toks):
        name = toks[i]
        assert toks[i + 1] == "{", toks[i:i + 3]
        d, i = _parse_dict_body(toks, i + 2)
        d["name"] = name
        out.append(d)
    assert len(out) == n
    return out


def read_internal_field(path):
    """volScalarField / volVectorField internalField -> numpy array (uniform fields return a 0-d / (3,) array)."""
    b = _binary(path)
    if b is not None:
        m = b.seek(rb"internalField\s+(nonuniform\s+List<(\w+)>|uniform)\s*")
        if m.group(1) == b"uniform":
            e = b.b.index(b";", b.i)
            return np.array(b.b[b.i:e].decode().replace("(", " ").replace(")", " ").split(), dtype=np.float64).squeeze()
        w = {b"scalar": 1, b"vector": 3, b"tensor": 9}[m.group(2)]
        n = b.integer()
        a = b.block(np.float64, n * w)
        return a if w == 1 else a.reshape(n, w)
    body = _body(open(path).read())
    m = re.search(r"internalField\s+(nonuniform\s+List<(\w+)>|uniform)\s*", body)
    if not m:
        raise ValueError(f"{path}: no internalField")
    if m.group(1) == "uniform":
        e = body.index(";", m.end())
        vals = body[m.end():e].replace("(", " ").replace(")", " ").split()
        return np.array(vals, dtype=np.float64).squeeze()
    n, inner, _, _ = _sized_block(body, m.end())
    a = np.array(inner.replace("(", " ").replace(")", " ").split(), dtype=np.float64)
    return a if m.group(2) == "scalar" else a.reshape(n, -1)


def read_patch_field(path, patch):
    """boundaryField value of one patch of a volScalarField / volVectorField -> numpy array ([n] or [n, 3]); a `uniform` value
    comes back 0-d / (3,)."""
    bb = _binary(path)
    if bb is not None:
        bb.seek(rb"boundaryField")
        bb.seek(rb"\b" + re.escape(patch).encode() + rb"\s*\{")
        v = bb.seek(rb"value\s+(nonuniform\s+List<(\w+)>|uniform)\s*")   # a patch without a value would take the next patch's: callers ask for calculated patches
        if v.group(1) == b"uniform":
            e = bb.b.index(b";", bb.i)
            return np.array(bb.b[bb.i:e].decode().replace("(", " ").replace(")", " ").split(), dtype=np.float64).squeeze()
        w = {b"scalar": 1, b"vector": 3, b"tensor": 9}[v.group(2)]
        n = bb.integer()
        a = bb.block(np.float64, n * w)
        return a if w == 1 else a.reshape(n, w)
    body = _body(open(path).read())
    b = body.index("boundaryField")
    m = re.search(r"\b" + re.escape(patch) + r"\s*\{", body[b:])
    if not m:
        raise KeyError(patch)
    start = b + m.end()
    end = body.index("}", start)
    block = body[start:end]
    v = re.search(r"value\s+(nonuniform\s+List<(\w+)>|uniform)\s*", block)
    if not v:
        raise ValueError(f"{path}: patch {patch} has no value")
    if v.group(1) == "uniform":
        e = block.index(";", v.end())
        return np.array(block[v.end():e].replace("(", " ").replace(")", " ").split(), dtype=np.float64).squeeze()
    n, inner, _, _ = _sized_block(block, v.end())
    a = np.array(inner.replace("(", " ").replace(")", " ").split(), dtype=np.float64)
    return a if v.group(2) == "scalar" else a.reshape(n, -1)


# ---------------------------------------------------------------------------------------------
# writers (cloud + field files in the layout of DSMC/parcels/dsmcParcelIO.C:338-450)
# ---------------------------------------------------------------------------------------------
_BANNER = """/*--------------------------------*- C++ -*----------------------------------*\\
| =========                 |                                                 |
| \\\\      /  F ield         | OpenFOAM: The Open Source CFD Toolbox           |
|  \\\\    /   O peration     | Version:  v1706                                 |
|   \\\\  /    A nd           | Web:      www.OpenFOAM.com                      |
|    \\\\/     M anipulation  |                                                 |
\\*---------------------------------------------------------------------------*/
"""


def header(cls, location, obj, binary=False, label64=False):
    return (_BANNER + "FoamFile\n{\n    version     2.0;\n    format      " +
            (f"binary;\n    arch        \"LSB;label={64 if label64 else 32};scalar=64\";\n" if binary else "ascii;\n") +
            f"    class       {cls};\n    location    \"{location}\";\n    object      {obj};\n}}\n"
            "// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n\n")


def _binary_list(a, dtype) -> bytes:
    a = np.ascontiguousarray(a, dtype=dtype)
    n = len(a)
    return f"\n{n}\n".encode() + (b"(" + a.tobytes() + b")" if n else b"")


def write_scalar_list(path, cls, location, obj, a, fmt="%.10g", binary=False, label64=False):
    a = np.asarray(a)
    if binary:
        lab = np.int64 if label64 else np.int32
        with open(path, "wb") as f:
            f.write(header(cls, location, obj, True, label64).encode() + _binary_list(a, lab if np.issubdtype(a.dtype, np.integer) else np.float64) + b"\n")
        return
    with open(path, "w") as f:
        f.write(header(cls, location, obj))
        if len(a) and np.all(a == a[0]):
            f.write(f"{len(a)}{{{fmt % a[0]}}}\n")
        else:
            f.write(f"{len(a)}\n(\n" + "\n".join(fmt % v for v in a) + "\n)\n")


def write_vector_list(path, cls, location, obj, a, fmt="%.10g", binary=False, label64=False):
    if binary:
        a = np.ascontiguousarray(a, dtype=np.float64)
        with open(path, "wb") as f:
            f.write(header(cls, location, obj, True, label64).encode() + f"\n{len(a)}\n".encode() + (b"(" + a.tobytes() + b")" if len(a) else b"") + b"\n")
        return
    with open(path, "w") as f:
        f.write(header(cls, location, obj))
        f.write(f"{len(a)}\n(\n" + "\n".join("(" + " ".join(fmt % c for c in v) + ")" for v in a) + "\n)\n")


def write_positions(path, location, xyz, cell, fmt="%.10g", binary=False, label64=False):
    if binary:
        lab = np.int64 if label64 else np.int32
        rec = np.zeros(len(cell), np.dtype([("x", np.float64, 3), ("cell", lab), ("face", lab), ("stepFraction", np.float64)]))
        rec["x"], rec["cell"], rec["face"] = xyz, cell, -1
        with open(path, "wb") as f:
            f.write(header("Cloud<dsmcParcel>", location, "positions", True, label64).encode() + f"{len(cell)}\n(\n".encode())
            f.write(b"".join(b"(" + r.tobytes() + b")\n" for r in rec) + b")\n")
        return
    with open(path, "w") as f:
        f.write(header("Cloud<dsmcParcel>", location, "positions"))
        f.write(f"{len(cell)}\n(\n" + "\n".join("(" + " ".join(fmt % c for c in p) + f") {c}" for p, c in zip(xyz, cell)) + "\n)\n")


def write_label_list_list(path, cls, location, obj, a, binary=False, label64=False):
    a = np.asarray(a)
    if binary:
        lab = np.int64 if label64 else np.int32
        with open(path, "wb") as f:
            f.write(header(cls, location, obj, True, label64).encode() + f"{len(a)}\n(".encode() + b"".join(_binary_list(r, lab) for r in a) + b"\n)\n")
        return
    with open(path, "w") as f:
        f.write(header(cls, location, obj))
        f.write(f"{len(a)}\n(\n" + "\n".join(f"{len(r)}(" + " ".join(str(int(x)) for x in r) + ")" for r in a) + "\n)\n")


def write_faces(path, location, offsets, labels, binary=False, label64=False):
    if binary:   # faceCompactIOList
        lab = np.int64 if label64 else np.int32
        with open(path, "wb") as f:
            f.write(header("faceCompactList", location, "faces", True, label64).encode() + _binary_list(offsets, lab) + _binary_list(labels, lab) + b"\n")
        return
    with open(path, "w") as f:
        f.write(header("faceList", location, "faces"))
        n = len(offsets) - 1
        f.write(f"{n}\n(\n" + "\n".join(f"{offsets[i + 1] - offsets[i]}(" + " ".join(str(int(x)) for x in labels[offsets[i]:offsets[i + 1]]) + ")"
                                         for i in range(n)) + "\n)\n")


def convert_case_to_binary(case_dir, time_name, label64=False):
    """Rewrite the polyMesh and the cloud of <time_name> of a case in `format binary;` (what `foamFormatConvert` does after
    `writeFormat binary;`), with 32- or 64-bit labels: the same numbers, so a reader must find the same mesh and cloud."""
    import os
    pm = os.path.join(case_dir, "constant", "polyMesh")
    write_vector_list(os.path.join(pm, "points"), "vectorField", "constant/polyMesh", "points", read_vector_list(os.path.join(pm, "points")), binary=True, label64=label64)
    offs, labels = read_faces(os.path.join(pm, "faces"))
    write_faces(os.path.join(pm, "faces"), "constant/polyMesh", offs, labels, binary=True, label64=label64)
    for name in ("owner", "neighbour"):
        write_scalar_list(os.path.join(pm, name), "labelList", "constant/polyMesh", name, read_scalar_list(os.path.join(pm, name), np.int32), binary=True, label64=label64)
    loc = f"{time_name}/lagrangian/dsmc"
    d = os.path.join(case_dir, time_name, "lagrangian", "dsmc")
    xyz, cell = read_positions(os.path.join(d, "positions"))
    write_positions(os.path.join(d, "positions"), loc, xyz, cell, binary=True, label64=label64)
    for name in sorted(os.listdir(d)):
        path = os.path.join(d, name)
        if name == "positions" or not os.path.isfile(path):
            continue
        cls = re.search(rb"class\s+([^;]+);", open(path, "rb").read(2000)).group(1).strip().decode()
        if cls == "vectorField":
            write_vector_list(path, cls, loc, name, read_vector_list(path), binary=True, label64=label64)
        elif cls == "scalarField":
            write_scalar_list(path, cls, loc, name, read_scalar_list(path), binary=True, label64=label64)
        elif cls == "labelField":
            write_scalar_list(path, cls, loc, name, read_scalar_list(path, np.int32), binary=True, label64=label64)
        elif cls in ("labelFieldField", "labelListList"):
            write_label_list_list(path, cls, loc, name, read_label_list_list(path), binary=True, label64=label64)
