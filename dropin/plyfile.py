"""Minimal `plyfile` stand-in (the pip package is not installable offline): just what the reference's
scene/gaussian_model.py:264-296,337-389, scene/dataset_readers.py:131-163 and vectree/utils.py:70-103 use --
one-element binary_little_endian (or ascii) vertex tables of scalar properties."""
from __future__ import annotations

import numpy as np

_PLY2NP = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2",
           "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4", "double": "f8", "float64": "f8"}
_NP2PLY = {"i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint", "f4": "float", "f8": "double"}


class PlyProperty:
    def __init__(self, name, dtype):
        self.name, self.dtype = name, np.dtype(dtype)


class PlyElement:
    def __init__(self, name, data):
        self.name, self.data = name, data
        self.properties = tuple(PlyProperty(n, data.dtype[n]) for n in data.dtype.names)

    @staticmethod
    def describe(data, name):
        return PlyElement(name, np.asarray(data))

    @property
    def count(self):
        return len(self.data)

    def __getitem__(self, key):
        return self.data[key]

    def __len__(self):
        return len(self.data)


class PlyData:
    def __init__(self, elements, text=False):
        self.elements = list(elements)
        self.text = text

    def __getitem__(self, name):
        for el in self.elements:
            if el.name == name:
                return el
        raise KeyError(name)

    def write(self, path):
        with open(path, "wb") as f:
            hdr = ["ply", "format ascii 1.0" if self.text else "format binary_little_endian 1.0"]
            for el in self.elements:
                hdr.append(f"element {el.name} {len(el.data)}")
                for p in el.properties:
                    hdr.append(f"property {_NP2PLY[p.dtype.newbyteorder('=').str[1:]]} {p.name}")
            hdr.append("end_header")
            f.write(("\n".join(hdr) + "\n").encode("ascii"))
            for el in self.elements:
                if self.text:
                    np.savetxt(f, np.stack([el.data[n] for n in el.data.dtype.names], axis=1), fmt="%.9g")
                else:
                    f.write(el.data.astype(el.data.dtype.newbyteorder("<")).tobytes())

    @staticmethod
    def read(path):
        with open(path, "rb") as f:
            if f.readline().strip() != b"ply":
                raise ValueError("not a PLY file")
            fmt, elements, cur = None, [], None
            while True:
                line = f.readline().decode("ascii").strip()
                if line == "end_header":
                    break
                tok = line.split()
                if not tok or tok[0] == "comment":
                    continue
                if tok[0] == "format":
                    fmt = tok[1]
                elif tok[0] == "element":
                    cur = (tok[1], int(tok[2]), [])
                    elements.append(cur)
                elif tok[0] == "property":
                    if tok[1] == "list":
                        raise NotImplementedError("list properties are not needed on this path")
                    cur[2].append((tok[2], _PLY2NP[tok[1]]))
            out = []
            for name, count, props in elements:
                if fmt == "ascii":
                    rows = np.loadtxt(f, max_rows=count, ndmin=2)
                    data = np.empty(count, dtype=[(n, t) for n, t in props])
                    for k, (n, _) in enumerate(props):
                        data[n] = rows[:, k]
                else:
                    order = "<" if fmt == "binary_little_endian" else ">"
                    dt = np.dtype([(n, order + t) for n, t in props])
                    data = np.frombuffer(f.read(dt.itemsize * count), dtype=dt, count=count).astype([(n, t) for n, t in props])
                out.append(PlyElement(name, data))
        return PlyData(out, text=(fmt == "ascii"))
