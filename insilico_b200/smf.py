"""SMF mesh files, the reference's mesh format (base/io/smf/Reader.hpp:81-313, Writer.hpp:47-117):

    # comment lines                      (before and after the header)
    ! elementShape hexahedron            (point | line | triangle | quadrilateral | tetrahedron | hexahedron)
    ! elementNumPoints 8
    ! externalNodes <file>               (optional: coordinates / connectivity in separate files)
    ! externalElements <file>
    <numNodes>  <numElements>
    x y z                                (one node per line; the reader takes the first `dim` numbers of the line)
    n0 n1 ... n7                         (one element per line, node indices in the reference's hierarchic order)

read() gives the flat arrays isl_mesh_set takes; write() produces what the reference's Reader accepts (coordinates with 17
significant digits so that doubles survive the round trip; the reference's own Writer prints 6).
"""
import io
import os

import numpy as np

from . import engine as E

SHAPE_NAME = {E.TRI: "triangle", E.QUAD: "quadrilateral", E.TET: "tetrahedron", E.HEX: "hexahedron"}
_NAME_SHAPE = {v: k for k, v in SHAPE_NAME.items()}
_NAME_SHAPE["line"] = 1
SHAPE_DIM = {1: 1, E.TRI: 2, E.QUAD: 2, E.TET: 3, E.HEX: 3}


class SmfError(ValueError):
    pass


def _numbers(stream, count, per_line, dtype, take):
    """`count` lines, the first `take` numbers of each (the rest of the line is ignored, Reader.hpp:271-277)"""
    out = np.empty((count, take), dtype=dtype)
    for k in range(count):
        line = stream.readline()
        while line and not line.strip():
            line = stream.readline()
        if not line:
            raise SmfError("SMF file ends after %d of %d %s" % (k, count, per_line))
        tok = line.split()
        if len(tok) < take:
            raise SmfError("SMF %s line %d has %d entries, expected %d" % (per_line, k, len(tok), take))
        out[k] = [dtype(t) for t in tok[:take]]
    return out


def read(source, shape=None, dim=None):
    """source: path or text stream.  shape: expected element shape (engine constant) -- a mismatch is an error like the
    reference's "Smf file has unexpected shape value"; dim: coordinates per node to keep (default: dimension of the shape).
    Returns (shape, geom_deg, coords [n, dim] f64, conn [ne, npe] int32)."""
    own = isinstance(source, (str, os.PathLike))
    stream = open(source) if own else source
    base_dir = os.path.dirname(os.fspath(source)) if own else "."
    try:
        found_shape, npe, ext_nodes, ext_elems = None, None, None, None
        pos = stream.tell()
        line = stream.readline()
        while line.startswith("#"):
            pos = stream.tell(); line = stream.readline()
        while line.startswith("!"):
            tok = line[1:].split()
            if tok and "elementShape" in tok[0]:
                name = tok[1].lower()
                match = [s for n, s in _NAME_SHAPE.items() if n in name]
                if not match:
                    raise SmfError("Smf file has unknown shape value: " + tok[1])
                found_shape = match[0]
            elif tok and "elementNumPoints" in tok[0]:
                npe = int(tok[1])
            elif tok and "externalNodes" in tok[0]:
                ext_nodes = tok[1]
            elif tok and "externalElements" in tok[0]:
                ext_elems = tok[1]
            pos = stream.tell(); line = stream.readline()
        while line.startswith("#"):
            pos = stream.tell(); line = stream.readline()
        if found_shape is None or npe is None:
            raise SmfError("Smf header is invalid")
        if shape is not None and shape != found_shape:
            raise SmfError("Smf file has unexpected shape value: %s != %s" % (SHAPE_NAME.get(found_shape, found_shape), SHAPE_NAME.get(shape, shape)))
        tok = line.split()
        if len(tok) < 2:
            raise SmfError("SMF file: expected '<numNodes> <numElements>'")
        n_nodes, n_elems = int(tok[0]), int(tok[1])
        dim = SHAPE_DIM[found_shape] if dim is None else dim

        def section(ext):
            return open(ext if os.path.isabs(ext) else os.path.join(base_dir, ext)) if ext else stream
        s = section(ext_nodes)
        coords = _numbers(s, n_nodes, "node", float, dim)
        if ext_nodes:
            s.close()
        s = section(ext_elems)
        conn = _numbers(s, n_elems, "element", int, npe)
        if ext_elems:
            s.close()
    finally:
        if own:
            stream.close()
    if conn.size and (conn.min() < 0 or conn.max() >= n_nodes):
        raise SmfError("SMF connectivity refers to node %d of %d" % (int(conn.max()), n_nodes))
    # geometry degree from the number of points per element (Lagrange elements)
    geom_deg = None
    for deg in ((1, 2, 3) if found_shape in (E.QUAD, E.HEX, 1) else (1, 2)):
        if E.shape_nfun(found_shape, deg) == npe:
            geom_deg = deg
    if geom_deg is None:
        raise SmfError("Smf file has unexpected number of element nodes: %d" % npe)
    return found_shape, geom_deg, np.ascontiguousarray(coords, dtype=np.float64), np.ascontiguousarray(conn, dtype=np.int32)


def write(target, shape, coords, conn, comment=None):
    """target: path or text stream"""
    own = isinstance(target, (str, os.PathLike))
    stream = open(target, "w") if own else target
    try:
        if comment:
            stream.write("# %s\n" % comment)
        stream.write("! elementShape %s\n! elementNumPoints %d\n%d  %d\n" % (SHAPE_NAME[shape], conn.shape[1], len(coords), len(conn)))
        c3 = np.zeros((len(coords), 3))
        c3[:, :coords.shape[1]] = coords
        for x in c3:
            stream.write("%.17g %.17g %.17g \n" % tuple(x))
        for e in conn:
            stream.write(" ".join(str(int(v)) for v in e) + " \n")
    finally:
        if own:
            stream.close()


def loads(text, **kw):
    return read(io.StringIO(text), **kw)
