"""Shared test helpers: SMF fixtures, Dirichlet flows, parity metric."""
import gzip
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "tests", "golden", "ref")

SHAPE_BY_NAME = {"line": 1, "triangle": 2, "quadrilateral": 3, "tetrahedron": 4, "hexahedron": 5}
SHAPE_DIM = {1: 1, 2: 2, 3: 2, 4: 3, 5: 3}


def read_smf(path):
    """Semantics of base/io/smf/Reader.hpp: header keys, then 'nNodes nElems', coordinates (3 per node)
    and connectivities; optional externalNodes / externalElements files."""
    keys = {}
    with open(path) as f:
        lines = [l.strip() for l in f if l.strip() and not l.startswith("#")]
    k = 0
    while lines[k].startswith("!"):
        parts = lines[k][1:].split()
        keys[parts[0]] = parts[1]
        k += 1
    toks = " ".join(lines[k:]).split()
    n_nodes, n_elems = int(toks[0]), int(toks[1])
    npe = int(keys["elementNumPoints"])
    shape = SHAPE_BY_NAME[keys["elementShape"]]
    rest = toks[2:]
    d = os.path.dirname(path)
    if "externalNodes" in keys:
        coords = np.loadtxt(os.path.join(d, keys["externalNodes"]), dtype=np.float64).reshape(n_nodes, 3)
    else:
        coords = np.array(rest[:3 * n_nodes], dtype=np.float64).reshape(n_nodes, 3)
        rest = rest[3 * n_nodes:]
    if "externalElements" in keys:
        conn = np.loadtxt(os.path.join(d, keys["externalElements"]), dtype=np.int64).reshape(n_elems, npe)
    else:
        conn = np.array(rest[:npe * n_elems], dtype=np.int64).reshape(n_elems, npe)
    dim = SHAPE_DIM[shape]
    return shape, np.ascontiguousarray(coords[:, :dim]), conn


def read_pairs(path):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rt") as f:
        rows = [l.split() for l in f if l.strip() and not l.startswith("#")]
    return np.array(rows, dtype=np.int64)


def fund_sol_laplace(x, src):
    """base/auxi/FundamentalSolution.hpp:110-130 (3-D / 2-D)."""
    x = np.atleast_2d(x)
    dim = x.shape[1]
    dist = np.sqrt(((x - src) ** 2).sum(axis=1))
    if dim == 3:
        return (1.0 / (4.0 * np.pi)) * (1.0 / dist)
    return (1.0 / (2.0 * np.pi)) * np.log(1.0 / dist)


def fund_sol_elastostatic(x, y0, direction, lam, mu):
    """base/auxi/FundamentalSolution.hpp:176-251: Kelvin tensor U(y,x) * dir, vectorised over points."""
    x = np.atleast_2d(x)
    dim = x.shape[1]
    G, nu = mu, lam / 2.0 / (lam + mu)
    surf = 2.0 * (dim - 1.0) * np.pi
    fac1 = 1.0 / (4.0 * surf * G * (1.0 - nu))
    fac2 = 3.0 - 4.0 * nu
    R = x - y0
    dist = np.sqrt((R ** 2).sum(axis=1))
    kern = 1.0 / dist if dim == 3 else np.log(1.0 / dist)
    U = np.zeros((x.shape[0], dim, dim))
    for i in range(dim):
        for j in range(dim):
            U[:, i, j] = fac1 * ((fac2 * kern if i == j else 0.0) + (R[:, i] * R[:, j]) / dist ** dim)
    return np.einsum("nij,j->ni", U, direction)


def constrain_boundary(prob, fe_deg, dof_size, elem_dof, n_obj, fun):
    """dof::constrainBoundary flow on flat arrays using the oracle's boundary list / support points.
    fun(x[n,dim]) -> values[n,dof_size].  Returns status[n_obj,ds] (u8) and prescribed[n_obj,ds]."""
    pairs = prob.mesh_boundary()
    elem, loc, x = prob.boundary_dof_points(fe_deg, pairs)
    vals = np.asarray(fun(x), dtype=np.float64).reshape(len(elem), dof_size)
    status = np.zeros((n_obj, dof_size), dtype=np.uint8)
    prescribed = np.zeros((n_obj, dof_size))
    objs = elem_dof[elem, loc]
    for k in range(len(objs)):  # sequential: later visits overwrite (DegreeOfFreedom::constrainValue)
        status[objs[k], :] = 1
        prescribed[objs[k], :] = vals[k]
    return status, prescribed


def csr_rel_diff(rowptr, val_a, val_b):
    """SURVEY 8(d) parity metric: max |a-b| / max(|a|,|b|, max_row |A|)."""
    val_a = np.asarray(val_a); val_b = np.asarray(val_b)
    n = len(rowptr) - 1
    counts = np.diff(rowptr)
    rows = np.repeat(np.arange(n), counts)
    rowmax = np.zeros(n)
    np.maximum.at(rowmax, rows, np.maximum(np.abs(val_a), np.abs(val_b)))
    scale = np.maximum(rowmax[rows], 1e-300)
    return float(np.max(np.abs(val_a - val_b) / scale)) if len(val_a) else 0.0


def vec_rel_diff(a, b):
    a = np.asarray(a); b = np.asarray(b)
    s = max(np.max(np.abs(a)), np.max(np.abs(b)), 1e-300) if len(a) else 1.0
    return float(np.max(np.abs(a - b)) / s) if len(a) else 0.0
