"""Synthetic meshes for the benchmark and the parity tests (numpy, vectorised).

Follows the recipe of the reference's mesh tools, restated in memory (no SMF text round trip):
  tools/meshGeneration/unitCube/unitCube.hpp:85-265  nodes x-fastest at (h*i1, h*i2, h*i3), elements x-fastest,
      hex connectivity in hierarchic order; simplices by the 6-tet split of base/cut/DecomposeHyperCube.hpp:83-90
  tools/converter/smfRandom/smfRandom.cpp:96-224     interior nodes moved by d*e, e a random unit direction,
      d uniform in [-1,1]*maxDist*h_min, boundary nodes fixed (time-seeded there, fixed seed here)
"""
import numpy as np

from .engine import HEX, QUAD, TET, TRI

# hierarchic position of the lexicographic corners of a Q1 square / cube (base/mesh/HierarchicOrder.hpp)
_H_QUAD = np.array([0, 1, 3, 2])
_H_HEX = np.array([0, 1, 3, 2, 4, 5, 7, 6])
_TETS = np.array([[0, 1, 3, 4], [1, 3, 4, 5], [3, 4, 5, 7], [1, 3, 5, 2], [3, 5, 2, 7], [5, 2, 7, 6]])
_TRIS = np.array([[0, 1, 3], [2, 3, 1]])


def unit_cube_hex(e1, e2, e3, k0=0, k1=None, global_ids=False):
    """Q1 hexahedra on the unit cube.  With k0/k1 only the element layers k0 <= i3 < k1 (a z-slab) are produced;
    node numbers are then either local to the slab (global_ids=False) or those of the full mesh.
    Returns (coords[n,3] f64, conn[ne,8] int32, node_offset) where node_offset is the global id of local node 0."""
    if k1 is None:
        k1 = e3
    n1, n2 = e1 + 1, e2 + 1
    h1, h2, h3 = 1.0 / e1, 1.0 / e2, 1.0 / e3
    i1 = np.arange(n1, dtype=np.float64) * h1
    i2 = np.arange(n2, dtype=np.float64) * h2
    i3 = np.arange(k0, k1 + 1, dtype=np.float64) * h3
    coords = np.empty((len(i3), n2, n1, 3))
    coords[..., 0] = i1[None, None, :]
    coords[..., 1] = i2[None, :, None]
    coords[..., 2] = i3[:, None, None]
    coords = coords.reshape(-1, 3)
    node_offset = k0 * n1 * n2
    ex = np.arange(e1, dtype=np.int64)[None, None, :]
    ey = np.arange(e2, dtype=np.int64)[None, :, None]
    ez = np.arange(0, k1 - k0, dtype=np.int64)[:, None, None]
    base = (ex + ey * n1 + ez * n1 * n2).reshape(-1)
    if global_ids:
        base = base + node_offset
    lex = np.array([0, 1, n1, n1 + 1, n1 * n2, n1 * n2 + 1, n1 * n2 + n1, n1 * n2 + n1 + 1], dtype=np.int64)
    conn = np.empty((base.size, 8), dtype=np.int32)
    conn[:, _H_HEX] = (base[:, None] + lex[None, :]).astype(np.int32)
    return coords, conn, node_offset


def unit_square_quad(e1, e2):
    n1 = e1 + 1
    x = np.arange(n1, dtype=np.float64) / e1
    y = np.arange(e2 + 1, dtype=np.float64) / e2
    coords = np.empty((e2 + 1, n1, 2))
    coords[..., 0] = x[None, :]
    coords[..., 1] = y[:, None]
    base = (np.arange(e1)[None, :] + np.arange(e2)[:, None] * n1).reshape(-1)
    lex = np.array([0, 1, n1, n1 + 1])
    conn = np.empty((base.size, 4), dtype=np.int32)
    conn[:, _H_QUAD] = base[:, None] + lex[None, :]
    return coords.reshape(-1, 2), conn


def unit_cube_tet(e1, e2, e3):
    """P1 tetrahedra: every cube split into six (unitCube.hpp:107-134)."""
    coords, hexc, _ = unit_cube_hex(e1, e2, e3)
    # hexc is in hierarchic order already; the split table is written in hierarchic vertex numbers and is
    # applied through HierarchicOrder to the lexicographic corner list, which is an involution for Q1.
    lexi = hexc[:, _H_HEX]                       # lexicographic corner list
    conn = lexi[:, _H_HEX[_TETS]].reshape(-1, 4)  # cube[HO(v)]
    return coords, np.ascontiguousarray(conn, dtype=np.int32)


def unit_square_tri(e1, e2):
    coords, quad = unit_square_quad(e1, e2)
    lexi = quad[:, _H_QUAD]
    conn = lexi[:, _H_QUAD[_TRIS]].reshape(-1, 3)
    return coords, np.ascontiguousarray(conn, dtype=np.int32)


def boundary_node_mask(coords, tol=1e-12):
    """Nodes on the surface of the unit square / cube."""
    return np.any((coords < tol) | (coords > 1.0 - tol), axis=1)


def perturb_interior(coords, h_min, max_dist=0.1, seed=12345):
    """smfRandom recipe with a fixed-seed generator: x -> x + d*e for interior nodes."""
    rng = np.random.default_rng(seed)
    n, dim = coords.shape
    direction = rng.uniform(-1.0, 1.0, size=(n, dim))
    norm = np.linalg.norm(direction, axis=1, keepdims=True)
    direction = np.where(norm > 1e-10, direction / np.maximum(norm, 1e-300), direction)
    dist = rng.uniform(-1.0, 1.0, size=(n, 1)) * (h_min * max_dist)
    out = coords.copy()
    interior = ~boundary_node_mask(coords)
    out[interior] += (dist * direction)[interior]
    return out


def permute_elements(conn, seed=54321):
    """Random element order to defeat locality for the 'unstructured' runs."""
    rng = np.random.default_rng(seed)
    return np.ascontiguousarray(conn[rng.permutation(conn.shape[0])])
