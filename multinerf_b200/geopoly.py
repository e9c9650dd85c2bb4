"""Host-side constant generator for the IPE projection basis.

The hot path projects every Gaussian onto K unit vectors taken from a tessellated
icosahedron / octahedron (reference: internal/geopoly.py:78-124 `generate_basis`,
used at internal/models.py:388-389).  The ORDER and SIGN of the K vectors decide
which MLP input column a feature lands in, so this generator reproduces the
reference enumeration exactly (pinned by tests/golden/geopoly_*.npz, which were
produced by running the reference file itself, and by the golden bases of the
reference's tests/geopoly_test.py:78-136).

Enumeration (same as the reference, written as explicit first-occurrence scans):
  1. for each base face, in face order, emit the barycentric lattice points
     (i, j, v-i-j)/v, i outer / j inner, pushed onto the unit sphere;
  2. keep a vertex only if no earlier-emitted vertex lies within sqrt(eps);
  3. keep a vertex only if its antipode appears at the same or a later position
     (drops the second member of every +/- pair);
  4. reverse the xyz column order.
"""
import itertools

import numpy as np

_PHI = (np.sqrt(5.0) + 1.0) / 2.0

_ICOSA_VERTS = np.array(
    [(-1, 0, _PHI), (1, 0, _PHI), (-1, 0, -_PHI), (1, 0, -_PHI), (0, _PHI, 1), (0, _PHI, -1),
     (0, -_PHI, 1), (0, -_PHI, -1), (_PHI, 1, 0), (-_PHI, 1, 0), (_PHI, -1, 0),
     (-_PHI, -1, 0)]) / np.sqrt(_PHI + 2.0)
_ICOSA_FACES = [(0, 4, 1), (0, 9, 4), (9, 5, 4), (4, 5, 8), (4, 8, 1), (8, 10, 1), (8, 3, 10),
                (5, 3, 8), (5, 2, 3), (2, 7, 3), (7, 10, 3), (7, 6, 10), (7, 11, 6), (11, 0, 6),
                (0, 1, 6), (6, 1, 10), (9, 0, 11), (9, 11, 2), (9, 2, 5), (7, 2, 11)]
_OCTA_VERTS = np.array([(0, 0, -1), (0, 0, 1), (0, -1, 0), (0, 1, 0), (-1, 0, 0), (1, 0, 0)],
                       dtype=np.float64)


def _octa_faces():
  # One face per cube corner: the three axis vertices at squared distance 2 from it,
  # listed in increasing vertex index; corners in lexicographic (-1,+1)^3 order.
  # The reference derives the same table through argwhere/reshape (geopoly.py:111-113):
  # its [3, -1] reshape of the per-corner triples followed by a transpose regroups the
  # 24 hits column-wise, which is reproduced literally here.
  corners = np.array(list(itertools.product([-1, 1], repeat=3)), dtype=np.float64)
  hits = []
  for c in corners:
    for vi, v in enumerate(_OCTA_VERTS):
      if np.sum((c - v) ** 2) == 2:
        hits.append(vi)
  hits = np.array(hits)
  return np.sort(hits.reshape(3, -1).T, axis=1)


def _lattice_weights(v):
  if v < 1:
    raise ValueError(f'v {v} must be >= 1')
  rows = [(i, j, v - i - j) for i in range(v + 1) for j in range(v + 1 - i)]
  return np.array(rows, dtype=np.float64) / v


def _tessellate(base_verts, faces, v, eps):
  if not isinstance(v, int):
    raise ValueError(f'v {v} must an integer')
  bary = _lattice_weights(v)
  pts = []
  for face in faces:
    p = bary @ base_verts[list(face), :]
    pts.append(p / np.sqrt(np.sum(p * p, axis=1, keepdims=True)))
  pts = np.concatenate(pts, axis=0)
  kept = []
  for i in range(pts.shape[0]):
    d2 = np.sum((pts[:i + 1] - pts[i]) ** 2, axis=1)
    # the reference computes ||x||^2+||y||^2-2x.y and clamps at 0; same decisions at 1e-4
    first = int(np.argmax(d2 <= eps))
    if first == i:
      kept.append(i)
  return pts[kept]


def generate_basis(base_shape, angular_tesselation, remove_symmetries=True, eps=1e-4):
  """Returns the [K, 3] float64 basis (the caller transposes it to [3, K])."""
  if base_shape == 'icosahedron':
    verts = _tessellate(_ICOSA_VERTS, _ICOSA_FACES, angular_tesselation, eps)
  elif base_shape == 'octahedron':
    verts = _tessellate(_OCTA_VERTS, _octa_faces(), angular_tesselation, eps)
  else:
    raise ValueError(f'base_shape {base_shape} not supported')
  if remove_symmetries:
    keep = []
    for i in range(verts.shape[0]):
      d2 = np.sum((verts[i:] + verts[i]) ** 2, axis=1)   # distance to antipodes at j >= i
      if np.any(d2 < eps):
        keep.append(i)
    verts = verts[keep]
  return np.ascontiguousarray(verts[:, ::-1])
