"""Rock meshes: Wavefront .obj parsing, synthetic star-convex rocks and the
device mesh bank the rasteriser indexes.

The reference ships 10 005 convex rocks as .obj/.urdf pairs
(stackrl/envs/data/generated, ~86 faces each) produced by an offline trimesh
generator (envs/data/generator.py); only their plain ``v``/``f`` text format is
needed here.  The synthetic rocks of BASELINE config 3 are
icosphere subdivisions with a seeded radial scale.
"""
import glob
import os
import re

import numpy as np


def load_obj(path):
  """(verts [V,3] float32, tris [T,3] int32) of a triangle/polygon .obj."""
  verts, tris = [], []
  with open(path) as f:
    for line in f:
      tok = line.split()
      if not tok:
        continue
      if tok[0] == 'v':
        verts.append((float(tok[1]), float(tok[2]), float(tok[3])))
      elif tok[0] == 'f':
        idx = [int(t.split('/')[0]) for t in tok[1:]]
        idx = [i - 1 if i > 0 else len(verts) + i for i in idx]
        for k in range(1, len(idx) - 1):          # fan-triangulate polygons
          tris.append((idx[0], idx[k], idx[k + 1]))
  return np.asarray(verts, dtype='float32').reshape(-1, 3), \
    np.asarray(tris, dtype='int32').reshape(-1, 3)


def load_urdf(path):
  """(verts, tris, inertial origin xyz) of a single-link URDF in the layout of
  stackrl/envs/data/template.urdf (visual = collision = one mesh, no scale)."""
  text = open(path).read()
  mesh = re.search(r'<visual.*?<mesh\s+filename="([^"]+)"', text, re.S).group(1)
  if not os.path.isabs(mesh):
    mesh = os.path.join(os.path.dirname(path), mesh)
  origin = re.search(r'<inertial>.*?<origin\s+xyz="([^"]+)"', text, re.S)
  com = [float(x) for x in origin.group(1).split()] if origin else [0., 0., 0.]
  verts, tris = load_obj(mesh)
  return verts, tris, np.asarray(com, dtype='float64')


def generated(directory, name=None, test=False):
  """URDF files of a rock set, the reference's ``stackrl.envs.data.generated``
  (stackrl/envs/data/__init__.py:39-83) with the data directory given explicitly
  (the assets stay in the reference's package): ``<directory>/[test/]<name>_*.urdf``,
  falling back to ``<directory>/compat/<name>*.urdf`` for the old names.  The
  registered environments use ``name='[5-9]?'`` (envs/stack/__init__.py:3-24).
  Sorted, so that mesh ids are reproducible."""
  sub = os.path.join(directory, 'test') if test else directory
  pattern = '{}_*.urdf'.format(name) if name is not None else '*.urdf'
  files = glob.glob(os.path.join(sub, pattern))
  if not files and name is not None:
    files = glob.glob(os.path.join(directory, 'compat', '{}*.urdf'.format(name)))
  return sorted(files)


def icosphere(subdivisions):
  """Unit icosphere: 20 * 4^s triangles, 10 * 4^s + 2 vertices."""
  t = (1.0 + 5.0 ** 0.5) / 2.0
  verts = [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t),
           (0, -1, -t), (0, 1, -t), (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]
  verts = [np.asarray(v, dtype='float64') / np.linalg.norm(v) for v in verts]
  faces = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9),
           (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8), (3, 9, 4), (3, 4, 2),
           (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10),
           (8, 6, 7), (9, 8, 1)]
  for _ in range(subdivisions):
    cache, out = {}, []

    def mid(a, b):
      key = (a, b) if a < b else (b, a)
      if key not in cache:
        m = verts[a] + verts[b]
        verts.append(m / np.linalg.norm(m))
        cache[key] = len(verts) - 1
      return cache[key]
    for a, b, c in faces:
      ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
      out += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
    faces = out
  return np.asarray(verts, dtype='float64'), np.asarray(faces, dtype='int32')


def geodesic_icosphere(frequency):
  """Unit class-I geodesic sphere: every icosahedron face cut into frequency^2
  triangles (20 f^2 triangles, 10 f^2 + 2 vertices; f = 10 gives the 2 000
  triangles / 1 002 vertices of BASELINE config 3)."""
  base, faces = icosphere(0)
  f = int(frequency)
  index, verts, tris = {}, [], []

  def vertex(a, b, c, i, j):
    # barycentric grid point of face (a, b, c); the key is the same for both faces
    # that share an edge or a corner
    w = sorted(((a, f - i - j), (b, i), (c, j)))
    key = tuple((n, k) for n, k in w if k)
    if key not in index:
      p = ((f - i - j) * base[a] + i * base[b] + j * base[c]) / f
      verts.append(p / np.linalg.norm(p))
      index[key] = len(verts) - 1
    return index[key]
  for a, b, c in faces:
    for i in range(f):
      for j in range(f - i):
        tris.append((vertex(a, b, c, i, j), vertex(a, b, c, i + 1, j), vertex(a, b, c, i, j + 1)))
        if i + j < f - 1:
          tris.append((vertex(a, b, c, i + 1, j), vertex(a, b, c, i + 1, j + 1),
                       vertex(a, b, c, i, j + 1)))
  return np.asarray(verts, dtype='float64'), np.asarray(tris, dtype='int32')


def synthetic_rocks(seed, count, subdivisions=3, max_dimension=0.16, smooth=0.6,
                    frequency=None):
  """``count`` star-convex rocks: an icosphere whose
  vertices are pushed radially by a seeded triangular(0.1, 0.4, 1.0) factor
  blended with a per-rock ellipsoid, scaled to fit a ball of ``max_dimension``.
  subdivisions=3 gives 1280 triangles / 642 vertices, 4 gives 5120 / 2562;
  ``frequency`` (overrides ``subdivisions``) uses the geodesic sphere instead:
  frequency=10 gives config 3's 2 000 triangles / 1 002 vertices.
  Returns (verts [count, V, 3] float32, tris [T, 3] int32 shared by all)."""
  base, tris = geodesic_icosphere(frequency) if frequency else icosphere(subdivisions)
  rng = np.random.default_rng(seed)
  axes = rng.uniform(0.55, 1.0, (count, 1, 3))
  radial = rng.triangular(0.1, 0.4, 1.0, (count, base.shape[0], 1))
  scale = smooth + (1 - smooth) * radial
  v = base[None] * axes * scale
  v *= (0.5 * max_dimension) / np.linalg.norm(v, axis=-1).max(axis=1)[:, None, None]
  return v.astype('float32'), tris


class MeshBank(object):
  """Meshes packed for the rasteriser: one vertex buffer, one index buffer
  (indices local to each mesh) and per-mesh ranges.  Host arrays plus lazily
  uploaded device copies."""

  def __init__(self):
    self._verts, self._tris = [], []
    self.ranges = []            # (vert_begin, vert_count, tri_begin, tri_count)
    self.coms = []              # inertial origin of each mesh (link frame)
    self.names = {}
    self._nv = self._nt = 0
    self._device = None

  def add(self, verts, tris, com=(0., 0., 0.), name=None):
    verts = np.ascontiguousarray(verts, dtype='float32').reshape(-1, 3)
    tris = np.ascontiguousarray(tris, dtype='int32').reshape(-1, 3)
    if len(tris) and (tris.min() < 0 or tris.max() >= len(verts)):
      raise ValueError('triangle index outside the mesh')
    self._verts.append(verts)
    self._tris.append(tris)
    self.ranges.append((self._nv, len(verts), self._nt, len(tris)))
    self.coms.append(np.asarray(com, dtype='float64'))
    self._nv += len(verts)
    self._nt += len(tris)
    self._device = None
    index = len(self.ranges) - 1
    if name is not None:
      self.names[name] = index
    return index

  def add_urdf(self, path):
    name = os.path.splitext(os.path.basename(path))[0]
    if name in self.names:
      return self.names[name]
    verts, tris, com = load_urdf(path)
    return self.add(verts, tris, com, name=name)

  @classmethod
  def from_urdfs(cls, paths):
    """Bank of the given URDF files, mesh id = position in ``paths``."""
    bank = cls()
    for path in paths:
      bank.add_urdf(path)
    return bank

  def save(self, path):
    """Packed on-disk form (one .npz: vertex buffer, index buffer, ranges, inertial
    origins, names): reloading the reference's 10 005 rocks takes milliseconds
    instead of re-parsing 20 010 text files."""
    names = [''] * len(self.ranges)
    for name, index in self.names.items():
      names[index] = name
    np.savez_compressed(
      path, verts=self.verts, tris=self.tris,
      ranges=np.asarray(self.ranges, dtype='int64').reshape(-1, 4),
      coms=np.asarray(self.coms, dtype='float64').reshape(-1, 3),
      names=np.asarray(names, dtype='U'))

  @classmethod
  def load(cls, path):
    """Inverse of ``save``."""
    data = np.load(path if str(path).endswith('.npz') else str(path) + '.npz')
    verts, tris = data['verts'], data['tris']      # (an NpzFile decompresses per access)
    bank = cls()
    for (vb, vn, tb, tn), com, name in zip(data['ranges'], data['coms'], data['names']):
      bank.add(verts[vb:vb + vn], tris[tb:tb + tn], com, name=str(name) or None)
    return bank

  def __len__(self):
    return len(self.ranges)

  @property
  def verts(self):
    return np.concatenate(self._verts) if self._verts else np.zeros((0, 3), 'float32')

  @property
  def tris(self):
    return np.concatenate(self._tris) if self._tris else np.zeros((0, 3), 'int32')

  def device(self, device):
    """(verts, tris) CUDA tensors (cached per device)."""
    import torch
    if self._device is None or self._device[0] != device:
      self._device = (device, torch.from_numpy(self.verts).to(device),
                      torch.from_numpy(self.tris).to(device))
    return self._device[1], self._device[2]
