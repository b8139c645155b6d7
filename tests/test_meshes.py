"""On-disk rock formats: Wavefront .obj + single-link URDF in the
layout of stackrl/envs/data/template.urdf, the ``generated`` file listing
(stackrl/envs/data/__init__.py:39-83) and the packed mesh-bank cache.  CPU only; the
files are written by the test in the reference's format (trimesh export: ``v`` / ``f``
lines, 1-based indices)."""
import os

import numpy as np

from stackrl_b200 import meshes

URDF = """<?xml version="1.0"?>
<robot name="{name}">
  <link name="link">
    <contact>
      <lateral_friction value="0.6"/>
    </contact>
    <inertial>
      <origin xyz="{com}" rpy="0 0 0"/>
      <mass value = "0.25"/>
      <inertia ixx="1e-4" ixy="0" ixz="0" iyy="1e-4" iyz="0" izz="1e-4" />
    </inertial>
    <visual name="visual">
      <geometry>
        <mesh filename="{name}.obj"/>
      </geometry>
    </visual>
    <collision name="collision">
      <geometry>
        <mesh filename="{name}.obj"/>
      </geometry>
    </collision>
  </link>
</robot>
"""


def _write_rock(directory, name, verts, tris, com=(0., 0., 0.), style='plain'):
  with open(os.path.join(directory, name + '.obj'), 'w') as f:
    f.write('# https://github.com/mikedh/trimesh\n')
    for v in verts:
      f.write('v {:.8f} {:.8f} {:.8f}\n'.format(*v))
    if style == 'normals':
      f.write('vn 0 0 1\n')
    for t in tris:
      if style == 'normals':
        f.write('f {}//1 {}//1 {}//1\n'.format(*(t + 1)))
      elif style == 'negative':
        f.write('f {} {} {}\n'.format(*(t - len(verts))))
      else:
        f.write('f {} {} {}\n'.format(*(t + 1)))
  with open(os.path.join(directory, name + '.urdf'), 'w') as f:
    f.write(URDF.format(name=name, com=' '.join(repr(float(c)) for c in com)))


def test_obj_and_urdf_round_trip(tmp_path):
  verts, tris = meshes.icosphere(1)
  verts = (verts * 0.05).astype('float32')
  for style in ('plain', 'normals', 'negative'):
    _write_rock(str(tmp_path), '50_%s' % style, verts, tris, com=(0.00626, -0.00097, 0.0019),
                style=style)
    v, t, com = meshes.load_urdf(str(tmp_path / ('50_%s.urdf' % style)))
    assert v.dtype == np.float32 and t.dtype == np.int32
    np.testing.assert_allclose(v, verts, rtol=0, atol=1e-8)
    np.testing.assert_array_equal(t, tris)
    np.testing.assert_array_equal(com, [0.00626, -0.00097, 0.0019])


def test_polygon_faces_are_fan_triangulated(tmp_path):
  with open(tmp_path / 'quad.obj', 'w') as f:
    f.write('v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nf 1 2 3 4\n')
  v, t = meshes.load_obj(str(tmp_path / 'quad.obj'))
  assert v.shape == (4, 3)
  np.testing.assert_array_equal(t, [[0, 1, 2], [0, 2, 3]])


def test_generated_listing_and_bank_cache(tmp_path):
  verts, tris = meshes.icosphere(0)
  os.makedirs(tmp_path / 'test')
  os.makedirs(tmp_path / 'compat')
  for name in ('50_001', '50_000', '95_003', '20_000'):
    _write_rock(str(tmp_path), name, verts.astype('float32') * 0.04, tris, com=(0.001, 0., 0.))
  _write_rock(str(tmp_path / 'test'), '50_900', verts.astype('float32'), tris)
  _write_rock(str(tmp_path / 'compat'), 'old7', verts.astype('float32'), tris)
  base = lambda files: [os.path.basename(f) for f in files]
  # the registered environments' pattern (envs/stack/__init__.py): irregularity 50-95
  assert base(meshes.generated(str(tmp_path), '[5-9]?')) == ['50_000.urdf', '50_001.urdf',
                                                            '95_003.urdf']
  assert base(meshes.generated(str(tmp_path), '50', test=True)) == ['50_900.urdf']
  assert base(meshes.generated(str(tmp_path), 'old')) == ['old7.urdf']      # compat fallback
  assert len(meshes.generated(str(tmp_path))) == 4
  bank = meshes.MeshBank.from_urdfs(meshes.generated(str(tmp_path), '[5-9]?'))
  assert len(bank) == 3 and bank.names['95_003'] == 2
  assert bank.add_urdf(str(tmp_path / '50_001.urdf')) == 1                  # known name: no copy
  bank.save(str(tmp_path / 'bank'))
  again = meshes.MeshBank.load(str(tmp_path / 'bank'))
  assert again.ranges == bank.ranges and again.names == bank.names
  np.testing.assert_array_equal(again.verts, bank.verts)
  np.testing.assert_array_equal(again.tris, bank.tris)
  np.testing.assert_array_equal(np.asarray(again.coms), np.asarray(bank.coms))
