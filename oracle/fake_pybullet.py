"""A static, render-only stand-in for the ``pybullet`` module.

Test infrastructure (see oracle/__init__.py).  Implements the names the
reference's simulator.py / observer.py call (SURVEY 8c lists them) so that the
UNMODIFIED reference StackEnv runs end to end without physics: a body stays
exactly where it is placed, has zero velocity and always reports three contact
points, which makes ``Simulator.step`` return at once (simulator.py:213, 239,
337-341).  ``getCameraImage`` is the software z-buffer of oracle/raster_np.py.

Every function accepts and ignores ``physicsClientId`` because the reference
forwards it on every call (simulator.py:57-61).
"""
import os
import re

import numpy as np

from oracle import raster_np as R

GUI, DIRECT, GEOM_BOX, COV_ENABLE_GUI = 1, 2, 3, 1
GEOM_MESH = 5


def load_obj(path):
  verts, tris = [], []
  with open(path) as f:
    for line in f:
      if line.startswith('v '):
        verts.append([float(x) for x in line.split()[1:4]])
      elif line.startswith('f '):
        idx = [int(tok.split('/')[0]) - 1 for tok in line.split()[1:]]
        for k in range(1, len(idx) - 1):
          tris.append([idx[0], idx[k], idx[k + 1]])
  return np.asarray(verts, dtype='float32'), np.asarray(tris, dtype='int32')


def parse_urdf(path):
  """(mesh path, inertial origin xyz) of a single-link URDF like
  stackrl/envs/data/template.urdf."""
  text = open(path).read()
  mesh = re.search(r'<visual.*?<mesh\s+filename="([^"]+)"', text, re.S).group(1)
  origin = re.search(r'<inertial>.*?<origin\s+xyz="([^"]+)"', text, re.S)
  xyz = [float(x) for x in origin.group(1).split()] if origin else [0., 0., 0.]
  if not os.path.isabs(mesh):
    mesh = os.path.join(os.path.dirname(path), mesh)
  return mesh, np.asarray(xyz, dtype='float64')


def box_mesh(half):
  hx, hy, hz = half
  v = np.array([[sx * hx, sy * hy, sz * hz] for sx in (-1, 1) for sy in (-1, 1)
                for sz in (-1, 1)], dtype='float32')
  t = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1],
                [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]],
               dtype='int32')
  return v, t


class _Body(object):
  def __init__(self, verts, tris, pos, orn, com=(0., 0., 0.), visible=True, shape=None):
    self.verts, self.tris = verts, tris
    # (geometry type, dimensions, mesh file, local visual frame position) as
    # pybullet.getVisualShapeData reports them
    self.shape = shape
    self.pos, self.orn = tuple(pos), tuple(orn)
    self.com = np.asarray(com, dtype='float64')   # inertial origin in the link frame
    self.visible = visible
    self.mass, self.inertia = 1.0, (1., 1., 1.)

  def world(self):
    """(rotation, translation) of the VISUAL mesh: the base pose is the pose of
    the inertial frame, the mesh sits at -com inside it."""
    rot = R.quat_matrix(self.orn)
    return rot, np.asarray(self.pos) - rot.dot(self.com)


class FakeBullet(object):
  """One instance = one fake physics server.  Use as ``pybullet`` module."""
  GUI, DIRECT, GEOM_BOX, COV_ENABLE_GUI = GUI, DIRECT, GEOM_BOX, COV_ENABLE_GUI

  def __init__(self):
    self._connected = False
    self._bodies = {}
    self._shapes = {}
    self._next = 0
    self.camera_calls = 0

  # -- server / world ---------------------------------------------------------- #
  def connect(self, mode=DIRECT, **_):
    # a new connection is a fresh server: nothing survives a disconnect
    self._connected = True
    self._bodies, self._shapes = {}, {}
    return 0

  def disconnect(self, physicsClientId=0, **_):
    self._connected = False

  def isConnected(self, physicsClientId=0, **_):
    return self._connected

  def resetSimulation(self, **_):
    self._bodies, self._shapes = {}, {}

  def setTimeStep(self, *a, **_):
    pass

  def setGravity(self, *a, **_):
    pass

  def configureDebugVisualizer(self, *a, **_):
    pass

  def resetDebugVisualizerCamera(self, *a, **_):
    pass

  def stepSimulation(self, **_):
    pass

  # -- bodies ------------------------------------------------------------------ #
  def _new_id(self):
    self._next += 1
    return self._next

  def createCollisionShape(self, *a, **_):
    return -1

  def createVisualShape(self, shapeType=GEOM_BOX, halfExtents=(1, 1, 1), rgbaColor=None,
                        visualFramePosition=(0, 0, 0), **_):
    sid = self._new_id()
    v, t = box_mesh(halfExtents)
    self._shapes[sid] = (v + np.asarray(visualFramePosition, dtype='float32'), t,
                         (GEOM_BOX, tuple(2. * h for h in halfExtents), '',
                          tuple(float(x) for x in visualFramePosition)))
    return sid

  def createMultiBody(self, baseMass=0, baseCollisionShapeIndex=-1,
                      baseVisualShapeIndex=-1, basePosition=(0, 0, 0),
                      baseOrientation=(0, 0, 0, 1), **_):
    bid = self._new_id()
    if baseVisualShapeIndex in self._shapes:
      v, t, shape = self._shapes[baseVisualShapeIndex]
      self._bodies[bid] = _Body(v, t, basePosition, baseOrientation, shape=shape)
    else:
      self._bodies[bid] = _Body(np.zeros((0, 3), 'float32'), np.zeros((0, 3), 'int32'),
                                basePosition, baseOrientation, visible=False)
    return bid

  def loadURDF(self, fileName, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), **_):
    mesh, com = parse_urdf(fileName)
    v, t = load_obj(mesh)
    bid = self._new_id()
    self._bodies[bid] = _Body(v, t, basePosition, baseOrientation, com=com,
                              shape=(GEOM_MESH, (1., 1., 1.), mesh,
                                     tuple(float(-c) for c in com)))
    return bid

  def removeBody(self, bodyUniqueId, **_):
    del self._bodies[bodyUniqueId]

  def resetBasePositionAndOrientation(self, bodyUniqueId, posObj, ornObj, **_):
    b = self._bodies[bodyUniqueId]
    b.pos, b.orn = tuple(posObj), tuple(ornObj)

  def resetBaseVelocity(self, *a, **_):
    pass

  def getBasePositionAndOrientation(self, bodyUniqueId, **_):
    b = self._bodies[bodyUniqueId]
    return b.pos, b.orn

  def getBaseVelocity(self, bodyUniqueId, **_):
    return (0., 0., 0.), (0., 0., 0.)

  def getContactPoints(self, bodyA=None, **_):
    return [None, None, None]

  def getDynamicsInfo(self, bodyUniqueId, linkIndex, **_):
    b = self._bodies[bodyUniqueId]
    return (b.mass, 0.5, b.inertia)

  def changeDynamics(self, bodyUniqueId, linkIndex, mass=None, localInertiaDiagonal=None, **_):
    b = self._bodies[bodyUniqueId]
    if mass is not None:
      b.mass = mass
    if localInertiaDiagonal is not None:
      b.inertia = tuple(localInertiaDiagonal)

  # -- scene queries (what stackrl_b200.observer.PybulletScene reads) ---------------- #
  def getNumBodies(self, **_):
    return len(self._bodies)

  def getBodyUniqueId(self, serialIndex, **_):
    return list(self._bodies.keys())[serialIndex]

  def getVisualShapeData(self, objectUniqueId, **_):
    """[(body, link, geometry type, dimensions, mesh file, local visual frame position,
    local visual frame orientation, rgba)], the visual frame relative to the inertial
    frame getBasePositionAndOrientation reports."""
    b = self._bodies[objectUniqueId]
    if not b.visible or b.shape is None:
      return []
    geom, dims, filename, lpos = b.shape
    return [(objectUniqueId, -1, geom, dims, filename.encode(), lpos, (0., 0., 0., 1.),
             (1., 1., 1., 1.))]

  # -- transforms -------------------------------------------------------------- #
  def getQuaternionFromEuler(self, eulerAngles, **_):
    return R.quat_from_euler(eulerAngles)

  def multiplyTransforms(self, positionA, orientationA, positionB, orientationB, **_):
    p = np.asarray(positionA, dtype='float64') + R.quat_rotate(orientationA, positionB)
    return tuple(p), R.quat_mul(orientationA, orientationB)

  def invertTransform(self, position, orientation, **_):
    qi = R.quat_conj(orientation)
    return tuple(-np.asarray(R.quat_rotate(qi, position))), qi

  def getDifferenceQuaternion(self, quaternionStart, quaternionEnd, **_):
    return R.quat_mul(quaternionEnd, R.quat_conj(quaternionStart))

  # -- camera ------------------------------------------------------------------ #
  def computeViewMatrix(self, cameraEyePosition, cameraTargetPosition, cameraUpVector, **_):
    return R.look_at(cameraEyePosition, cameraTargetPosition, cameraUpVector)

  def computeProjectionMatrix(self, left, right, bottom, top, nearVal, farVal, **_):
    return R.frustum(left, right, bottom, top, nearVal, farVal)

  def scene(self):
    """[(verts, tris, rot, pos)] of every visible body."""
    out = []
    for b in self._bodies.values():
      if b.visible and len(b.tris):
        rot, pos = b.world()
        out.append((b.verts, b.tris, rot, pos))
    return out

  def getCameraImage(self, width, height, viewMatrix, projectionMatrix, **_):
    self.camera_calls += 1
    depth = R.render_depth(viewMatrix, projectionMatrix, height, width, self.scene())
    return width, height, None, depth, None
