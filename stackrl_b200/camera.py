"""Camera and pose algebra of the observation path, float64 on the host.

The reference builds its cameras with pybullet's ``computeViewMatrix`` /
``computeProjectionMatrix`` and quaternion helpers (observer.py:84-141,
148-237).  Those are plain GL / Hamilton conventions; they are restated here so
the GPU observer does not need a physics server to place its cameras.  Matrices
are returned as 16-tuples in column-major order, like pybullet returns them.
"""
import math

import numpy as np

FAR = 10 ** 3          # Observer.far (observer.py:6)
# The float32 elevation formulas (observer.py:259-260, 274-275) subtract two numbers
# of magnitude ~FAR, so every height is a multiple of ulp(FAR) = 2^-14 m.
HEIGHT_QUANTUM_LOG2 = -14


def view_matrix(eye, target, up):
  """GL look-at (computeViewMatrix)."""
  ex, ey, ez = (float(c) for c in eye)
  fx, fy, fz = (float(t) - e for t, e in zip(target, (ex, ey, ez)))
  inv = 1.0 / math.sqrt(fx * fx + fy * fy + fz * fz)
  fx, fy, fz = fx * inv, fy * inv, fz * inv
  ux, uy, uz = (float(c) for c in up)
  inv = 1.0 / math.sqrt(ux * ux + uy * uy + uz * uz)
  ux, uy, uz = ux * inv, uy * inv, uz * inv
  sx, sy, sz = fy * uz - fz * uy, fz * ux - fx * uz, fx * uy - fy * ux
  inv = 1.0 / math.sqrt(sx * sx + sy * sy + sz * sz)
  sx, sy, sz = sx * inv, sy * inv, sz * inv
  ux, uy, uz = sy * fz - sz * fy, sz * fx - sx * fz, sx * fy - sy * fx
  return (sx, ux, -fx, 0.0,
          sy, uy, -fy, 0.0,
          sz, uz, -fz, 0.0,
          -(sx * ex + sy * ey + sz * ez), -(ux * ex + uy * ey + uz * ez),
          fx * ex + fy * ey + fz * ez, 1.0)


def projection_matrix(left, right, bottom, top, near, far):
  """glFrustum (computeProjectionMatrix with explicit planes)."""
  return (2.0 * near / (right - left), 0.0, 0.0, 0.0,
          0.0, 2.0 * near / (top - bottom), 0.0, 0.0,
          (right + left) / (right - left), (top + bottom) / (top - bottom),
          -(far + near) / (far - near), -1.0,
          0.0, 0.0, -2.0 * far * near / (far - near), 0.0)


def quaternion_from_yaw(angle):
  """[x, y, z, w] of a rotation by ``angle`` about +z
  (getQuaternionFromEuler([0, 0, angle]))."""
  return (0.0, 0.0, math.sin(0.5 * angle), math.cos(0.5 * angle))


def quaternion_inverse(q):
  return (-q[0], -q[1], -q[2], q[3])


def quaternion_multiply(a, b):
  ax, ay, az, aw = a
  bx, by, bz, bw = b
  return (aw * bx + ax * bw + ay * bz - az * by,
          aw * by - ax * bz + ay * bw + az * bx,
          aw * bz + ax * by - ay * bx + az * bw,
          aw * bw - ax * bx - ay * by - az * bz)


def rotation_matrix(q):
  """Row-major 3x3 rotation of the unit quaternion [x, y, z, w]."""
  x, y, z, w = (float(c) for c in q)
  s = 2.0 / (x * x + y * y + z * z + w * w)
  return np.array([
    [1.0 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w)],
    [s * (x * y + z * w), 1.0 - s * (x * x + z * z), s * (y * z - x * w)],
    [s * (x * z - y * w), s * (y * z + x * w), 1.0 - s * (x * x + y * y)]])


def rotate(q, v):
  return tuple(rotation_matrix(q).dot(np.asarray(v, dtype='float64')))


class ObserverGeometry(object):
  """Sizes and cameras of one Observer (observer.py:51-141), shared by the
  single-environment and the batched GPU observers."""

  def __init__(self, overhead_resolution=192, object_resolution=32,
               pixel_size=2. ** (-8), max_z=1, orientation_freedom=0):
    if np.isscalar(pixel_size):
      self.pixel_h = self.pixel_w = pixel_size
    else:
      self.pixel_h, self.pixel_w = pixel_size[0], pixel_size[1]
    if np.isscalar(overhead_resolution):
      self.overhead_h = self.overhead_w = int(overhead_resolution)
    else:
      self.overhead_h, self.overhead_w = (int(r) for r in overhead_resolution[:2])
    self.overhead_x = self.overhead_h * self.pixel_h
    self.overhead_y = self.overhead_w * self.pixel_w
    self.overhead_z = max_z
    if np.isscalar(object_resolution):
      self.object_h = self.object_w = int(object_resolution)
    else:
      self.object_h, self.object_w = (int(r) for r in object_resolution[:2])
    self.object_x = self.object_h * self.pixel_h
    self.object_y = self.object_w * self.pixel_w
    self.object_z = max(self.object_x, self.object_y)
    far = FAR
    # Overhead camera: above the centre of the observable area, looking down,
    # image rows along +x (observer.py:84-104).
    self.overhead_view = view_matrix(
      (self.overhead_x / 2, self.overhead_y / 2, far),
      (self.overhead_x / 2, self.overhead_y / 2, 0), (-1, 0, 0))
    self.overhead_projection = projection_matrix(
      -self.overhead_y / 2, self.overhead_y / 2, -self.overhead_x / 2,
      self.overhead_x / 2, far - self.overhead_z, far)
    # Object camera: below the spawned rock, looking up (observer.py:112-119).
    self.object_projection = projection_matrix(
      -self.object_y / 2, self.object_y / 2, -self.object_x / 2, self.object_x / 2,
      far - self.object_z / 2, far + self.object_z / 2)
    # One up-vector per orientation and the orientation handed back to the
    # simulator for it (observer.py:128-141).
    self.n_orientations = 2 ** orientation_freedom
    self.up_vectors, self.orientations = [], []
    for i in range(self.n_orientations):
      q = quaternion_from_yaw(i * 2 * np.pi / self.n_orientations)
      self.up_vectors.append(rotate(q, (-1, 0, 0)))
      self.orientations.append(quaternion_inverse(q))

  def object_view(self, pose, k=0):
    """View matrix of the camera under a rock at ``pose`` for orientation k
    (observer.py:148-164, 167-185)."""
    position, orientation = pose
    eye = tuple(np.asarray(position, dtype='float64') +
                np.asarray(rotate(orientation, (0, 0, -FAR))))
    up = rotate(orientation, self.up_vectors[k])
    return view_matrix(eye, position, up)

  @property
  def shape(self):
    return (self.overhead_h, self.overhead_w), (self.object_h, self.object_w)

  @property
  def size(self):
    return self.overhead_x, self.overhead_y, self.overhead_z

  @property
  def max_z(self):
    """Highest z of a rock that is still fully visible (observer.py:379-382)."""
    return self.overhead_z - self.object_z
