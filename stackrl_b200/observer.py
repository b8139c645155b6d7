"""GPU observers: heightmap capture behind the reference ``Observer`` interface.

Mirrors /root/reference/stackrl/envs/stack/observer.py.  Two classes:

* ``gpu_observer_class(base)`` -> a drop-in ``Observer`` (same constructor,
  ``__call__``, ``state``, ``shape``, ``size``, ``max_z``, ``num_objects``,
  ``pixel_to_xy``, ``xy_to_pixel``, ``pose``, ``visualize``) for ONE
  environment.  ``base`` is the reference's own ``Observer`` class when the
  object has to pass ``Rewarder``'s isinstance gate (rewarder.py:58-63); see
  INTEGRATION.md.  Where the reference asks pybullet for a camera image
  (observer.py:252-257, 267-272, 283-288) this class rasterises the simulator's
  scene with the sm_100a kernel (capi.raster) and fuses the depth->elevation
  conversion and the column mirror into the same launch.
* ``BatchedObserver``: the same for E environments at once with device-resident
  maps, instance tables and cameras (what BatchedStackEnv and bench.py use).

The simulator only has to expose its visual scene: either a ``scene()`` method
returning ``[(verts [V,3], tris [T,3], rot [3,3], pos [3]), ...]`` or pybullet's
own query API (``PybulletScene`` adapts the latter).
"""
import numpy as np
import torch

from stackrl_b200 import camera, capi
from stackrl_b200.camera import FAR


def _device():
  if not torch.cuda.is_available():
    raise RuntimeError('stackrl_b200 needs a CUDA device (no CPU fallback)')
  return torch.device('cuda', torch.cuda.current_device())


class PybulletScene(object):
  """Visual scene of a live pybullet client as rasteriser instances.

  Written against pybullet's documented query API (getNumBodies,
  getBodyUniqueId, getVisualShapeData, getBasePositionAndOrientation); pybullet
  is not installable in the build image, so this adapter is exercised through the
  fake backend's restatement of that surface (oracle/fake_pybullet.py;
  tests/test_observer_host.py)."""

  GEOM_BOX, GEOM_MESH = 3, 5

  def __init__(self, sim):
    self._sim = sim
    self._meshes = {}

  def _mesh(self, shape):
    from stackrl_b200 import meshes
    geom, dims, filename = shape[2], shape[3], shape[4]
    filename = filename.decode() if isinstance(filename, bytes) else filename
    key = (geom, tuple(dims), filename)
    if key not in self._meshes:
      if geom == self.GEOM_MESH:
        v, t = meshes.load_obj(filename)
        v = v * np.asarray(dims, dtype='float32')
      elif geom == self.GEOM_BOX:
        hx, hy, hz = (0.5 * d for d in dims)
        v = np.array([[sx * hx, sy * hy, sz * hz] for sx in (-1, 1) for sy in (-1, 1)
                      for sz in (-1, 1)], dtype='float32')
        t = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1],
                      [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]],
                     dtype='int32')
      else:
        raise NotImplementedError('visual geometry type {}'.format(geom))
      self._meshes[key] = (v, t)
    return self._meshes[key]

  def __call__(self):
    sim, out = self._sim, []
    for k in range(sim.getNumBodies()):
      body = sim.getBodyUniqueId(k)
      pos, orn = sim.getBasePositionAndOrientation(body)
      for shape in sim.getVisualShapeData(body):
        v, t = self._mesh(shape)
        # visual frame relative to the (inertial) base frame
        lpos, lorn = shape[5], shape[6]
        rot = camera.rotation_matrix(orn)
        out.append((v, t, rot.dot(camera.rotation_matrix(lorn)),
                    np.asarray(pos) + rot.dot(np.asarray(lpos))))
    return out


class GpuCamera(object):
  """Seam b2: a simulator proxy whose ``getCameraImage`` is the CUDA
  rasteriser, so that the reference's UNMODIFIED ``Observer`` -- which asks its
  simulator for ``getCameraImage(width=, height=, viewMatrix=, projectionMatrix=)``
  and applies the depth -> elevation arithmetic itself (observer.py:252-260,
  267-277) -- runs on the GPU path without a line changed:

      env = StackEnv(..., simulator=lambda **kw: GpuCamera(Simulator(**kw)))

  Every other attribute (camera matrices, transforms, ``new_pose``,
  ``has_new_object``, the physics calls) is forwarded to the wrapped simulator.  The
  scene comes from the simulator's ``scene()`` when it has one, else from pybullet's
  own query API (``PybulletScene``).  Returns pybullet's 5-tuple
  ``(width, height, rgb, depth, segmentation)`` with ``depth`` a fresh float32
  [height, width] array in [0, 1] (GL convention, background 1) and ``rgb`` /
  ``segmentation`` ``None`` (the reference discards them)."""

  def __init__(self, simulator, device=None):
    object.__setattr__(self, '_sim', simulator)
    object.__setattr__(self, '_scene', simulator.scene if hasattr(simulator, 'scene')
                       else PybulletScene(simulator))
    object.__setattr__(self, '_device', device)

  def __getattr__(self, name):
    return getattr(object.__getattribute__(self, '_sim'), name)

  def __setattr__(self, name, value):
    setattr(object.__getattribute__(self, '_sim'), name, value)

  def getCameraImage(self, width, height, viewMatrix, projectionMatrix, **_):
    dev = object.__getattribute__(self, '_device') or _device()
    bodies = object.__getattribute__(self, '_scene')()
    verts, tris, inst = _instances(bodies)
    job = _job(viewMatrix, projectionMatrix, 0, len(inst), 0.)
    depth = capi.raster(torch.from_numpy(verts).to(dev), torch.from_numpy(tris).to(dev), inst,
                        job, int(height), int(width), capi.RASTER_DEPTH)
    return int(width), int(height), None, depth[0].cpu().numpy(), None


def gpu_simulator_class(base):
  """Seam b2 as a subclass: ``base`` is the reference's ``Simulator``
  (simulator.py); the returned class is a ``Simulator`` (so it passes Rewarder's
  isinstance gate, rewarder.py:52-57) whose ``getCameraImage`` -- which the base
  class forwards to pybullet's TinyRenderer (simulator.py:57-61) -- is the CUDA
  rasteriser.  ``StackEnv(simulator=gpu_simulator_class(Simulator))`` then runs the
  reference's unmodified Observer, Rewarder and env on GPU depth images."""

  class GpuCameraSimulator(base):
    def getCameraImage(self, width, height, viewMatrix, projectionMatrix, **_):
      if not hasattr(self, '_gpu_scene'):
        self._gpu_scene = PybulletScene(self)      # pybullet's own scene queries
      dev = _device()
      verts, tris, inst = _instances(self._gpu_scene())
      job = _job(viewMatrix, projectionMatrix, 0, len(inst), 0.)
      depth = capi.raster(torch.from_numpy(verts).to(dev), torch.from_numpy(tris).to(dev),
                          inst, job, int(height), int(width), capi.RASTER_DEPTH)
      return int(width), int(height), None, depth[0].cpu().numpy(), None

  return GpuCameraSimulator


def _instances(bodies):
  """Concatenated (verts, tris) + INSTANCE_DTYPE rows for world-placed bodies."""
  inst = np.zeros(len(bodies), dtype=capi.INSTANCE_DTYPE)
  verts, tris, nv, nt = [], [], 0, 0
  for k, (v, t, rot, pos) in enumerate(bodies):
    v = np.ascontiguousarray(v, dtype='float32').reshape(-1, 3)
    t = np.ascontiguousarray(t, dtype='int32').reshape(-1, 3)
    inst[k]['rot'] = np.asarray(rot, dtype='float64').ravel()
    inst[k]['pos'] = np.asarray(pos, dtype='float64').ravel()
    inst[k]['vert_begin'], inst[k]['vert_count'] = nv, len(v)
    inst[k]['tri_begin'], inst[k]['tri_count'] = nt, len(t)
    verts.append(v)
    tris.append(t)
    nv += len(v)
    nt += len(t)
  verts = np.concatenate(verts) if verts else np.zeros((0, 3), 'float32')
  tris = np.concatenate(tris) if tris else np.zeros((0, 3), 'int32')
  return verts, tris, inst


def _job(view, proj, inst_begin, inst_count, zrange):
  job = np.zeros(1, dtype=capi.JOB_DTYPE)
  job['view'], job['proj'] = view, proj
  job['inst_begin'], job['inst_count'], job['zrange'] = inst_begin, inst_count, zrange
  return job


def gpu_observer_class(base=object):
  """Returns a GPU ``Observer`` class deriving from ``base`` (pass the
  reference's ``stackrl.envs.stack.observer.Observer`` for a drop-in that
  satisfies Rewarder's isinstance check)."""

  class GpuObserver(base):
    far = FAR

    def __init__(self, simulator, overhead_resolution=192, object_resolution=32,
                 pixel_size=2. ** (-8), max_z=1, object_pose=None,
                 orientation_freedom=0):
      # (the reference's constructor is deliberately not called: it would ask
      # the simulator for pybullet camera matrices this class does not need)
      if hasattr(simulator, 'scene'):
        self._scene = simulator.scene
      elif hasattr(simulator, 'getVisualShapeData'):
        self._scene = PybulletScene(simulator)
      else:
        raise TypeError('simulator must expose scene() or the pybullet query API')
      self._sim = simulator
      self._geo = g = camera.ObserverGeometry(
        overhead_resolution, object_resolution, pixel_size, max_z, orientation_freedom)
      # attribute names of the reference class, for code that peeks at them
      self._pixel_h, self._pixel_w = g.pixel_h, g.pixel_w
      self._overhead_h, self._overhead_w = g.overhead_h, g.overhead_w
      self._overhead_x, self._overhead_y, self._overhead_z = g.overhead_x, g.overhead_y, g.overhead_z
      self._object_h, self._object_w = g.object_h, g.object_w
      self._object_x, self._object_y, self._object_z = g.object_x, g.object_y, g.object_z
      if object_pose is None:
        if hasattr(simulator, 'new_pose'):
          if not isinstance(simulator.new_pose, list):
            object_pose = simulator.new_pose
        else:
          raise ValueError(
            "If object_pose is not provided, simulator must have 'new_pose' attribute.")
      self._object_pose = object_pose
      self._multi_view = g.n_orientations > 1
      self._multi_object = object_pose is None
      self._object_orientations = [] if self._multi_view else None
      self._object_indexes = [] if self._multi_object else None
      # The CUDA device is taken at the first capture, not here: the constructor
      # has to work wherever the reference's own does (e.g. to pass Rewarder's
      # isinstance gate, rewarder.py:58-63, in a process that only builds the env).
      self._dev_cache = None
      self._wall_d = None
      self._rock_d = None
      self._overhead_map = np.zeros((g.overhead_h, g.overhead_w), dtype='float32')
      if self._multi_view or self._multi_object:
        self._object_map = []
      else:
        self._object_map = np.zeros((g.object_h, g.object_w), dtype='float32')

    @property
    def _dev(self):
      if self._dev_cache is None:
        self._dev_cache = _device()
      return self._dev_cache

    # -- capture ---------------------------------------------------------------- #
    def _render(self, jobs, rows, cols, mode, bodies):
      verts, tris, inst = _instances(bodies)
      jobs = np.concatenate(jobs)
      jobs['inst_begin'], jobs['inst_count'] = 0, len(inst)
      return capi.raster(torch.from_numpy(verts).to(self._dev),
                         torch.from_numpy(tris).to(self._dev), inst, jobs, rows, cols, mode,
                         far_plane=self.far)

    def __call__(self):
      """Capture new elevation maps (observer.py:249-352)."""
      g = self._geo
      bodies = self._scene()
      self._wall_d = self._render(
        [_job(g.overhead_view, g.overhead_projection, 0, 0, g.overhead_z)],
        g.overhead_h, g.overhead_w, capi.RASTER_WALL, bodies)
      self._overhead_map = self._wall_d[0].cpu().numpy()
      if self._sim.has_new_object or not getattr(self, '_last_new_poses', None):
        if self._multi_object:
          self._last_new_poses = self._sim.new_pose
          poses = list(self._last_new_poses)
        else:
          poses = [self._object_pose]
        jobs, orientations, indexes = [], [], []
        for i, pose in enumerate(poses):
          for k in range(g.n_orientations):
            jobs.append(_job(g.object_view(pose, k), g.object_projection, 0, 0, g.object_z))
            orientations.append(g.orientations[k])
            indexes.append(i)
        if jobs:
          self._rock_d = self._render(jobs, g.object_h, g.object_w, capi.RASTER_ROCK, bodies)
          maps = list(self._rock_d.cpu().numpy())
        else:
          self._rock_d, maps = None, []
        if self._multi_view or self._multi_object:
          self._object_map = maps
          if self._multi_view:
            self._object_orientations = orientations
          if self._multi_object:
            self._object_indexes = indexes
        else:
          self._object_map = maps[0]
      elif self._multi_object:
        # No new object: drop the views of the objects that were used up and
        # renumber the rest (observer.py:329-352).
        new_poses = self._sim.new_pose
        renumber = [new_poses.index(p) if p in new_poses else None
                    for p in self._last_new_poses]
        keep = [j for j, i in enumerate(self._object_indexes) if renumber[i] is not None]
        self._object_map = [self._object_map[j] for j in keep]
        if self._object_orientations is not None:
          self._object_orientations = [self._object_orientations[j] for j in keep]
        self._object_indexes = [renumber[self._object_indexes[j]] for j in keep]
        if self._rock_d is not None:
          self._rock_d = self._rock_d[torch.as_tensor(keep, device=self._dev, dtype=torch.long)] \
            if keep else None
        self._last_new_poses = new_poses

    # -- reference properties ---------------------------------------------------- #
    @property
    def size(self):
      return self._geo.size

    @property
    def shape(self):
      return self._geo.shape

    @property
    def state(self):
      return self._overhead_map, self._object_map

    @property
    def num_objects(self):
      if isinstance(self._object_map, list):
        return len(self._object_map)
      return int(np.any(self._object_map))

    @property
    def max_z(self):
      return self._geo.max_z

    def pixel_to_xy(self, pixel):
      return pixel[0] * self._pixel_h, pixel[1] * self._pixel_w

    def xy_to_pixel(self, position):
      return position[0] // self._pixel_h, position[1] // self._pixel_w

    def pose(self, pixel, index=None):
      """Placement pose for the rock map overlapped at ``pixel``
      (observer.py:392-421); the drop height is the single-position max-plus
      evaluated by the GPU kernel on the device copies of the maps."""
      x, y = self.pixel_to_xy(pixel)
      view = 0 if index is None else int(index)
      picks = torch.tensor([[view, int(pixel[0]), int(pixel[1])]], dtype=torch.int32,
                           device=self._dev)
      rocks = self._rock_d[None].contiguous()            # [1, views, h, w]
      z = capi.drop_height_f32(self._wall_d, rocks, picks, threshold=10 ** (-4))
      z = z.cpu().numpy()[0]
      x += self._object_x / 2
      y += self._object_y / 2
      z -= self._object_z / 2
      ret = {'position': (x, y, z)}
      if index is not None:
        if self._object_orientations:
          ret['orientation'] = self._object_orientations[index]
        if self._object_indexes:
          ret['index'] = self._object_indexes[index]
      return ret

    def visualize(self, **kwargs):
      if hasattr(self._sim, 'draw_rectangle'):
        self._sim.draw_rectangle(self.size, **kwargs)

  return GpuObserver


GpuObserver = gpu_observer_class(object)


class BatchedObserver(object):
  """E environments observed at once; everything stays on the GPU.

  ``bank`` is the MeshBank of rock meshes.  Per environment the observer keeps a
  fixed-capacity table of placed instances on the device plus the episode
  bookkeeping of ``srl_env_state`` (rock order, cursor, pose history), so a step is
  a chain of kernels with no device->host round trip: ``poses_device`` (a4) ->
  ``advance`` (instance append + next rock) -> ``observe_walls`` / ``observe_rocks``
  (a2/a3).  ``walls`` [E,H,W] and ``rocks`` [E,R,h,h] are float32 CUDA tensors in the
  reference's float32 elevation arithmetic."""

  def __init__(self, bank, envs, capacity, overhead_resolution=128, object_resolution=32,
               pixel_size=0.125 / 32, max_z=0.375, orientation_freedom=0,
               spawn_pose=None, device=None, episode_length=None,
               rock_cache_bytes=1 << 30):
    """``rock_cache_bytes``: the image of a spawned rock depends only on its mesh (fixed
    spawn pose and orientation list), so when the images of the whole bank fit this many
    bytes they are rasterised once and a step fetches them by mesh id (0: rasterise the
    spawned rocks every step)."""
    self.geo = g = camera.ObserverGeometry(
      overhead_resolution, object_resolution, pixel_size, max_z, orientation_freedom)
    self.bank = bank
    self.E, self.cap, self.R = int(envs), int(capacity), g.n_orientations
    self.dev = device if device is not None else _device()
    self.spawn_pose = spawn_pose if spawn_pose is not None else \
      ((0., 0., max_z + g.object_z), (0., 0., 0., 1.))
    self._verts, self._tris = bank.device(self.dev)
    self._ranges = np.ascontiguousarray(np.asarray(bank.ranges, dtype='int32').reshape(-1, 4))
    self._coms = np.ascontiguousarray(np.asarray(bank.coms, dtype='float64').reshape(-1, 3))
    E, cap, R = self.E, self.cap, self.R
    isz, jsz = capi.INSTANCE_DTYPE.itemsize, capi.JOB_DTYPE.itemsize
    # Instance row of every mesh of the bank at the spawn pose: observing a new
    # rock is a device-side row copy by mesh id, no per-step host work.
    n = len(bank)
    spawn_rows = self._rows(np.arange(n), [self.spawn_pose[0]] * n, [self.spawn_pose[1]] * n)
    self._spawn_rows = self._upload_rows(spawn_rows)
    self.state = capi.EnvState(
      E, cap, int(episode_length if episode_length is not None else cap),
      torch.from_numpy(self._ranges).to(self.dev), torch.from_numpy(self._coms).to(self.dev),
      self._spawn_rows, self.dev)
    self._inst = self.state.instances
    self._rock_inst = self.state.rock_instances
    self.counts = self.state.counts
    jobs = np.zeros(E, dtype=capi.JOB_DTYPE)
    jobs['view'], jobs['proj'] = g.overhead_view, g.overhead_projection
    jobs['inst_begin'] = np.arange(E) * cap
    jobs['zrange'] = g.overhead_z
    self._wall_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(-1)).to(self.dev)
    jobs = np.zeros((E, R), dtype=capi.JOB_DTYPE)
    for k in range(R):
      jobs[:, k]['view'] = g.object_view(self.spawn_pose, k)
    jobs['proj'] = g.object_projection
    jobs['inst_begin'] = np.arange(E)[:, None]
    jobs['inst_count'] = 1
    jobs['zrange'] = g.object_z
    self._rock_jobs = torch.from_numpy(jobs.view(np.uint8).reshape(-1)).to(self.dev)
    # every orientation's camera looks at the same instance: the per-job instance
    # count of the wall images is the environment's `counts` entry, read on the device
    self._max_verts = int(self._ranges[:, 1].max()) if n else 0
    self._orientations = torch.from_numpy(
      np.ascontiguousarray(np.asarray(g.orientations, dtype='float64').reshape(R, 4))).to(self.dev)
    self._pose_geometry = (g.pixel_h, g.pixel_w, g.object_x, g.object_y, g.object_z)
    self.pose_buf = torch.zeros((E, 7), dtype=torch.float64, device=self.dev)
    self.status = torch.zeros((E,), dtype=torch.int32, device=self.dev)
    self._status_zero = torch.zeros((E,), dtype=torch.int32, device=self.dev)
    self.walls = torch.zeros((E, g.overhead_h, g.overhead_w), dtype=torch.float32,
                             device=self.dev)
    self.rocks = torch.zeros((E, R, g.object_h, g.object_w), dtype=torch.float32,
                             device=self.dev)
    # GL depth image of the placed rocks of every environment, kept between steps:
    # a step that only appends a rock draws that rock alone onto it (same bits as
    # re-drawing the scene: the depth image is a minimum over triangles).
    self._wall_depth = torch.ones((E, g.overhead_h, g.overhead_w), dtype=torch.float32,
                                  device=self.dev)
    self._wall_depth_valid = False
    # image rows the last observe_walls may have changed, per environment (for consumers
    # that keep a derived image, e.g. the packed observation, up to date row by row)
    self.wall_rows = torch.zeros((E, 2), dtype=torch.int32, device=self.dev)
    self._rock_cache = None
    self._cache_rocks = 0 < n * R * g.object_h * g.object_w * 4 <= int(rock_cache_bytes)

  # -- instance rows -------------------------------------------------------------- #
  def _rows(self, mesh_ids, positions, quaternions):
    """INSTANCE_DTYPE rows for rocks whose INERTIAL frames sit at the given
    poses (the visual mesh is offset by -com, like a URDF base)."""
    mesh_ids = np.asarray(mesh_ids, dtype='int64')
    q = np.asarray(quaternions, dtype='float64').reshape(-1, 4)
    p = np.asarray(positions, dtype='float64').reshape(-1, 3)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    s = 2.0 / (x * x + y * y + z * z + w * w)
    rot = np.stack([1.0 - s * (y * y + z * z), s * (x * y - z * w), s * (x * z + y * w),
                    s * (x * y + z * w), 1.0 - s * (x * x + z * z), s * (y * z - x * w),
                    s * (x * z - y * w), s * (y * z + x * w), 1.0 - s * (x * x + y * y)], -1)
    rows = np.zeros(len(mesh_ids), dtype=capi.INSTANCE_DTYPE)
    rows['rot'] = rot
    com = self._coms[mesh_ids]
    r3 = rot.reshape(-1, 3, 3)
    # same left-to-right sum as the device kernel (srl_env_advance)
    rows['pos'] = p - ((r3[:, :, 0] * com[:, None, 0] + r3[:, :, 1] * com[:, None, 1]) +
                       r3[:, :, 2] * com[:, None, 2])
    rng = self._ranges[mesh_ids]
    rows['vert_begin'], rows['vert_count'] = rng[:, 0], rng[:, 1]
    rows['tri_begin'], rows['tri_count'] = rng[:, 2], rng[:, 3]
    return rows

  def _upload_rows(self, rows):
    return torch.from_numpy(rows.view(np.uint8).reshape(len(rows), -1)).to(
      self.dev, non_blocking=True).view(torch.float64)

  # -- episodes --------------------------------------------------------------------- #
  def begin(self, orders, env_ids=None):
    """Start episodes (env.py:266-293): ``orders`` is the full [E, length] host table
    of mesh ids in pop order (uploaded whole; None: ``state.order`` was filled on the
    device, srl_env_draw); ``env_ids``: the environments to restart (None: all)."""
    st = self.state
    if orders is not None:
      st.order.copy_(torch.from_numpy(np.ascontiguousarray(orders, dtype='int32')),
                     non_blocking=True)
    ids = None if env_ids is None else torch.from_numpy(
      np.ascontiguousarray(env_ids, dtype='int32')).to(self.dev, non_blocking=True)
    capi.env_reset(st, ids)
    self._wall_depth_valid = False
    if env_ids is None:
      self.status.copy_(self._status_zero, non_blocking=True)     # device memcpy

  def reset(self, env_ids=None):
    """Forget the placed rocks of the given environments (all by default)."""
    if env_ids is None:
      self.counts.zero_()
    else:
      self.counts[torch.as_tensor(env_ids, device=self.dev, dtype=torch.long)] = 0

  def place(self, mesh_ids, positions, quaternions, env_ids=None):
    """Append one placed rock per environment from HOST poses (set-up paths and
    tests; a step uses ``advance``)."""
    env_ids = torch.arange(self.E, device=self.dev) if env_ids is None else \
      torch.as_tensor(env_ids, device=self.dev, dtype=torch.long)
    rows = self._upload_rows(self._rows(mesh_ids, positions, quaternions))
    slot = env_ids * self.cap + self.counts[env_ids].long()
    inst_rows = self._inst.view(torch.float64).view(self.E * self.cap, -1)
    inst_rows.index_copy_(0, torch.clamp(slot, max=self.E * self.cap - 1), rows)
    self.counts[env_ids] = torch.clamp(self.counts[env_ids] + 1, max=self.cap)
    self._wall_depth_valid = False

  def poses_device(self, views, flat):
    """Observer.pose for every environment on the device (observer.py:392-421):
    ``views`` [E] int64 or None, ``flat`` [E] int64 CUDA tensors (any stride).
    Fills ``pose_buf`` [E,7] float64 = (x, y, z, qx, qy, qz, qw) and ``status``."""
    capi.place_poses(self.walls, self.rocks, views, flat, self._orientations,
                     self._pose_geometry, threshold=10 ** (-4), poses=self.pose_buf,
                     status=self.status)
    return self.pose_buf

  def advance(self, rest=None, placed=None):
    """The spawned rocks come to rest at ``rest`` [E,7] (default: where they were
    placed, ``pose_buf``); next rock of every episode spawned (env.py:245-249)."""
    capi.env_advance(self.state, self.pose_buf if rest is None else rest, placed)

  def set_poses(self, poses):
    """Rewrite the rest poses of the first n placed rocks of every environment:
    ``poses`` [E,n,7] float64 (a physics step that disturbed earlier rocks)."""
    poses = torch.as_tensor(np.ascontiguousarray(poses, dtype='float64')) \
      if not isinstance(poses, torch.Tensor) else poses
    capi.env_set_poses(self.state, poses.to(self.dev).contiguous())
    self._wall_depth_valid = False

  # -- capture -------------------------------------------------------------------- #
  def observe_walls(self, appended=False):
    """Rasterise every environment's placed rocks into ``walls``
    (observer.py:252-260).  ``appended``: since the last call exactly one rock was
    appended to every environment (``advance``) and nothing else moved -- only that
    rock is drawn, onto the kept depth image; otherwise the whole scene."""
    g = self.geo
    last = bool(appended) and self._wall_depth_valid
    capi.raster(self._verts, self._tris, self._inst, self._wall_jobs, g.overhead_h,
                g.overhead_w, capi.RASTER_WALL, far_plane=FAR, out=self.walls,
                inst_counts=self.counts, depth_state=self._wall_depth,
                only_last=2 if last else 0,       # walls still holds the kept image
                rows_out=self.wall_rows,
                # (the true mesh size: up to 128 vertices the appended rock is drawn by the
                # warp-per-image kernel)
                max_cached_verts=self._max_verts if last else
                min(2048, max(256, self._max_verts * self.cap)))
    self._wall_depth_valid = True
    return self.walls

  def observe_rocks(self, mesh_ids=None):
    """Rasterise the undersides of the spawned rocks at every orientation into
    ``rocks`` (observer.py:262-293).  ``mesh_ids`` (host array) replaces the
    device-side episode state with explicit meshes (set-up paths)."""
    g = self.geo
    if mesh_ids is None and self._cache_rocks:
      if self._rock_cache is None:
        self._rock_cache = self._render_bank()
      capi.gather_rows(self._rock_cache, self.state.current, self.rocks.view(self.E, -1))
      return self.rocks
    if mesh_ids is not None:
      ids = torch.as_tensor(np.asarray(mesh_ids, dtype='int64')).to(self.dev, non_blocking=True)
      torch.index_select(self._spawn_rows, 0, ids,
                         out=self._rock_inst.view(torch.float64).view(self.E, -1))
    capi.raster(self._verts, self._tris, self._rock_inst, self._rock_jobs, g.object_h,
                g.object_w, capi.RASTER_ROCK, far_plane=FAR,
                out=self.rocks.view(self.E * self.R, g.object_h, g.object_w),
                max_cached_verts=max(256, self._max_verts))
    return self.rocks

  def _render_bank(self):
    """[n_meshes, R*h*w] float32: every mesh of the bank at the spawn pose, seen by the
    R orientation cameras -- the same instance rows, jobs and kernel as the per-step
    rasterisation, so the fetched images have its bits."""
    g, n, R = self.geo, len(self.bank), self.R
    jobs = np.zeros((n, R), dtype=capi.JOB_DTYPE)
    for k in range(R):
      jobs[:, k]['view'] = g.object_view(self.spawn_pose, k)
    jobs['proj'] = g.object_projection
    jobs['inst_begin'] = np.arange(n)[:, None]
    jobs['inst_count'] = 1
    jobs['zrange'] = g.object_z
    jobs_d = torch.from_numpy(jobs.view(np.uint8).reshape(-1)).to(self.dev)
    cache = torch.empty((n, R * g.object_h * g.object_w), dtype=torch.float32, device=self.dev)
    insts = self._spawn_rows.view(torch.uint8).view(-1)
    capi.raster(self._verts, self._tris, insts, jobs_d, g.object_h, g.object_w,
                capi.RASTER_ROCK, far_plane=FAR, out=cache.view(n * R, g.object_h, g.object_w),
                max_cached_verts=max(256, self._max_verts))
    return cache

  def poses(self, views, flat_actions):
    """Observer.pose for every environment, on the host (observer.py:392-421):
    ``views`` [E] orientation index, ``flat_actions`` [E] row-major position.
    Returns (positions [E,3] float64 numpy, quaternions [E,4]); raises on an
    action outside the action space (env.py:237)."""
    as_dev = lambda a: None if a is None else torch.as_tensor(
      np.asarray(a.cpu() if isinstance(a, torch.Tensor) else a, dtype='int64')).to(self.dev)
    poses = self.poses_device(as_dev(views), as_dev(flat_actions)).cpu().numpy()
    if bool(self.status.any()):
      raise AssertionError('Invalid action.')
    return poses[:, :3].copy(), poses[:, 3:].copy()
