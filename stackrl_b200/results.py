"""On-disk result formats of the reference's evaluation harness.

``stackrl.test`` stores what a run produced in two files that its analysis and
heat-map tools read back (stackrl/test.py:46-148, 812-826; stackrl/heatmap.py):

* ``data.npz``   -- ``np.savez_compressed`` of ``keys`` (policy names), ``actions``
  [policies, steps, 2] (row, column; uint8, or uint16 for maps of 256+ positions a
  side), ``values`` [policies, steps, positions] float32 (every policy scores every
  observation, test.py:221-224, 269-280), ``rewards`` [policies, steps per policy]
  float32 and ``episode_bounds`` (uint16 / uint32, strictly increasing);
* ``results.csv`` -- one line per policy; columns are the keyword names in
  CamelCase; a file whose header matches is appended to, lines with a repeated
  ``Keys`` entry are replaced unless the old line has a higher ``Priority``.

These writers produce the same files from GPU-side results (value maps of
``PlacementScorer(want_shown=True)``, actions, rewards), so the reference's
``analyse`` / ``heatmap`` run on them unchanged.  Host code, numpy only.
"""
import os

import numpy as np


def _camel(name):
  return ''.join(part[:1].upper() + part[1:] for part in name.split('_'))


def write(fname, force=False, **kwargs):
  """``stackrl.test.write`` (test.py:46-148): append ``kwargs`` (column name ->
  iterable, scalars broadcast) to the CSV file ``fname``."""
  size = None
  for v in kwargs.values():
    if not np.isscalar(v):
      size = len(v)
      break
  if size is None:
    raise ValueError('at least one column must be an iterable')
  cols = {_camel(k): (np.array([v] * size) if np.isscalar(v) else np.array(v))
          for k, v in kwargs.items()}

  def lines_of(columns, skip=()):
    for i, row in enumerate(zip(*columns.values())):
      if i not in skip:
        yield ','.join(str(x) for x in row) + '\n'

  if os.path.isfile(fname):
    with open(fname) as f:
      head_line = f.readline()
      old = f.readlines()
    header = head_line[:-1].split(',')
    if set(header) == set(cols):
      keep, skip, rewrite = [head_line], [], False
      if 'Keys' in header:
        ik = header.index('Keys')
        ip = header.index('Priority') if 'Priority' in header else None
        for line in old:
          fields = line[:-1].split(',')
          if fields[ik] in cols['Keys']:
            if ip is not None:
              i = int(np.where(cols['Keys'] == fields[ik])[0][0])
              if float(fields[ip]) > cols['Priority'][i]:
                keep.append(line)          # the old line outranks the new one
                skip.append(i)
              else:
                rewrite = True
            else:
              rewrite = True
          else:
            keep.append(line)
      if rewrite:
        with open(fname, 'w') as f:
          f.writelines(keep)
      with open(fname, 'a') as f:
        f.writelines(lines_of({k: cols[k] for k in header}, skip))
      return
    if not force:
      raise ValueError("kwargs don't match the existing file's header.")
  dirname = os.path.dirname(fname)
  if dirname and not os.path.isdir(dirname):
    os.makedirs(dirname)
  with open(fname, 'w') as f:
    f.write(','.join(cols.keys()) + '\n')
    f.writelines(lines_of(cols))


def pack_data(keys, actions, values, rewards, episode_bounds, map_shape):
  """The dictionary ``stackrl.test.run`` returns (test.py:150-345), with its dtypes.

  ``actions`` [policies, steps] flat indices (or [policies, steps, 2] already
  unravelled), ``values`` [policies, steps, positions] (any float dtype; the
  reference stores float32), ``rewards`` [policies, steps per policy],
  ``episode_bounds`` any iterable of step indices (made unique and sorted, with the
  total step count appended like test.py:333-341)."""
  keys = np.array([str(k) for k in keys])
  actions = np.asarray(actions)
  if actions.ndim == 2:
    actions = np.stack(np.unravel_index(actions, map_shape), axis=-1)
  total = actions.shape[1]
  bounds = np.unique(np.array(list(episode_bounds) + [total],
                              dtype='uint16' if total < 2 ** 16 else 'uint32'))
  return {
    'keys': keys,
    'actions': actions.astype('uint8' if max(map_shape) < 2 ** 8 else 'uint16'),
    'values': np.asarray(values, dtype='float32').reshape(len(keys), total, -1),
    'rewards': np.asarray(rewards, dtype='float32'),
    'episode_bounds': bounds,
  }


def save_data(dirname, **data):
  """``np.savez_compressed(os.path.join(dirname, 'data'), **data)`` (test.py:812-815)."""
  if not os.path.isdir(dirname):
    os.makedirs(dirname)
  path = os.path.join(dirname, 'data')
  np.savez_compressed(path, **data)
  return path + '.npz'
