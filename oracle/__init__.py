"""CPU oracle for the stackrl observation + placement-scoring path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``stackrl_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` may use it, and only as the checker
or the timed CPU baseline -- never as a product code path.

Contents
  refload.py       loads the UNMODIFIED reference modules from /root/reference
                   behind stub ``gin``/``gym``/``tensorflow``/``pybullet``
                   modules (only works where /root/reference exists; used to
                   pin the restatement and to generate tests/golden/*.npz).
  scoring_np.py    numpy restatement of stackrl/baselines.py (live part),
                   agents/policies.py:PyGreedy and Observer.pose.
  observe_np.py    numpy restatement of the depth->elevation conversion
                   (observer.py:259-260, 274-277), StackEnv.observation /
                   _return (env.py:171-180, 226-231) and the IoU / OR rewards
                   (rewarder.py:162-179, 297-307).
  raster_np.py     software z-buffer standing in for pybullet's TinyRenderer
                   (NOT in the reference tree: raster parity is UNPINNED).
  nets_np.py       numpy restatement of the Siamese correlation layer
                   (nets/layers.py:21-38; PARITY UNPINNED against TensorFlow).
  fake_pybullet.py duck-typed pybullet camera API + static bodies around it.
  csrc/            plain-C restatement of the same arithmetic for sizes the
                   numpy versions are too slow for (built to oracle/_build/).

Parity pinning status
  scoring / conversion / rewards / packing: PINNED against the reference's own
  code run in the build container (tests/golden/make_golden.py ->
  tests/golden/*.npz, checked by tests/test_oracle_golden.py).
  rasterisation: PARITY UNPINNED -- the reference delegates it to pybullet
  (unvendored, unpinned, absent here); see DESIGN.md.
"""
