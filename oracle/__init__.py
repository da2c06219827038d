"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's flow-warp + occlusion-blend path
(/root/reference/src/utils/ops.py:183-202, src/modules/generator/generator.py:80-96,
src/utils/utils.py:346-354).  Nothing in the product package ``c2m_b200`` imports this
package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs do, and there only as the checker or as the timed CPU baseline.

Pinning status
--------------
* ``reference_torch`` is pinned against golden vectors produced by importing the *unmodified*
  reference sources in the build container (``oracle/make_golden.py`` ->
  ``tests/golden/*.npz``).
* ``warp_numpy`` (which also restates ATen's ``grid_sampler_2d`` forward/backward, a third-party
  dependency that is not vendored under /root/reference -- torch 2.11.0, header
  ``ATen/native/cuda/GridSampler.cuh``) is pinned against the same golden vectors.
The reference ships no tests of its own, so these goldens are the pin.
"""
