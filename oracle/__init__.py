"""CPU oracle for the Project-NeRF ray-marching hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline -- never as a fallback for the CUDA path.

Parity status
-------------
* tcnn-free functions (Fourier encoding, NeRFDecoder, DeformationNetwork,
  TimeModulationNetwork, stratified sampling, occupancy mask, compositing,
  ``render_rays`` wiring, ``DensityGrid.update``): PINNED against the
  reference's own ``src/`` imported from ``/root/reference`` -- see
  ``tests/golden/make_golden.py`` (generator, run in the build container) and
  ``tests/test_oracle_golden.py`` (checker, runs anywhere).
* tiny-cuda-nn arithmetic (HashGrid encoding, FullyFusedMLP): PARITY UNPINNED.
  ``tinycudann`` is an un-vendored, un-pinned dependency of the reference
  (absent from requirements.txt, imported at src/embeddings.py:57,
  src/decoders.py:107,281) and the reference holds no tests or golden vectors
  at that boundary.  ``nerf_oracle.hash_encode`` / ``fused_mlp`` restate the
  upstream library's published algorithm;
  the reference's own wiring around them (src/embeddings.py:75-89,
  src/decoders.py:136-162,300-318, src/core.py) is pinned by running the
  reference files verbatim on top of ``tcnn_shim``.
"""
