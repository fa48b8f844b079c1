"""Drop-in ``src`` package: the reference's module API (src/abstract.py, core.py,
embeddings.py, decoders.py, renderer.py) on top of the sm_100a kernels in
libb2nerf.so.  Put ``project-nerf_b200/`` ahead of the reference on sys.path
and the reference's unchanged run.py drives this implementation."""
