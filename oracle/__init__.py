"""CPU oracle for the emulator forward pass.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (torch-CPU convs in fp64/fp32, numpy integer
index math, scipy quadrature for the cosmology scalars) of the algorithm in
``/root/reference/src/jax_nbody_emulator``.  It exists to check the CUDA path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package
(``jax_nbody_emulator_with_dj_b200``) never imports, links or executes anything
from here; it fails loudly when its CUDA library is missing.

PARITY UNPINNED (network level): the reference is pure JAX/Flax; jax, jaxlib and
flax are not installed in this image (nor in /opt/wheelhouse) and the pretrained
weight blob is absent (``/root/reference/.MISSING_LARGE_BLOBS``), so the reference
itself cannot be run here and its own tests contain no golden output of the net.
What *is* pinned against the reference's tests (see tests/test_oracle_*.py):
integer tiling tables and their invariants (tests/test_subbox.py:86-204),
LeakyReLU known answers incl. the x==0 tangent branch (tests/test_layers_vel.py:
268-334), the shape law out = in - 96, parameter-tree names/shapes, the algebraic
invariants of the velocity branch, and the cosmology values quoted in
README.md:178-180.
"""
