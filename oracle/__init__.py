"""CPU oracle for the emulator forward pass.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (torch-CPU convs in fp64/fp32, numpy integer
index math, scipy quadrature for the cosmology scalars) of the algorithm in
``/root/reference/src/jax_nbody_emulator``.  It exists to check the CUDA path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package
(``jax_nbody_emulator_with_dj_b200``) never imports, links or executes anything
from here; it fails loudly when its CUDA library is missing.

PINNING (round 2).  The reference is pure JAX/Flax; jax, jaxlib and flax are not
installed in this image (nor in /opt/wheelhouse) and the pretrained weight blob is
absent (``/root/reference/.MISSING_LARGE_BLOBS``), so the reference cannot run on
the JAX runtime here and its own tests contain no golden output of the net.
Instead ``oracle/jaxshim`` provides stand-in ``jax`` / ``flax`` packages (numpy and
torch-CPU underneath) for exactly the API subset the reference calls, and
``tools/make_reference_golden.py`` executes the UNMODIFIED reference package over
them: its four models, both premodulation functions, its cosmology module and
``SubboxProcessor.process_box``.  The outputs are committed as
``tests/golden/ref_*.npz``; ``tests/test_oracle_vs_reference.py`` checks this oracle
against them (layers / blocks 1e-13, whole network and the committed oracle
fixtures 1e-12 in fp64) and ``tests/test_gpu_reference.py`` checks the CUDA path
against them directly.  That pins the reference's ALGORITHM — every formula, index,
crop and loop of its source — but not XLA's last-bit floating point: the
convolution primitive, the fp32 hyp2f1 series and the summation order underneath
are the stand-in's (``oracle/jaxshim/README.md``).  With the fixed-seed parameter
tree only: the pretrained weights do not exist in either checkout.

What *is* pinned against the reference's tests (see tests/test_oracle_*.py):
integer tiling tables and their invariants (tests/test_subbox.py:86-204),
LeakyReLU known answers incl. the x==0 tangent branch (tests/test_layers_vel.py:
268-334), the shape law out = in - 96, parameter-tree names/shapes, the algebraic
invariants of the velocity branch, and the cosmology values quoted in
README.md:178-180.
"""
