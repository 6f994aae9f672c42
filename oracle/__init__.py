"""CPU oracle for the galax hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in numpy / plain C, the algorithm of the reference
(GalacticDynamics/galax + the diffrax 0.7.0 solvers it calls) for the one hot
path this repository accelerates.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.  Nothing under ``galax_b200/`` imports it, and the
product path fails loudly when the CUDA library is missing.

Parity status:

* potentials / gradients / Hessians / densities: PINNED against every known-answer
  test the reference holds for MiyamotoNagai, Hernquist, NFW, PowerLawCutoff,
  MN3Exponential, MN3Sech2, MilkyWayPotential, MilkyWayPotential2022,
  BovyMWPotential2014 (``tests/golden/potential_kats.json``, transcribed from
  ``/root/reference/tests/unit/potential/builtin/*.py``).
* integrators (SemiImplicitEuler, Dopri8+PID) and the mock-stream path:
  **parity unpinned** at the 1e-12 / 10*tol level -- the arithmetic lives in
  diffrax 0.7.0 / jax 0.8.0, which are not vendored in ``/root/reference`` and are
  not importable in the authoring container.  The restatement is anchored on the
  reference's call sites and its 4-8 digit orbit doctests (``tests/golden/orbit_kats.json``).
"""
