"""CPU oracle for the SlotDiffusion hot path -- TEST INFRASTRUCTURE ONLY.

A from-scratch restatement (torch CPU tensor ops, fp32 or fp64) of the reference
algorithm, each function citing the reference file:line it follows.  The oracle
is pinned against golden vectors produced by the reference modules themselves
(imported in the build container, see tools/make_golden.py -> tests/golden/).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (slotdiffusion_b200/) never
imports it and has no CPU fallback.
"""
