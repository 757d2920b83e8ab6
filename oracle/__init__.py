"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the LearnMultigrid V-cycle hot path used only as the
checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
Nothing under learnmultigrid_b200/ imports this package.

Parity pinning: the restatement is pinned against the reference's own code executed in the
authoring container through oracle/refshim.py (real learn_multigrid modules + two shims), with
the resulting known answers committed under tests/golden/ by tests/golden/make_golden.py.
"""
