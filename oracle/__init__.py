"""CPU oracle for the FDFD operator hot path (TEST INFRASTRUCTURE ONLY).

This package is a CPU restatement (numpy/scipy + a small C SpMV) of the algorithm the
reference package MaxwellFDFD.jl uses for the hot path named in BASELINE.json:
assembling and applying  A = Cm * (Pmu \\ Ce) - w^2 * Peps  (reference
src/model/model.jl:225-246) from the difference/averaging operators of its un-vendored
dependency MaxwellBase ^0.1.6 -> StaggeredGridCalculus (Project.toml:10,15), whose
published algorithm is restated from SURVEY.md Appendix A.

PARITY UNPINNED at the operator boundary: the reference holds no golden vector or test
for create_curls/create_paramops/create_A (test/source.jl covers sources only), there is
no Julia in this image, and MaxwellBase is not on disk.  The operator restatement is
therefore pinned by mathematical property tests (tests/test_oracle_properties.py), and
the source restatement (oracle/source.py) IS pinned by the reference's own known-answer
table (test/source.jl:5-230, restated in tests/test_oracle_source.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package - as the checker or the timed CPU baseline, never as
something the product path routes through.
"""
