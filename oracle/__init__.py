"""oracle/ — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU (PyTorch, eager) restatement of the reference's algorithm for BAGEL's inference forward path, written
as pure functions over a state dict with the reference's parameter names. Every function cites the
reference file:line it follows (paths relative to ByteDance-Seed/Bagel @ 056b5fd).

Pinning: the reference ships no tests or golden vectors, so the oracle is pinned against
outputs of the reference itself: tests/golden/make_golden.py imports the unmodified reference in this
container (oracle/ref_shims.py) and commits input/output fixtures under tests/golden/; tests/test_oracle_*.py
check the restatement against those fixtures bit-for-bit on CPU.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package, and only as the checker. bagel_b200/ never imports it.
"""
