"""CPU oracle for the CLIP-guidance hot path — TEST INFRASTRUCTURE, not product code.

A float32 PyTorch-on-CPU restatement of the reference algorithm (perceptor v0.6.7), each function citing the
reference file:line it follows.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference legs
may import this package; perceptor_b200/ never does (tests/test_boundary.py enforces it).

Pinning: the reference's own tests hold no numeric vectors for this path, so the oracle is
pinned against outputs of the reference's in-tree files run in the build container: oracle/make_golden.py imports
perceptor/transforms/resize/resize_right.py and perceptor/models/ruclip/model.py BY PATH from /root/reference and
writes tests/golden/*.npz; tests/test_oracle_golden.py checks every oracle function against those vectors.
The cutout sampler (S0) has no referent in the reference: its parity is between this restatement and the product
spec only ("cutout-index parity vs the reference: unpinned", as DESIGN.md states).
"""
