"""CPU oracle for the RQAE residual-quantization hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the shipped
product: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import or execute it, and there
only as the checker or as the CPU arm being timed, never as the GPU path.

Parity status: the reference (harish-kamath/rqae) ships no tests, golden vectors
or fixtures, so the oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF: ``tests/golden/make_golden.py`` imports the unmodified
``rqae.model.RQAE`` from /root/reference, runs it on seeded inputs and commits
the results under ``tests/golden/``; ``tests/test_oracle.py`` checks both oracle
implementations against those files.
"""
