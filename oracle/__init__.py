"""CPU oracle for the DeViT hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this package, and only as the checker / the timed CPU baseline.
Nothing under ``devit_b200/`` imports it; the product path fails loudly without its CUDA
library instead of falling back to this code.

Parity pin: ``tests/golden/*.npz`` were produced by the ``tests/golden/make_*golden.py`` scripts
running the UNMODIFIED reference modules from /root/reference (through ``oracle/ref_shim.py``) on
the seeded synthetic weights/inputs; ``tests/test_oracle_golden.py``, ``test_cct_oracle_golden.py``,
``test_edge_oracle_golden.py`` and ``test_hsic_cpu.py`` check the restatements against them.
"""
