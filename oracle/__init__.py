"""ORACLE — test infrastructure only.

CPU restatements of the reference's hot path (Swin-UNet forward + residual-shift sampling
loop) used as the parity checker.  Nothing under ``oracle/`` is part of the product path:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs import it.
"""
