"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatements of the reference's frame-synthesis hot path, used as the parity
checker.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this package.  The product package (fvfi) never
does, and fails loudly when its CUDA library is missing.
"""
