"""CPU oracle for the EFGHNet lattice/BCL hot path - TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (efgh_b200/) never does: it fails loudly when its CUDA library is missing.
"""
