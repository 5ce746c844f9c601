"""CPU oracle for the kzg_rust blob path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this package.  The product package (kzg_rust_b200) never does.
"""
