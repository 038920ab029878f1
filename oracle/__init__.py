"""TEST INFRASTRUCTURE (CPU oracle). Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this."""
