"""CPU oracle of the gDCA hot path -- TEST INFRASTRUCTURE ONLY (see gdca_oracle.py).
Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may import this package."""
