"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT: CPU restatements of the reference `advance_mu_t`.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may import this."""
