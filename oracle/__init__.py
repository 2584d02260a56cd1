"""CPU oracle of the normflow hot path -- test infrastructure only (see nf_oracle.py)."""
