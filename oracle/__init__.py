"""CPU oracle of the lshrs hot path -- test infrastructure only (see lshrs_oracle.py)."""
