"""CPU oracle (test infrastructure only - see moe_oracle.py header)."""
