"""CPU oracle (test infrastructure only -- see grafp_oracle.py header)."""
