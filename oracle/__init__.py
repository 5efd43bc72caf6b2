"""CPU oracle (test infrastructure only; see sgmse_oracle.py header).  Never imported by the product path."""
