"""CPU oracle (test infrastructure only — see scarplet_oracle.py header)."""
