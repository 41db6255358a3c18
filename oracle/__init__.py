"""CPU oracle (test infrastructure only) – see pfb_oracle.py."""
