"""Test-only CPU oracle (see score_oracle.py header)."""
