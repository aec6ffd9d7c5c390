"""Oracle package -- test infrastructure only (see ullava_oracle.py header)."""
