"""Oracle package: TEST INFRASTRUCTURE ONLY (see oracle/glfer_oracle.py header)."""
