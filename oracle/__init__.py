"""CPU oracle (TEST INFRASTRUCTURE ONLY - see ncf_oracle.py).  Never imported by the product package."""
