"""TEST INFRASTRUCTURE — the CPU oracle.  Never import this from lc_b200/ (the product)."""
