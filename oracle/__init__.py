"""Oracle package: CPU restatements of the reference algorithm.  TEST INFRASTRUCTURE ONLY
(see oracle/fem_oracle.py header).  Never imported by the product package."""
