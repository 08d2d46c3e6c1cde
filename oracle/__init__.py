"""CPU oracle for the ChronoClust hot path -- TEST INFRASTRUCTURE ONLY (see chronoclust_oracle.c)."""
