"""CPU oracle package (test infrastructure only; see oracle/ac_oracle.h)."""
