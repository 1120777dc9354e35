"""CPU oracle (test infrastructure only; see oracle/fsp_oracle.c header)."""
