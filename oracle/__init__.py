"""CPU oracle + reference-compiled libraries: TEST INFRASTRUCTURE ONLY (see oracle/gptq_oracle.c)."""
