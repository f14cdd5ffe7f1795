"""CPU oracle + reference-frontend harness.  TEST INFRASTRUCTURE ONLY (see tfhe_ref.c header)."""
