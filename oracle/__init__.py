"""CPU oracle (test infrastructure only — never imported by spark_rapids_ml_b200)."""
