"""Test infrastructure only: CPU float64 oracle, reference build recipe and loader. Never imported by the product."""
