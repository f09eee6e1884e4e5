"""Full-size parity + timing scripts (they use the oracle as the checker, hence they live under tests/; not collected by pytest)."""
