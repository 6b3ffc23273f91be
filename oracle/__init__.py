"""CPU oracle (test infrastructure): never imported by the product package."""
