"""TEST INFRASTRUCTURE: CPU oracles for the fm3d hot path (never imported by the product)."""
