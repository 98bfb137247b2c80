"""bench-only baselines (never imported by the product)"""
