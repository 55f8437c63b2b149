"""Benchmark models of the hot path (MA2, Gaussian noise, g-and-k) on the elfi_b200 node API."""
