"""Drivers mirroring python/verify of the reference (test_backends, purity) on the B200 backend."""
