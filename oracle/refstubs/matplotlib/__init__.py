"""Stub (test infra): pyani_config imports matplotlib.colors at import time only."""
