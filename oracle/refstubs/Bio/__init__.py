"""Stub of the tiny Biopython surface the reference's hot path touches (test infra only)."""
