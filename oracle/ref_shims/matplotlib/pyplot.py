"""Stub (see package docstring)."""
