"""Stub: the reference imports matplotlib at module import time (data.py:6); nothing on the
hot path uses it.  TEST INFRASTRUCTURE ONLY."""
def use(*a, **k):
    pass
