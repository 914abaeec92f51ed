"""Stub (TEST INFRASTRUCTURE ONLY) for the decorator used at reference trainer.py:164-173."""


def if_exception_type(*a, **k):
    return lambda e: False


class Retry:
    def __init__(self, *a, **k):
        pass

    def __call__(self, fn):
        return fn
