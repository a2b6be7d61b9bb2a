"""Stub for ``more_itertools`` (TEST INFRASTRUCTURE): the reference imports
``intersperse`` (dctn/eps_plus_linear.py:7, unused), ``chunked`` (dctn/conv_sbs.py:9) and
``last`` (dctn/training.py:11)."""
from itertools import islice


def intersperse(e, iterable, n=1):
    it = iter(iterable)
    first = True
    while True:
        chunk = list(islice(it, n))
        if not chunk:
            return
        if not first:
            yield e
        first = False
        yield from chunk


def chunked(iterable, n):
    it = iter(iterable)
    while True:
        chunk = list(islice(it, n))
        if not chunk:
            return
        yield chunk


def last(iterable, *default):
    item = None
    found = False
    for item in iterable:
        found = True
    if not found:
        if default:
            return default[0]
        raise ValueError("last() was called on an empty iterable")
    return item
