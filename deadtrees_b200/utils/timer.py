"""``record_execution_time`` as ``deadtrees/utils/timer.py:1-8`` (wall-clock of a ``with`` block)."""
from contextlib import contextmanager
from time import perf_counter


@contextmanager
def record_execution_time():
    start = perf_counter()
    yield lambda: perf_counter() - start
