"""``partialize``: the decorator python/data_gen wraps its entry points with
(channelize.py:19, synthesize.py:27, generate_test_vector.py:74).  Calling the
decorated function with keyword arguments only returns a partially-applied
callable (``channelize(backend="python")``, test_backends.py:80, purity.py:112);
any positional argument triggers the call."""
from __future__ import annotations

import functools


def partialize(func):
    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        if len(args) == 0 and kwargs:
            inner = functools.partial(wrapper_call, **kwargs)
            functools.update_wrapper(inner, func)
            return inner
        return func(*args, **kwargs)

    def wrapper_call(*args, **kwargs):
        return func(*args, **kwargs)

    wrapper.__wrapped__ = func
    return wrapper
