"""Temporary-file disposal context manager (python/data_gen/dispose.py:13-85)."""
from __future__ import annotations

import contextlib
import os

__all__ = ["dispose"]


@contextlib.contextmanager
def dispose(*callbacks, dispose=True):
    files = []
    for res in callbacks:
        if callable(res):
            res = res()
        files.append(res)
    yield files[0] if len(files) == 1 else files

    def remove_file(f):
        if hasattr(f, "file_path"):
            f = f.file_path
        if isinstance(f, str) and os.path.exists(f):
            os.remove(f)

    if dispose:
        for f in files:
            if isinstance(f, (tuple, list)):
                for g in f:
                    remove_file(g)
            else:
                remove_file(f)
