"""Minimal progress-bar shim with the call surface the hot path uses from the reference's
`utils/pbar.py`: `pbar(iterable, desc=...)`, `pbar.quiet()`, `pbar.print(...)`,
`pbar.descnext(...)`.  The module object itself is callable, like the reference's.
"""
import sys
import types
from contextlib import contextmanager

try:
    from tqdm import tqdm as _tqdm
except Exception:  # pragma: no cover
    _tqdm = None

_state = {'quiet': 0, 'desc': None}


def _iterate(iterable, desc=None, **kwargs):
    if desc is None:
        desc, _state['desc'] = _state['desc'], None
    if _state['quiet'] or _tqdm is None or not sys.stderr.isatty():
        return iter(iterable)
    return _tqdm(iterable, desc=desc, **kwargs)


def descnext(desc):
    _state['desc'] = desc


def print(*args):  # noqa: A001  (API name)
    if not _state['quiet']:
        sys.stderr.write(' '.join(str(a) for a in args) + '\n')


@contextmanager
def quiet():
    _state['quiet'] += 1
    try:
        yield
    finally:
        _state['quiet'] -= 1


class _CallableModule(types.ModuleType):
    def __call__(self, iterable, **kwargs):
        return _iterate(iterable, **kwargs)


sys.modules[__name__].__class__ = _CallableModule
