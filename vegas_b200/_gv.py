"""``gv`` = the real ``gvar`` package when installed, else the built-in minimal implementation.

(The test-only stand-in under ``oracle/gvar_shim`` identifies itself with a ``-shim`` version and
is never used by product code, even if a test process has it on ``sys.path``.)
"""
try:                                    # pragma: no cover - depends on the environment
    import gvar as gv
    if str(getattr(gv, '__version__', '')).endswith('shim'):
        raise ImportError('test shim')
    HAVE_GVAR = True
except ImportError:
    from . import _gvbuiltin as gv
    HAVE_GVAR = False
