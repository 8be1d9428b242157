"""
Locates the host framework (myokit) for the oracle.

TEST INFRASTRUCTURE ONLY. The oracle restates the arithmetic of the reference's
``SimulationOpenCL`` on the CPU; it walks a ``myokit.Model`` with myokit's own
model API, so it needs the ``myokit`` package. Search order: an installed
``myokit``; ``<repo>/baseline/_ref`` (the offline ``pip --target`` install of the
unmodified reference; git-ignored, travels to the GPU box); ``/root/reference``
(this container only).
"""
import os
import sys
import warnings

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _first_import_lock:
    """
    myokit creates ~/.config/myokit (bare os.makedirs) and writes myokit.ini
    there on its first import. N ranks starting together on a fresh machine
    race on both; one dies with FileExistsError. The directory is created
    tolerantly here and the import itself is serialised with a file lock.
    """
    def __enter__(self):
        self._f = None
        try:
            import fcntl
            d = os.path.join(os.path.expanduser('~'), '.config', 'myokit')
            os.makedirs(d, exist_ok=True)
            self._f = open(os.path.join(d, '.import.lock'), 'w')
            fcntl.flock(self._f, fcntl.LOCK_EX)
        except Exception:
            self._f = None
        return self

    def __exit__(self, *args):
        if self._f is not None:
            try:
                import fcntl
                fcntl.flock(self._f, fcntl.LOCK_UN)
                self._f.close()
            except Exception:
                pass
        return False


def import_myokit():
    """Returns the ``myokit`` module, or raises ImportError."""
    try:
        with _first_import_lock(), warnings.catch_warnings():
            warnings.simplefilter('ignore')
            import myokit
        return myokit
    except ImportError:
        pass
    for path in (os.path.join(_REPO, 'baseline', '_ref'), '/root/reference'):
        if os.path.isdir(os.path.join(path, 'myokit')):
            sys.path.insert(0, path)
            try:
                with _first_import_lock(), warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    import myokit
                return myokit
            except ImportError:
                sys.path.remove(path)
    raise ImportError(
        'myokit not found: install it, or create baseline/_ref with '
        '`python -m pip install --no-index --no-build-isolation --no-deps '
        '--target baseline/_ref /root/reference`.')
