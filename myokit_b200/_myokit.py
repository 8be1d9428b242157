"""
Locates the host framework. ``myokit_b200`` is a back-end plugin for Myokit: the
model / expression / protocol / DataLog machinery is Myokit's own and is
imported, not rebuilt (north_star: "reusing myokit's expression writers",
"pacing via myokit.Protocol").

Search order: an installed ``myokit``; then ``<repo>/baseline/_ref`` — the
offline ``pip install --target`` of the unmodified reference, which is
git-ignored but travels to the GPU box.
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
    with _first_import_lock(), warnings.catch_warnings():
        warnings.simplefilter('ignore')
        try:
            import myokit
            return myokit
        except ImportError:
            pass
        path = os.path.join(_REPO, 'baseline', '_ref')
        if os.path.isdir(os.path.join(path, 'myokit')):
            sys.path.insert(0, path)
            try:
                import myokit
                return myokit
            except ImportError:
                sys.path.remove(path)
    raise ImportError(
        'myokit_b200 is a back-end for Myokit and needs the `myokit` package. '
        'Install it, or create baseline/_ref with `python -m pip install '
        '--no-index --no-build-isolation --no-deps --target baseline/_ref '
        '<path to myokit source>`.')
