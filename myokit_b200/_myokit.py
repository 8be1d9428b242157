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


def import_myokit():
    with warnings.catch_warnings():
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
