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


def import_myokit():
    """Returns the ``myokit`` module, or raises ImportError."""
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            import myokit
        return myokit
    except ImportError:
        pass
    for path in (os.path.join(_REPO, 'baseline', '_ref'), '/root/reference'):
        if os.path.isdir(os.path.join(path, 'myokit')):
            sys.path.insert(0, path)
            try:
                with warnings.catch_warnings():
                    warnings.simplefilter('ignore')
                    import myokit
                return myokit
            except ImportError:
                sys.path.remove(path)
    raise ImportError(
        'myokit not found: install it, or create baseline/_ref with '
        '`python -m pip install --no-index --no-build-isolation --no-deps '
        '--target baseline/_ref /root/reference`.')
