"""Import the UNMODIFIED reference (ZackHodari/morgana) for fixture generation.  Test infrastructure only.

Kept for ``tests/golden/make_golden.py`` and the live comparisons; the loading itself (stand-ins for the third-party packages
this image lacks, choice between ``/root/reference`` and its mirror ``oracle/_ref``) lives in ``oracle/ref_loader.py``.
Fixtures are always generated from ``/root/reference`` itself, so the root is pinned here.
"""
import os
import sys

REPO_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if REPO_ROOT not in sys.path:
    sys.path.insert(0, REPO_ROOT)

REFERENCE_ROOT = os.environ.get('MORGANA_REFERENCE_ROOT', '/root/reference')


def import_reference():
    """Returns the reference's top-level ``morgana`` package, imported from ``REFERENCE_ROOT``."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, 'morgana')):
        raise FileNotFoundError('reference tree not found at {} (it only exists in the build container)'
                                .format(REFERENCE_ROOT))
    os.environ['MORGANA_REFERENCE_ROOT'] = REFERENCE_ROOT
    from oracle import ref_loader
    return ref_loader.import_reference()
