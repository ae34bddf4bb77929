"""Import the UNMODIFIED reference (ZackHodari/morgana at /root/reference) in the build container.

Test infrastructure only.  The reference cannot travel to the GPU box (``/root/reference`` does not
exist there), so this module is used by exactly one thing: ``tests/golden/make_golden.py``, which
runs the reference's own functions on seeded inputs and commits the results as fixtures.  Nothing in
``morgana_b200/``, ``bench.py`` or the ``-m gpu`` tests imports it.

The reference's ``morgana/__init__.py:1-10`` eagerly imports modules that depend on packages absent
from this image (``tts_data_tools``, ``tensorboardX``, ``bandmat``, ``matplotlib``, ``pyworld``).  None of
them carries arithmetic for the hot path, so empty stand-ins are registered
before the import.
"""
import json
import os
import sys
import types

REFERENCE_ROOT = os.environ.get('MORGANA_REFERENCE_ROOT', '/root/reference')


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _install_stand_ins():
    def load_json(path):
        with open(path) as f:
            return json.load(f)

    def _absent(*args, **kwargs):
        raise RuntimeError('stand-in for a package that is not installed in this image')

    class _Anything(object):
        def __init__(self, *args, **kwargs):
            pass

    if 'tts_data_tools' not in sys.modules:
        tdt = _module('tts_data_tools')
        tdt.file_io = _module('tts_data_tools.file_io', load_json=load_json, save_json=_absent, save_dir=_absent,
                              save_bin=_absent, save_wav=_absent, load_dir=_absent, load_bin=_absent)
        tdt.utils = _module('tts_data_tools.utils', get_file_ids=_absent)
        tdt.data_sources = _module('tts_data_tools.data_sources', TextSource=_Anything, NumpyBinarySource=_Anything)
        tdt.wav_gen = _module('tts_data_tools.wav_gen')
        tdt.wav_gen.world_with_reaper_f0 = _module('tts_data_tools.wav_gen.world_with_reaper_f0', synthesis=_absent)
    if 'tensorboardX' not in sys.modules:
        _module('tensorboardX', SummaryWriter=_Anything)
    if 'bandmat' not in sys.modules:
        bm = _module('bandmat')
        bm.linalg = _module('bandmat.linalg')
    if 'matplotlib' not in sys.modules:
        mpl = _module('matplotlib', use=lambda *a, **k: None)
        mpl.patches = _module('matplotlib.patches', Rectangle=_Anything)
        mpl.pyplot = _module('matplotlib.pyplot')
    if 'pyworld' not in sys.modules:
        _module('pyworld')


def import_reference():
    """Returns the reference's top-level ``morgana`` package, imported from ``REFERENCE_ROOT``."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, 'morgana')):
        raise FileNotFoundError('reference tree not found at {} (it only exists in the build container)'
                                .format(REFERENCE_ROOT))
    sys.dont_write_bytecode = True  # never write __pycache__ into the read-only reference tree
    _install_stand_ins()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import morgana
    return morgana
