"""Import the UNMODIFIED reference (ZackHodari/morgana).  TEST / BASELINE INFRASTRUCTURE ONLY.

Where it comes from, in this order: ``$MORGANA_REFERENCE_ROOT``; ``oracle/_ref/`` (the byte-for-byte mirror made by
``oracle/make_ref.py``, which travels to the GPU box like the built ``.so``); ``/root/reference`` (build container only).

The reference's ``morgana/__init__.py:1-10`` eagerly imports modules that depend on packages absent from this image and not
installable offline: ``tts_data_tools`` (same author; file I/O only: call sites ``data.py:135,142,345,516``),
``tensorboardX``, ``matplotlib``, ``pyworld`` -- none carries arithmetic for the hot path, so empty
stand-ins are registered before the import -- and ``bandmat``, MLPG's banded solver, which gets a functional stand-in
(``oracle/bandmat_standin.py``).

Used by: ``tests/golden/make_golden.py``, ``tests/test_oracle_vs_reference_live.py``, ``tests/test_reference_models_gpu.py``
and ``bench.py --impl reference`` / ``cpu_baseline``.  Nothing under ``morgana_b200/`` imports it.
"""
import importlib
import importlib.util
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
MIRROR = os.path.join(HERE, '_ref')


def reference_root():
    """Directory holding ``morgana/`` and ``models/`` of the reference, or None."""
    for cand in (os.environ.get('MORGANA_REFERENCE_ROOT'), MIRROR, '/root/reference'):
        if cand and os.path.isfile(os.path.join(cand, 'morgana', '__init__.py')):
            return cand
    return None


def available():
    return reference_root() is not None


def _module(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def install_stand_ins():
    """Stand-ins for third-party packages the reference imports at module level and this image lacks."""
    def load_json(path):
        with open(path) as f:
            return json.load(f)

    def save_json(obj, path):
        with open(path, 'w') as f:
            json.dump(obj, f)

    def _absent(*args, **kwargs):
        raise RuntimeError('stand-in for a package that is not installed in this image')

    class _Anything(object):
        def __init__(self, *args, **kwargs):
            pass

        def __getattr__(self, name):                 # e.g. SummaryWriter().add_scalar(...): accepted, ignored
            if name.startswith('__'):
                raise AttributeError(name)
            return lambda *a, **k: None

    if 'tts_data_tools' not in sys.modules:
        tdt = _module('tts_data_tools')
        tdt.file_io = _module('tts_data_tools.file_io', load_json=load_json, save_json=save_json, save_dir=_absent,
                              save_bin=_absent, save_wav=_absent, load_dir=_absent, load_bin=_absent)
        tdt.utils = _module('tts_data_tools.utils', get_file_ids=_absent)
        tdt.data_sources = _module('tts_data_tools.data_sources', TextSource=_Anything, NumpyBinarySource=_Anything)
        tdt.wav_gen = _module('tts_data_tools.wav_gen')
        tdt.wav_gen.world_with_reaper_f0 = _module('tts_data_tools.wav_gen.world_with_reaper_f0', synthesis=_absent)
    if 'tensorboardX' not in sys.modules:
        _module('tensorboardX', SummaryWriter=_Anything)
    try:
        importlib.import_module('bandmat')
    except ImportError:
        from oracle import bandmat_standin
        bandmat_standin.install()
    if 'matplotlib' not in sys.modules:
        try:
            importlib.import_module('matplotlib')
        except ImportError:
            mpl = _module('matplotlib', use=lambda *a, **k: None)
            mpl.patches = _module('matplotlib.patches', Rectangle=_Anything)
            mpl.pyplot = _module('matplotlib.pyplot')
    if 'pyworld' not in sys.modules:
        try:
            importlib.import_module('pyworld')
        except ImportError:
            _module('pyworld')


def import_reference():
    """Returns the reference's top-level ``morgana`` package."""
    root = reference_root()
    if root is None:
        raise FileNotFoundError('the reference is neither mirrored under oracle/_ref (python oracle/make_ref.py, build '
                                'container only) nor present at /root/reference')
    if 'morgana' in sys.modules:
        loaded = os.path.dirname(os.path.dirname(os.path.abspath(sys.modules['morgana'].__file__)))
        if os.path.realpath(loaded) != os.path.realpath(root):
            raise RuntimeError('a different `morgana` is already imported from ' + loaded)
        return sys.modules['morgana']
    sys.dont_write_bytecode = True       # never write __pycache__ next to the reference's sources
    install_stand_ins()
    if root not in sys.path:
        sys.path.insert(0, root)
    import morgana
    return morgana


def load_model_module(name):
    """``models/<name>.py`` of the reference as a module (e.g. 'RNN_SPSS', 'f0_test_model'); executed unmodified.

    Model scripts bind names at import (``from morgana.viz.synthesis import MLPG``, models/RNN_SPSS.py:9), so call
    ``morgana_b200.patch()`` *before* this when the kernels are wanted, and give each variant its own ``alias``.
    """
    return load_model_module_as(name, 'morgana_reference_models_' + name)


def load_model_module_as(name, alias):
    import_reference()
    path = os.path.join(reference_root(), 'models', name + '.py')
    spec = importlib.util.spec_from_file_location(alias, path)
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module
