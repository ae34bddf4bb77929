#!/usr/bin/env python
"""Recipe that makes the UNMODIFIED reference travel to the GPU box.  TEST / BASELINE INFRASTRUCTURE ONLY.

    python oracle/make_ref.py [--check]

ZackHodari/morgana is pure Python (no build step), so "compiling the reference" is a copy: this script mirrors
``/root/reference/morgana`` and ``/root/reference/models`` (plus ``LICENSE``, MIT) into ``oracle/_ref/`` byte for byte and
writes ``oracle/_ref/MANIFEST.json`` with a sha256 per file.  ``oracle/_ref/`` is git-ignored (no reference source ever
enters the history) but not gpurun-ignored, so -- exactly like the built ``.so`` -- it ships with the snapshot to the GPU
box, where ``/root/reference`` does not exist.  It is used there by

* ``bench.py --impl reference``: the reference's own functions on the box's host cores (``cpu_baseline.kind = "reference"``);
* the ``-m gpu`` tests of ``tests/test_reference_models_gpu.py``: ``patch()`` applied to the real package, the real
  ``models/RNN_SPSS.py`` / ``models/f0_test_model.py`` and the real ``ExperimentBuilder.train_epoch`` on CUDA against the
  unpatched reference on the CPU.

Nothing under ``morgana_b200/`` imports it.  ``oracle/ref_loader.py`` does the importing (with the stand-ins for the
third-party packages that are absent from this image).  ``__graft_entry__.build()`` runs this recipe whenever
``/root/reference`` is present; on the GPU box only the prebuilt copy is used.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get('MORGANA_REFERENCE_ROOT', '/root/reference')
DEST = os.path.join(HERE, '_ref')
TREES = ('morgana', 'models')
FILES = ('LICENSE',)


def _sha256(path):
    h = hashlib.sha256()
    with open(path, 'rb') as f:
        for block in iter(lambda: f.read(1 << 20), b''):
            h.update(block)
    return h.hexdigest()


def _source_files():
    for tree in TREES:
        for root, dirs, files in os.walk(os.path.join(REFERENCE_ROOT, tree)):
            dirs[:] = sorted(d for d in dirs if d != '__pycache__')
            for name in sorted(files):
                if name.endswith('.pyc'):
                    continue
                yield os.path.relpath(os.path.join(root, name), REFERENCE_ROOT)
    for name in FILES:
        if os.path.exists(os.path.join(REFERENCE_ROOT, name)):
            yield name


def is_current():
    """True when oracle/_ref holds exactly the files of the reference tree (by sha256)."""
    manifest_path = os.path.join(DEST, 'MANIFEST.json')
    if not os.path.exists(manifest_path):
        return False
    with open(manifest_path) as f:
        manifest = json.load(f)['files']
    if not os.path.isdir(REFERENCE_ROOT):
        return all(os.path.exists(os.path.join(DEST, rel)) and _sha256(os.path.join(DEST, rel)) == digest
                   for rel, digest in manifest.items())
    wanted = {rel: _sha256(os.path.join(REFERENCE_ROOT, rel)) for rel in _source_files()}
    return wanted == manifest and all(
        os.path.exists(os.path.join(DEST, rel)) and _sha256(os.path.join(DEST, rel)) == digest
        for rel, digest in manifest.items())


def make():
    """Mirror the reference into oracle/_ref (idempotent).  Returns the destination."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, 'morgana')):
        raise FileNotFoundError('reference tree not found at {} (it only exists in the build container)'.format(REFERENCE_ROOT))
    if is_current():
        return DEST
    tmp = DEST + '.tmp'
    shutil.rmtree(tmp, ignore_errors=True)
    manifest = {}
    for rel in _source_files():
        dst = os.path.join(tmp, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REFERENCE_ROOT, rel), dst)     # bytes only: never executes or edits the reference
        manifest[rel] = _sha256(dst)
    with open(os.path.join(tmp, 'MANIFEST.json'), 'w') as f:
        json.dump({'source': 'ZackHodari/morgana (MIT), mirrored unmodified from ' + REFERENCE_ROOT,
                   'files': manifest}, f, indent=1, sort_keys=True)
    shutil.rmtree(DEST, ignore_errors=True)
    os.replace(tmp, DEST)
    return DEST


if __name__ == '__main__':
    if '--check' in sys.argv:
        ok = is_current()
        print('oracle/_ref is', 'current' if ok else 'missing or stale')
        sys.exit(0 if ok else 1)
    print(make())
