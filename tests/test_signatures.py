"""The drop-in boundary as data: every callable of the reference that `morgana_b200` mirrors must accept the reference's
parameters -- same names, same order, same defaults -- as a prefix of its own signature (additive keyword parameters with
defaults are allowed).  `tests/golden/signatures.json` was dumped from the unmodified reference by
`tests/golden/make_golden.py`."""
import inspect
import json
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, 'golden', 'signatures.json')) as f:
    REFERENCE = json.load(f)

# Deliberate differences, each documented in DESIGN.md section 1:
ALLOWED = {
    # speaker_id_list gains a default (parameters can be set in memory instead of loaded from files)
    'data.SpeakerDependentMeanVarianceNormaliser.__init__': {'speaker_id_list'},
    'data.SpeakerDependentMinMaxNormaliser.__init__': {'speaker_id_list'},
}


def _resolve(name):
    import morgana_b200 as mg
    import morgana_b200.viz.synthesis  # noqa: F401
    obj = mg
    for part in name.split('.'):
        obj = getattr(obj, part)
    return obj


@pytest.mark.parametrize('name', sorted(REFERENCE))
def test_signature_is_a_superset_of_the_reference(name):
    ours = list(inspect.signature(_resolve(name), follow_wrapped=False).parameters.values())
    theirs = REFERENCE[name]
    catch_all = any(p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD) for p in ours)
    assert catch_all or len(ours) >= len([t for t in theirs if not t[0].startswith('*')]), (name, ours, theirs)
    for i, (ref_name, ref_default) in enumerate(theirs):
        if ref_name.startswith('*'):
            continue
        if i >= len(ours):
            assert catch_all, (name, ref_name)
            continue
        p = ours[i]
        if p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD):
            break
        assert p.name == ref_name, '{}: parameter {} is {!r}, the reference has {!r}'.format(name, i, p.name, ref_name)
        if ref_name in ALLOWED.get(name, ()):
            continue
        our_default = '<required>' if p.default is p.empty else repr(p.default)
        assert our_default == ref_default, '{}: default of {!r} is {}, the reference has {}'.format(name, ref_name, our_default, ref_default)
    for p in ours[len(theirs):]:      # anything we add must be optional
        assert p.default is not p.empty or p.kind in (p.VAR_POSITIONAL, p.VAR_KEYWORD), (name, p.name)
