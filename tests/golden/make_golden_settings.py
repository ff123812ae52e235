"""Golden fixture for ``Integrator.settings()`` / ``AdaptiveMap.settings()``: the strings the UNMODIFIED
reference prints (compiled ``_vegas`` of oracle/_ref, imported with the test-only gvar stand-in) for the
configurations in ``tests/golden/cases.py::SETTINGS``.  Build container only:

    make -C oracle ref && python tests/golden/make_golden_settings.py
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'gvar_shim'))
sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
sys.path.insert(0, ROOT)
import vegas           # noqa: E402  (the reference package)

from tests.golden.cases import SETTINGS, settings_limits      # noqa: E402

out = {}
for name, spec in SETTINGS.items():
    integ = vegas.Integrator(settings_limits(spec['limits']), **spec['kw'])
    out[name] = integ.settings(ngrid=spec.get('ngrid', 0))
    print('-----', name)
    print(out[name])
with open(os.path.join(HERE, 'ref_settings.json'), 'w') as f:
    json.dump(out, f, indent=1, sort_keys=True)
