"""Golden fixture for ``vegas.restratify``: runs the UNMODIFIED reference -- the compiled
``_vegas`` module of oracle/_ref plus the reference's own ``src/vegas/__init__.py`` (linked, not
copied, into a scratch package directory), imported with the test-only gvar stand-in -- on a 4-D
integrand whose first two axes carry all the structure.  Uniforms are injected through
``ran_array_generator`` (numpy default_rng), so the oracle can replay them.  Build container only:

    make -C oracle ref && python tests/golden/make_golden_restratify.py
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_INIT = '/root/reference/src/vegas/__init__.py'

pkg = tempfile.mkdtemp(prefix='refpkg_')
os.makedirs(os.path.join(pkg, 'vegas'))
refdir = os.path.join(ROOT, 'oracle', '_ref', 'vegas')
so = [n for n in os.listdir(refdir) if n.startswith('_vegas') and n.endswith('.so')][0]
os.symlink(os.path.join(refdir, so), os.path.join(pkg, 'vegas', so))
os.symlink(REF_INIT, os.path.join(pkg, 'vegas', '__init__.py'))
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'gvar_shim'))
sys.path.insert(0, pkg)
sys.path.insert(0, ROOT)
import gvar            # noqa: E402  (the shim)
import vegas           # noqa: E402  (the reference package)

from tests.golden.cases import RESTRATIFY, integrand      # noqa: E402

out = {}
for name, spec in RESTRATIFY.items():
    gvar.ranseed(1)
    rng = np.random.default_rng(spec['seed'])
    integ = vegas.Integrator(spec['limits'], ran_array_generator=lambda shape: rng.random(shape), **spec['kw'])
    f = vegas.lbatchintegrand(integrand(spec['f']))
    integ(f, nitn=spec['nitn_adapt'])
    out[name + '_grid'] = np.array(integ.map.grid, float)
    out[name + '_sigf'] = np.array(integ.sigf, float)
    out[name + '_sum_sigf'] = float(integ.sum_sigf)
    out[name + '_old_nstrat'] = np.array(integ.nstrat, np.int64)
    rng2 = np.random.default_rng(spec['seed'] + 1000)
    integ.set(ran_array_generator=lambda shape: rng2.random(shape))
    new = vegas.restratify(integ, f, nitn=spec['nitn'], ndy=spec['ndy'], **spec['opt'])
    out[name + '_I'] = np.array([new.I.mean, new.I.var])
    out[name + '_dI_mean'] = np.array(gvar.mean(new.dI), float)
    out[name + '_dI_var'] = np.array(gvar.var(new.dI), float)
    out[name + '_weight'] = np.array(gvar.mean(new.weight), float)
    out[name + '_new_nstrat'] = np.array(new.nstrat, np.int64)
    out[name + '_new_neval'] = np.int64(new.neval)
    print(name, 'old', list(integ.nstrat), 'new', list(new.nstrat), 'I', new.I, 'weights', out[name + '_weight'])
np.savez_compressed(os.path.join(HERE, 'ref_restratify.npz'), **out)
