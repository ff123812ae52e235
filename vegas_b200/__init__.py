"""vegas_b200 -- a B200-native (sm_100a) vegas / vegas+ sampling engine behind the API of
gplepage/vegas: ``Integrator``, ``AdaptiveMap``, ``@lbatchintegrand`` / ``@rbatchintegrand``,
``RAvg`` / ``RAvgArray`` / ``RAvgDict``.  ``import vegas_b200 as vegas`` is the intended use.

The per-iteration hot path (allocate -> sample -> map -> evaluate -> per-hypercube reduce -> train)
runs in hand-written CUDA kernels behind the C ABI of ``libvegas_b200.so``
(``include/vegas_b200.h``).  There is no CPU fallback: without the library and a B200 the
sampling calls raise.
"""
from ._gv import gv as _gv, HAVE_GVAR
from ._map import AdaptiveMap
from ._integrand import (VegasIntegrand, LBatchIntegrand, RBatchIntegrand, BatchIntegrand, VecIntegrand,
                         DeviceIntegrand, lbatchintegrand, rbatchintegrand, batchintegrand,
                         devicebatchintegrand, vecintegrand, MPIintegrand)
from ._results import RAvg, RAvgArray, RAvgDict, VegasResult, reporter, ravg
from ._integrator import Integrator
from ._restratify import restratify, restratifyIntegrator, stratification_profile
from ._pdf import PDFIntegrator, PDFEV, PDFEVArray, PDFEVDict, PDFAnalyzer
from . import integrands

__version__ = '0.1.0'
ranseed = _gv.ranseed

__all__ = ['Integrator', 'AdaptiveMap', 'RAvg', 'RAvgArray', 'RAvgDict', 'VegasResult', 'reporter',
           'VegasIntegrand', 'LBatchIntegrand', 'RBatchIntegrand', 'BatchIntegrand', 'DeviceIntegrand',
           'lbatchintegrand', 'rbatchintegrand', 'batchintegrand', 'devicebatchintegrand', 'integrands',
           'restratify', 'restratifyIntegrator', 'ravg', 'PDFIntegrator', 'PDFEV', 'PDFEVArray', 'PDFEVDict',
           'ranseed']
