"""surmise_b200 -- B200-native (sm_100a, fp64) PCGPwM emulation and directbayeswoodbury calibration
for bandframework/surmise, registered as additional surmise method modules.

    import surmise_b200; surmise_b200.register()
    emu = emulator(x, theta, f, method='PCGPwM_B200')
    cal = calibrator(emu, y, x, thetaprior, method='directbayeswoodbury_B200', yvar=yvar,
                     args={'sampler': 'PTLMC_B200'})      # or surmise's own 'PTLMC', 'LMC', 'metropolis_hastings'

The compute path is a C-ABI CUDA library (include/surmise_b200.h, surmise_b200/csrc); there is no CPU
fallback -- calling any method without a CUDA device raises.
"""
import importlib
import sys

EMULATION_METHODS = ('PCGPwM_B200', 'PCGPwImpute_B200', 'indGP_B200', 'PCSK_B200')
CALIBRATION_METHODS = ('directbayeswoodbury_B200', 'mlbayeswoodbury_B200')
UTILITIES_METHODS = ('PTLMC_B200', 'LMC_B200', 'metropolis_hastings_B200')

__version__ = '0.1.0'


def register():
    """Make the B200 method modules importable under surmise's plugin namespaces.

    surmise looks methods up with importlib.import_module('surmise.emulationmethods.' + method)
    (surmise/emulation.py:145-149) and 'surmise.calibrationmethods.' + method
    (surmise/calibration.py:196-200).  Entering the modules in sys.modules satisfies that lookup;
    setting them as attributes of the parent packages keeps emu.save_to / load_from working (dill
    pickles the method module by qualified name).  Call before load_from.
    """
    import surmise.emulationmethods as _em
    import surmise.calibrationmethods as _cm
    import surmise.utilitiesmethods as _um
    for name in EMULATION_METHODS:
        mod = importlib.import_module('surmise_b200.emulationmethods.' + name)
        sys.modules['surmise.emulationmethods.' + name] = mod
        setattr(_em, name, mod)
    for name in CALIBRATION_METHODS:
        mod = importlib.import_module('surmise_b200.calibrationmethods.' + name)
        sys.modules['surmise.calibrationmethods.' + name] = mod
        setattr(_cm, name, mod)
    for name in UTILITIES_METHODS:            # sampler plugins: 'surmise.utilitiesmethods.' + name (utilities.py:62-63)
        mod = importlib.import_module('surmise_b200.utilitiesmethods.' + name)
        sys.modules['surmise.utilitiesmethods.' + name] = mod
        setattr(_um, name, mod)
    return EMULATION_METHODS + CALIBRATION_METHODS + UTILITIES_METHODS


def covmat(x1, x2, gammav, return_gradhyp=False, return_gradx1=False):
    """Drop-in for surmise.emulationsupport.matern_covmat.covmat (numpy in / numpy out, computed on the GPU)."""
    from . import ops
    return ops.covmat(x1, x2, gammav, return_gradhyp, return_gradx1)
