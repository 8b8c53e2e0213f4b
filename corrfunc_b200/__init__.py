"""corrfunc_b200 -- B200-native (sm_100a) gridded pair counting behind Corrfunc's own API.

Drop-in for the reference's hot path only: ``corrfunc_b200.theory.{DD,DDrppi,DDsmu,wp,xi}`` and
``corrfunc_b200.mocks.DDtheta_mocks`` mirror ``Corrfunc.theory.*`` / ``Corrfunc.mocks.DDtheta_mocks``
and call the C-ABI library ``csrc/libcorrfunc_b200.so`` (CUDA, no CPU fallback).
"""
__version__ = "0.1.0"
