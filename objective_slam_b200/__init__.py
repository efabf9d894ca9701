"""objective_slam_b200 -- the Drost PPF recognition hot path of objective-slam, rebuilt for B200 (sm_100a).

Importing the package loads lib/libppf_b200.so and fails loudly if it is missing (no CPU fallback).
"""
from . import io, operators, synth, voxel  # noqa: F401
from .api import Lookup, LookupResult, Model, Scene, ppf_registration  # noqa: F401
from ._capi import PpfError  # noqa: F401
