"""pytassim_b200 — B200-native LETKF / ETKF analysis engine behind the pytassim.interface API.

Importing the package does not need a GPU; constructing an engine (or calling ``assimilate``) does, and raises if the
CUDA library ``libb200da.so`` has not been built or no sm_100 device is present.  There is no CPU fallback.
"""
__version__ = "0.1.0"
