"""B200-native RDF and Einstein / Green-Kubo hot paths behind the MDSuite calculator API.

Layout:
  csrc/        hand-written sm_100a CUDA kernels + the C ABI (include/mdk.h) -> libmdk.so
  _lib.py      ctypes binding of libmdk.so (fails loudly when it is missing)
  kernels.py   torch-CUDA-tensor entry points over the C ABI
  engine.py    HBM-resident batch drivers (frame packing, window plans)
  ...          host-side mirror of the MDSuite calculator interface
"""
__version__ = "0.1.0"
