"""h263_rs_b200 -- B200-native reconstruction path for h263-rs streams.

Layout:
  csrc/         CUDA kernels, the C ABI (include/h263cu.h), host front end, stream generator
  _lib.py       ctypes loader (fails loudly when libh263cu.so is missing; no CPU fallback)
  frontend.py   host parse -> side info
  api.py        mirrors of the reference API: H263State, yuv420_to_rgba, deblock
  synth.py      synthetic bitstream generator
  build.py      in-tree nvcc build (sm_100a)
"""
__all__ = ["_lib", "frontend", "synth"]
