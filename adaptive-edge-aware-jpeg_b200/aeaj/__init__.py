"""aeaj -- host side of the B200-native adaptive edge-aware JPEG hot path.

``native`` binds libaeaj.so (hand-written sm_100a CUDA behind a C ABI, include/aeaj.h) with ctypes;
``codec`` drives it with torch tensors (device memory, streams); ``tables`` holds the host-side
settings tables.  The reference-facing packages ``jpeg``, ``color`` and ``image`` next to this one
mirror the reference's call surface on top of it.  There is no CPU fallback anywhere.
"""
