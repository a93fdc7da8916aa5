"""tt_sketch -- B200-native drop-in for the sketching hot path of RikVoorhaar/tt-sketch.

Same module paths and call signatures as the reference package for the path
`stream_sketch` / `orthogonal_sketch` / `blocked_stream_sketch` -> `general_sketch` ->
`DRM.sketch_*` + `sketch_omega_* / sketch_psi_*`; every contraction runs in hand-written
sm_100a CUDA kernels behind the C ABI of include/ttsk.h (libttsk.so, loaded with ctypes).
There is no CPU fallback: without a CUDA device the compute entry points raise.
"""
__version__ = "0.1.0"
