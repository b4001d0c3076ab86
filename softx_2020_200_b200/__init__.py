"""B200-native GLS Navier–Stokes hot path (Jacobian/residual assembly + ILU(0)-GMRES) behind
Lethe's assemble_matrix_and_rhs / assemble_rhs / solve_linear_system interface.

The compute lives in libglsns.so (hand-written sm_100a CUDA behind the C ABI of include/glsns.h);
this package holds the ctypes binding and the host-side mirror of the reference interface."""
from .hotpath import GLSHotPath, GlsnsError, NoConvergence  # noqa: F401
