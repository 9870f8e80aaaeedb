"""
Duck-typed replacement of the `pmesh` package surface that nbodykit's FFTPower path touches
(SURVEY.md §8b): `pmesh.pm.{ParticleMesh, RealField, ComplexField, Field, BaseComplexField}`
and `pmesh.window.methods`.  The arithmetic behind every member is a CUDA kernel in
libnbk_b200.so; fields live in HBM as torch tensors (torch is the allocator / stream / NCCL
plumbing only).
"""
from . import window, pm  # noqa: F401
