"""unomol_b200 -- B200-native (sm_100a, FP64) integral-direct J/K Fock build behind the call surface of
PatNichols/unomol's TwoElectronInts / RestrictedHartreeFock / UnRestrictedHartreeFock.

The product is libunomol_b200.so (C ABI: include/unomol_b200.h; CUDA kernels: unomol_b200/csrc).  This
package is the thin Python view used by tests/ and bench.py: `capi` binds the C ABI with ctypes (and fails
loudly when the library is missing -- there is no CPU fallback), `basis` reads the reference's patin.dat
format, `scf` mirrors the reference's RHF/UHF drivers on top of the C ABI.
"""
from .basis import Basis, water_cluster  # noqa: F401
