"""peapods_b200 — B200-native drop-in for the spin-sim sweep path of PeaBrane/peapods.

Hand-written sm_100a CUDA kernels behind the C ABI of ``include/peapods_b200.h``; this package is
the host-side mirror of the reference's ``IsingSimulation`` / ``Ising`` interface.  The CUDA
extension is mandatory: there is no CPU code path."""
from ._core import IsingSimulation, colouring, metropolis_lookup, nccl_unique_id
from .spin_models import GEOMETRIES, Ising

__all__ = ["Ising", "IsingSimulation", "GEOMETRIES", "colouring", "metropolis_lookup", "nccl_unique_id"]
