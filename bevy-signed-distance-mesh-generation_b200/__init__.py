"""sdfmesh-b200: B200-native (sm_100a) drop-in for the mesh-generation hot path of
Meterius/bevy-signed-distance-mesh-generation.

The directory name carries the reference's name and is not a valid Python identifier; import it through the
`bsdmg_b200` alias package at the repository root (``import bsdmg_b200``).

The product is the C-ABI shared library ``libsdfmesh.so`` (csrc/, include/sdfmesh.h).  The Python here is the
host-side mirror of the reference's ``CudaHandler`` (src/cuda/mod.rs) over that ABI, the scene tables of the
benchmark configurations, and the multi-GPU shard/gather plumbing (torch.distributed / NCCL).
"""
from .handler import CudaHandler, CudaVoxelField, Mesh, SdfMeshError, lib_path, load_library  # noqa: F401
from . import scenes  # noqa: F401

__all__ = ["CudaHandler", "CudaVoxelField", "Mesh", "SdfMeshError", "scenes", "lib_path", "load_library"]
