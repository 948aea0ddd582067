"""cheeta-mpc_b200 -- batched B200-native centroidal-MPC condensed-QP solver.

Python here is a thin ctypes view of the C ABI in ``include/cmpc.h`` (the product is
``libcmpc_b200.so``: CUDA kernels for sm_100a + C ABI; host mirror in C++ under
``include/CentroidalMPC.h``).  The directory name carries a hyphen, so import it with
``__graft_entry__.load_package()`` (registers it as ``cheeta_mpc_b200``).

There is NO CPU fallback: every call goes through the CUDA library and raises if it is
missing or no GPU is present.
"""
from . import abi, workloads  # noqa: F401
from .abi import (CmpcConfig, CmpcGait, CmpcStats, make_gait, CentroidalMPC, CmpcError, lib_path, load_library,  # noqa: F401
                  make_config, STATUS_NAMES)

__all__ = ["abi", "workloads", "CmpcConfig", "CmpcStats", "CentroidalMPC", "CmpcError",
           "lib_path", "load_library", "make_config", "STATUS_NAMES"]
