"""rankcompv3.jl_b200 -- B200-native REO core of RankCompV3.jl behind the reference's own interface.

The directory name contains a dot, so load it with `__graft_entry__.load_package()` (importlib) rather
than a plain `import`.  Contents: csrc/ (CUDA kernels + the C ABI, built into libreo_cuda.so), _lib.py
(ctypes binding), api.py (identify_degs / McCullagh_test / get_major_reo_lower_count mirrors), reoa.py
(the reoa() driver mirror), synth.py (synthetic inputs of the benchmark shapes), dist.py (row-tile
sharding across one process per GPU).
"""
from . import _lib, api, synth  # noqa: F401
from .api import (DegResult, DeviceMatrix, McCullagh_test, Reo, get_major_reo_lower_count,  # noqa: F401
                  identify_degs, nccl_unique_id)
from .reoa import pseudobulk_group, reoa  # noqa: F401
