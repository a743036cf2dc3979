"""clrsdp — B200-native interior-point hot path for clustered low-rank SDPs.

Host-side mirror of the reference's interface (MPMP.jl exports `solvempmp, solverank1sdp,
get_block_info, prepareabc, laguerrebasis`, :19) over the C ABI in include/clrsdp.h. The compute path is
the CUDA library csrc/libclrsdp.so; nothing here computes on the CPU.
"""
from .wire import MpArray  # noqa: F401
from .solver import (BlockInfo, Constraint, get_block_info, precision, product_handle,  # noqa: F401
                     set_precision, solverank1sdp)
