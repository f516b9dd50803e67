"""asif_b200 -- B200-native batched safety-filter engine (drop-in for the filter path of DrewSingletary/asif).

The product is the C-ABI shared library ``libasif_b200.so`` (include/asif_b200.h) with the C++
host layer on top of it (asif_b200/host/).  This Python package is only a thin ctypes binding used
by the tests and bench.py; it contains no numerics and no fallback: if the CUDA library is missing
or there is no device, calls fail loudly.
"""
from .capi import (  # noqa: F401
    FILTER_EXPLICIT, FILTER_IMPLICIT_TB, FILTER_IMPLICIT, FILTER_ROBUST, FILTER_REALIZABLE, FILTER_IMPLICIT_RB,
    MODEL_DOUBLE_INTEGRATOR, MODEL_DOUBLE_INTEGRATOR_TB, MODEL_INVERTED_PENDULUM, MODEL_INVERTED_PENDULUM_TABLE, MODEL_INVERTED_PENDULUM_KERNEL,
    MODEL_SEGWAY, MODEL_SEGWAY_SHIPPED, MEM_HOST, MEM_DEVICE,
    AsifError, Engine, EngineGroup, PinnedArray, HOST_IO_STAGED, HOST_IO_OUT, HOST_IO_INOUT, EngineConfig, LoopConfig, loop_log_fields, device_count, lib_path, load_library, measure_fp64_peak, qp_solve_batch,
)
