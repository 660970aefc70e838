"""spada-sim_b200 -- B200-native SpGEMM engine behind the spada-sim interface.

The product is ``lib/libspada_b200.so`` (hand-written sm_100a CUDA + the C ABI declared in
``include/spada_b200.h``).  This package is the host-side mirror of the reference's interface
for the hot path (same names and argument meaning as gemm.rs / py2rust.rs / storage.rs /
frontend.rs / simulator.rs / main.rs) plus a thin ctypes object model (``Engine``).

The directory name contains a hyphen, so import it with
``importlib.import_module("spada-sim_b200")``.  There is no CPU fallback: importing works
without a GPU, creating an ``Engine`` does not.
"""
from . import _abi, workloads
from ._abi import SpadaB200Error
from .engine import CBuf, DeviceCsr, Engine, Group, Result, Shard, device_count
from .frontend import Cli, OmegaConfig, parse_args, parse_config
from .gemm import GEMM
from .py2rust import load_mm_mat, load_pickled_gemms
from .simulator import Simulator
from .storage import CsrMatStorage, CsrRow, Element, sort_by_length

__all__ = ["Engine", "DeviceCsr", "Result", "CBuf", "Shard", "Group", "device_count", "SpadaB200Error", "GEMM", "load_mm_mat",
           "load_pickled_gemms", "Simulator", "CsrMatStorage", "CsrRow", "Element", "sort_by_length", "Cli",
           "OmegaConfig", "parse_args", "parse_config", "workloads"]
