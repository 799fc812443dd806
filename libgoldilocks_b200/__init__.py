"""libgoldilocks_b200 -- host side of the B200-native batched Ed448-Goldilocks engine.

The product is the C-ABI shared library `libgoldilocks_b200.so` (CUDA, sm_100a only; header
include/goldilocks_b200.h).  This package only locates and binds it:

  load()            -> BatchLib over host (numpy) arrays: the `*_batch` entry points
  engine.Device...  -> device-resident entry points on torch CUDA tensors / streams, batch sharding

There is no CPU implementation here.  If the CUDA library is missing, load() raises.
"""
import os

from .capi import BatchLib, pack_messages, SUCCESS, FAILURE  # noqa: F401

# GOLDILOCKS_B200_LIB selects another build of the same CUDA library (kernel experiments); never a CPU library.
LIB_PATH = os.environ.get("GOLDILOCKS_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgoldilocks_b200.so")
_lib = None


def load():
    """Binds the CUDA library (built in-tree by `make lib` / __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libgoldilocks_b200.so is not built (run `make lib`); there is no CPU fallback")
        _lib = BatchLib(LIB_PATH)
    return _lib
