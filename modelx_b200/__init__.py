"""modelx-b200: B200-native blob digest-and-chunk engine for the modelx push/pull hot path.

The product is the C-ABI shared library ``libmodelxdigest.so`` (hand-written sm_100a CUDA);
this package is its Python host-side mirror of ``pkg/client``'s digest path.
"""
from ._native import LIB_PATH, MxdError, load  # noqa: F401
from .engine import (DEFAULT_CHUNK, DEFAULT_FANOUT, DEFAULT_LEAF, Engine, Hasher, batch_pays_off, calc_parts,  # noqa: F401
                     digest_parse, digest_string, server_part_count, tree_shape)

__all__ = ["Engine", "Hasher", "calc_parts", "server_part_count", "digest_string", "digest_parse", "tree_shape", "batch_pays_off",
           "DEFAULT_CHUNK", "DEFAULT_LEAF", "DEFAULT_FANOUT", "MxdError", "load", "LIB_PATH"]
