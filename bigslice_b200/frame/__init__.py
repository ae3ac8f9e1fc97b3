"""Columnar frame package (reference ``frame/``)."""
from .columns import (ListColumn, StringColumn, as_column, col_concat, col_len, col_slice,
                      col_take, col_tolist, col_zeros, col_from_list, is_column)
from .frame import (Frame, Copy, Make, Slices, Values, concat, copy, make,
                    slices, compatible, hash_column)
from .ops import Ops, RegisterOps, register_ops, unregister_ops, CanCompare, CanHash, can_compare, can_hash
from .codec import Key, Session
from . import hashing

__all__ = [
    "Frame", "ListColumn", "StringColumn", "Copy", "Make", "Slices", "Values", "concat", "copy",
    "make", "slices", "compatible", "hash_column", "Ops", "RegisterOps",
    "register_ops", "unregister_ops", "CanCompare", "CanHash", "can_compare",
    "can_hash", "Key", "Session", "hashing", "as_column", "col_concat", "col_len",
    "col_slice", "col_take", "col_tolist", "col_zeros", "col_from_list", "is_column",
]
