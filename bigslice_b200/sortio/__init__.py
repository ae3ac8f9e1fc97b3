"""Sort / merge / reduce streams (reference ``sortio/``)."""
from .sort import (FrameBuffer, MergeReader, NewMergeReader, SortReader,
                   merge_reader, sort_frame, sort_reader)
from .reader import Reduce, ReduceReader, combine_sorted, reduce_reader

__all__ = ["FrameBuffer", "MergeReader", "NewMergeReader", "SortReader",
           "merge_reader", "sort_frame", "sort_reader", "Reduce", "ReduceReader",
           "combine_sorted", "reduce_reader"]
