"""Streaming frame I/O (reference ``sliceio/``)."""
from .reader import (EOF, ClosingReader, EmptyReader, ErrReader, FrameReader,
                     FuncReader, LabelReader, MultiReader, PprofReader, ReadAll,
                     ReadFull, Reader, ReaderWithCloseFunc, Rebatch, RetryReader, current_label,
                     read_all, read_full)
from .codec import (Decoder, Encoder, NewDecodingReader, NewEncodingWriter, crc32,
                    decode_batch, encode_frame)
from .scanner import Scanner
from .spiller import NewSpiller, Spiller
from .writer import Writer, FuncWriter, DiscardWriter

__all__ = [
    "EOF", "Reader", "FuncReader", "EmptyReader", "ErrReader", "FrameReader",
    "MultiReader", "ReaderWithCloseFunc", "ClosingReader", "LabelReader",
    "PprofReader", "Rebatch", "RetryReader", "read_all", "read_full", "ReadAll", "ReadFull",
    "current_label", "Encoder", "Decoder", "NewEncodingWriter", "NewDecodingReader",
    "encode_frame", "decode_batch", "crc32", "Scanner", "Spiller", "NewSpiller",
    "Writer", "FuncWriter", "DiscardWriter",
]
