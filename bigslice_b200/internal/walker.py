"""Scanner-style recursive directory walker.

Capability parity: reference ``internal/walker/walker.go:14-99`` (``Walker.
{Init,Scan,Path,Info,Err}``): depth-first, lexical order within a directory,
errors sticky.
"""
from __future__ import annotations

import os
from typing import List, Optional


class Walker:
    def __init__(self, root: Optional[str] = None):
        self._todo: List[str] = []
        self._path = ""
        self._info: Optional[os.stat_result] = None
        self._err: Optional[BaseException] = None
        if root is not None:
            self.init(root)

    def init(self, root: str) -> None:
        self._todo = [root]
        self._path, self._info, self._err = "", None, None

    def scan(self) -> bool:
        """Advance to the next entry; False when done or on error."""
        if self._err is not None or not self._todo:
            return False
        self._path = self._todo.pop()
        try:
            self._info = os.stat(self._path)
        except OSError as e:
            if self._path is not None and os.path.islink(self._path):
                self._info = os.lstat(self._path)
                return True
            self._err = e
            return False
        if os.path.isdir(self._path):
            try:
                names = sorted(os.listdir(self._path), reverse=True)
            except OSError as e:
                self._err = e
                return False
            self._todo.extend(os.path.join(self._path, n) for n in names)
        return True

    def path(self) -> str:
        return self._path

    def info(self) -> Optional[os.stat_result]:
        return self._info

    def err(self) -> Optional[BaseException]:
        return self._err

    def __iter__(self):
        while self.scan():
            yield self._path
        if self._err is not None:
            raise self._err

    Init, Scan, Path, Info, Err = init, scan, path, info, err
