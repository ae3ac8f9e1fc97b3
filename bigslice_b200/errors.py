"""Error kinds and severities that drive retry vs abort decisions.

A task that fails with a *fatal* error aborts the evaluation (``TaskErr``);
any other failure marks the task *lost* and it is recomputed.  This mirrors
how the reference uses ``grailbio/base/errors`` kinds/severities
(``exec/bigmachine.go:697-725``, ``exec/local.go:69-94``).
"""
from __future__ import annotations

import enum
from typing import Optional


class Kind(enum.Enum):
    Other = "other"
    Integrity = "integrity error"
    TooManyTries = "too many tries"
    Net = "network error"
    Unavailable = "resource unavailable"
    Temporary = "temporary error"
    Canceled = "operation was canceled"
    Timeout = "operation timed out"
    Invalid = "invalid argument"
    NotExist = "resource does not exist"
    Remote = "remote error"


class Severity(enum.IntEnum):
    Retriable = -2
    Temporary = -1
    Unknown = 0
    Fatal = 1


class Error(Exception):
    """An error with a kind, a severity and an optional cause."""

    def __init__(self, msg: str = "", kind: Kind = Kind.Other,
                 severity: Severity = Severity.Unknown,
                 cause: Optional[BaseException] = None):
        self.msg = msg
        self.kind = kind
        self.severity = severity
        self.cause = cause
        super().__init__(self._render())

    def _render(self) -> str:
        parts = []
        if self.msg:
            parts.append(self.msg)
        if self.kind is not Kind.Other:
            parts.append(self.kind.value)
        if self.severity is Severity.Fatal:
            parts.append("fatal")
        s = ": ".join(parts) if parts else "error"
        if self.cause is not None:
            s += ": %s" % (self.cause,)
        return s


class IntegrityError(Error):
    def __init__(self, msg="", cause=None):
        super().__init__(msg, Kind.Integrity, Severity.Unknown, cause)


class TooManyTriesError(Error):
    def __init__(self, msg="", cause=None):
        super().__init__(msg, Kind.TooManyTries, Severity.Unknown, cause)


class TaskLostError(Error):
    """``exec.ErrTaskLost``"""

    def __init__(self, msg="task was lost"):
        super().__init__(msg, Kind.Unavailable)


def fatal(msg_or_exc, cause: Optional[BaseException] = None) -> Error:
    """Wrap into a Fatal-severity error (``errors.E(errors.Fatal, ...)``)."""
    if isinstance(msg_or_exc, BaseException):
        return Error(str(msg_or_exc) if not isinstance(msg_or_exc, Error) else msg_or_exc.msg,
                     getattr(msg_or_exc, "kind", Kind.Other), Severity.Fatal,
                     cause if cause is not None else msg_or_exc)
    return Error(str(msg_or_exc), Kind.Other, Severity.Fatal, cause)


def is_fatal(e: BaseException) -> bool:
    seen = set()
    while e is not None and id(e) not in seen:
        seen.add(id(e))
        if isinstance(e, Error) and e.severity is Severity.Fatal:
            return True
        e = getattr(e, "cause", None) or e.__cause__
    return False


def is_kind(e: BaseException, kind: Kind) -> bool:
    seen = set()
    while e is not None and id(e) not in seen:
        seen.add(id(e))
        if isinstance(e, Error) and e.kind is kind:
            return True
        e = getattr(e, "cause", None) or e.__cause__
    return False


def is_temporary(e: BaseException) -> bool:
    return isinstance(e, Error) and (e.severity in (Severity.Temporary, Severity.Retriable)
                                     or e.kind is Kind.Temporary)


E = Error
Fatal = Severity.Fatal
Integrity = Kind.Integrity
TooManyTries = Kind.TooManyTries
