"""Execution layer: compiler, tasks, evaluator, executors, sessions
(reference ``exec/``)."""
from .task import (Task, TaskDep, TaskName, TaskState, TaskSubscriber, TaskInit,
                   TaskWaiting, TaskRunning, TaskOk, TaskErr, TaskLost, iter_tasks,
                   make_graph)
from .compile import CompileEnv, ExecInvocation, compile_slice, pipeline
from .eval import Executor, Eval, evaluate, MAX_CONSECUTIVE_LOST
from .local import LocalExecutor, TaskBuffer
from .combiner import Combiner, CombiningFrame
from .session import (Bigmachine, DefaultMaxLoad, EventerOpt, ListEventer, Local,
                      LocalDevice, MachineCombiners, MaxLoad, NopEventer, Option, StoreOpt, Journal,
                      Parallelism, Result, Session, Spmd, Start, StatusOpt, TracePath,
                      start)
from .tracer import Tracer
from .topn import top_n

Status = StatusOpt
Eventer = EventerOpt

__all__ = [
    "Task", "TaskDep", "TaskName", "TaskState", "TaskSubscriber", "TaskInit",
    "TaskWaiting", "TaskRunning", "TaskOk", "TaskErr", "TaskLost", "iter_tasks",
    "make_graph", "CompileEnv", "ExecInvocation", "compile_slice", "pipeline",
    "Executor", "Eval", "evaluate", "MAX_CONSECUTIVE_LOST", "LocalExecutor",
    "TaskBuffer", "Combiner", "CombiningFrame", "Bigmachine", "DefaultMaxLoad",
    "Eventer", "EventerOpt", "ListEventer", "Local", "LocalDevice",
    "MachineCombiners", "MaxLoad", "NopEventer", "Option", "StoreOpt", "Journal", "Parallelism", "Result",
    "Session", "Spmd", "Start", "Status", "StatusOpt", "TracePath", "start",
    "Tracer", "top_n",
]
