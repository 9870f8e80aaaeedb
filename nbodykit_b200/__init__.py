"""
nbodykit_b200 -- a B200-native FFTPower pipeline behind the nbodykit API.

Keeps the runtime/config surface of `nbodykit/__init__.py` that the FFTPower path touches:
`CurrentMPIComm` (:107-191), `_global_options` / `set_options` (:22-25, 215-256),
`setup_logging` (:259-300).  The "communicator" is a thin shim over torch.distributed
(NCCL, one process per GPU) -- see nbodykit_b200/comm.py.
"""
import logging
import warnings
from contextlib import contextmanager

from .version import __version__  # noqa: F401

_global_options = {}
_global_options["global_cache_size"] = 1e8   # kept for API compatibility; there is no dask cache
_global_options["dask_chunk_size"] = 100000  # idem
_global_options["paint_chunk_size"] = 1024 * 1024 * 4   # nbodykit/__init__.py:25


class CurrentMPIComm(object):
    """get / set the current default communicator (nbodykit/__init__.py:107-191)"""
    _stack = []
    logger = logging.getLogger("CurrentMPIComm")

    @staticmethod
    def enable(func):
        """decorator: fill the ``comm`` keyword with the current communicator when it is None"""
        import functools

        @functools.wraps(func)
        def wrapped(*args, **kwargs):
            kwargs.setdefault("comm", None)
            if kwargs["comm"] is None:
                kwargs["comm"] = CurrentMPIComm.get()
            return func(*args, **kwargs)
        return wrapped

    @classmethod
    def _ensure(cls):
        if not cls._stack:
            from .comm import world
            cls._stack.append(world())
        elif len(cls._stack) == 1:
            from .comm import world, SelfComm
            if isinstance(cls._stack[0], SelfComm):
                cls._stack[0] = world()   # picks up a process group initialised after import

    @classmethod
    @contextmanager
    def enter(cls, comm):
        cls.push(comm)
        try:
            yield
        finally:
            cls.pop()

    @classmethod
    def push(cls, comm):
        cls._ensure()
        cls._stack.append(comm)
        if comm.rank == 0:
            cls.logger.info("Entering a current communicator of size %d" % comm.size)
        cls._stack[-1].barrier()

    @classmethod
    def pop(cls):
        comm = cls._stack[-1]
        if comm.rank == 0:
            cls.logger.info("Leaving current communicator of size %d" % comm.size)
        cls._stack[-1].barrier()
        cls._stack.pop()
        comm = cls._stack[-1]
        if comm.rank == 0:
            cls.logger.info("Restored current communicator to size %d" % comm.size)

    @classmethod
    def get(cls):
        cls._ensure()
        return cls._stack[-1]

    @classmethod
    def set(cls, comm):
        warnings.warn("CurrentMPIComm.set is deprecated. Use `with CurrentMPIComm.enter(comm):` instead")
        cls._ensure()
        cls._stack[-1].barrier()
        cls._stack[-1] = comm
        cls._stack[-1].barrier()


class set_options(object):
    """set global options, also usable as a context manager (nbodykit/__init__.py:215-256)"""

    def __init__(self, **kwargs):
        self.old = _global_options.copy()
        for key in sorted(kwargs):
            if key not in _global_options:
                raise KeyError("Option `%s` is not supported" % key)
        _global_options.update(kwargs)

    def __enter__(self):
        return

    def __exit__(self, type, value, traceback):
        _global_options.clear()
        _global_options.update(self.old)


_logging_handler = None


def setup_logging(log_level="info"):
    """turn on logging with elapsed-seconds + rank prefix (nbodykit/__init__.py:259-300)"""
    import time
    levels = {"info": logging.INFO, "debug": logging.DEBUG, "warning": logging.WARNING}
    logger = logging.getLogger()
    t0 = time.time()
    rank = CurrentMPIComm.get().rank

    class Formatter(logging.Formatter):
        def format(self, record):
            return ("[ %09.2f ] % 3d: " % (time.time() - t0, rank)) + logging.Formatter.format(self, record)

    fmt = Formatter(fmt="%(asctime)s %(name)-15s %(levelname)-8s %(message)s", datefmt="%m-%d %H:%M ")
    global _logging_handler
    if _logging_handler is None:
        _logging_handler = logging.StreamHandler()
        logger.addHandler(_logging_handler)
    _logging_handler.setFormatter(fmt)
    logger.setLevel(levels[log_level])
