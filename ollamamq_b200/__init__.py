"""ollamamq_b200 — B200-native hot path of ollamaMQ (fair-share dispatch + on-box sm_100a GPU workers).

The package is a thin ctypes mirror of the C ABI in include/ollamamq_b200.h.  There is no Python or CPU
implementation of the forward pass: importing the package without a built libollamamq_b200.so raises.
"""
from ._lib import lib, MQError, last_error, check, LIB_PATH  # noqa: F401
from .dispatcher import Scheduler, Dispatch  # noqa: F401
from .worker import Worker, Dispatcher, Stream, model_cfg, Encoder, encoder_cfg  # noqa: F401
from . import models  # noqa: F401

__all__ = ["lib", "MQError", "last_error", "check", "Scheduler", "Dispatch", "LIB_PATH", "Worker", "Dispatcher",
           "Stream", "model_cfg", "Encoder", "encoder_cfg"]
