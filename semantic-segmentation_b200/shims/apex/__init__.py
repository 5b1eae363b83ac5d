"""Minimal stand-in for NVIDIA Apex (absent from this image; the reference builds it unpinned, Dockerfile:21).

Only the names the reference touches on the hot path are provided: apex.amp.{float_function, half_function,
disable_casts, initialize, scale_loss} (network/mynn.py:10,42,51; loss/rmi.py:19,76; train.py:381,504) and
apex.parallel.{SyncBatchNorm, DistributedDataParallel} (config.py:218-220; network/__init__.py:38-39).
The b200seg path carries its own precision policy (bf16 storage, fp32 accumulate), so amp is a pass-through.
"""
from . import amp, parallel  # noqa: F401
