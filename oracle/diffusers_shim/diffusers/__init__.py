"""Minimal restatement of the 12 `diffusers` 0.15.x symbols the reference imports.

TEST INFRASTRUCTURE ONLY (oracle). `diffusers ^0.15.1` (reference pyproject.toml:14) is a
third-party dependency that is absent from /root/reference and from this image; this package
restates its published semantics so that /root/reference/tts/{models,ldm/*}.py import and run
UNCHANGED with `PYTHONPATH=oracle/diffusers_shim:/root/reference`.  Parity at this third-party
boundary is therefore *unpinned* (no golden vectors from the real library exist offline).
"""
from .schedulers import DDPMScheduler  # noqa: F401
