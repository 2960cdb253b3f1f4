"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front of oracle/rvq_oracle.c plus a pure-numpy restatement.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this.  See rvq_oracle.c for the
reference functions restated (encodec RVQ encode/decode; tts/dataloader.py code normalisation)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "librvq_oracle.so")
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def encode(latents: np.ndarray, codebooks: np.ndarray, threads: int = 1) -> np.ndarray:
    """latents [B, D, T] f32, codebooks [Q, K, D] f32 -> codes [B, Q, T] int64.
    threads > 1: clips are independent, so they are dealt to host threads (ctypes releases the GIL during the C call); every frame
    is still evaluated by one thread in the C file's fixed order, so the result does not depend on the thread count."""
    latents = np.ascontiguousarray(latents, np.float32)
    codebooks = np.ascontiguousarray(codebooks, np.float32)
    B, D, T = latents.shape
    Q, K, _ = codebooks.shape
    codes = np.empty((B, Q, T), np.int64)
    fn = lib().rvq_oracle_encode
    threads = max(1, min(int(threads), B))
    if threads == 1:
        fn(_ptr(latents), _ptr(codebooks), _ptr(codes), B, D, T, Q, K)
        return codes
    from concurrent.futures import ThreadPoolExecutor
    bounds = [B * i // threads for i in range(threads + 1)]

    def part(i):
        lo, hi = bounds[i], bounds[i + 1]
        if hi > lo:
            fn(_ptr(latents[lo:hi]), _ptr(codebooks), _ptr(codes[lo:hi]), hi - lo, D, T, Q, K)
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(part, range(threads)))
    return codes


def decode(codes: np.ndarray, codebooks: np.ndarray) -> np.ndarray:
    codes = np.ascontiguousarray(codes, np.int64)
    codebooks = np.ascontiguousarray(codebooks, np.float32)
    B, Q, T = codes.shape
    _, K, D = codebooks.shape
    lat = np.empty((B, D, T), np.float32)
    lib().rvq_oracle_decode(_ptr(codes), _ptr(codebooks), _ptr(lat), B, D, T, Q, K)
    return lat


def codes_affine(codes: np.ndarray) -> np.ndarray:
    codes = np.ascontiguousarray(codes, np.int64)
    out = np.empty(codes.shape, np.float32)
    lib().codes_affine_oracle(_ptr(codes), _ptr(out), C.c_long(codes.size))
    return out


def encode_fp64(latents: np.ndarray, codebooks: np.ndarray):
    """fp64 evaluation of the same argmax with the top-2 margin per (frame, stage) along the fp64 path:
    returns (codes [B,Q,T], margin [B,Q,T]) -- used to classify near-ties."""
    x = np.transpose(latents.astype(np.float64), (0, 2, 1)).reshape(-1, latents.shape[1])   # [N, D]
    B, D, T = latents.shape
    Q, K, _ = codebooks.shape
    codes = np.empty((Q, x.shape[0]), np.int64)
    margin = np.empty((Q, x.shape[0]), np.float64)
    r = x.copy()
    for q in range(Q):
        e = codebooks[q].astype(np.float64)
        dist = -((r * r).sum(1, keepdims=True) - 2 * r @ e.T + (e * e).sum(1)[None, :])
        idx = dist.argmax(1)
        top2 = np.partition(dist, -2, axis=1)[:, -2:]
        margin[q] = top2[:, 1] - top2[:, 0]
        codes[q] = idx
        r = r - e[idx]
    return (codes.reshape(Q, B, T).transpose(1, 0, 2).copy(), margin.reshape(Q, B, T).transpose(1, 0, 2).copy())
