"""Client side of the shim: connection to the GPU server, content keys of static arrays."""
import atexit
import os
import subprocess
import sys
import tempfile
import time
import zlib
from multiprocessing.connection import Client

import numpy as np

_ENV = "PB200_SHIM_ADDR"
_AUTH = b"pb200-shim"
_state = {"pid": None, "conn": None, "server": None}
_keys = {}          # (id, data pointer, shape, dtype) -> content key, per process
FULL_HASH_BYTES = 64 << 20


def _start_server():
    addr = os.path.join(tempfile.mkdtemp(prefix="pb200shim_"), "sock")
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    proc = subprocess.Popen([sys.executable, "-m", "pyratbay_b200.shim.server", addr,
                             str(os.getpid())], env=env)
    for _ in range(600):
        if os.path.exists(addr):
            break
        if proc.poll() is not None:
            raise RuntimeError("pb200 shim: the GPU server exited at start-up "
                               "(no usable CUDA device? there is no CPU fallback)")
        time.sleep(0.1)
    else:
        proc.kill()
        raise RuntimeError("pb200 shim: the GPU server did not come up")
    os.environ[_ENV] = addr
    _state["server"] = proc
    atexit.register(_stop_server, os.getpid())


def _stop_server(owner_pid):
    proc = _state.get("server")
    if proc is not None and os.getpid() == owner_pid and proc.poll() is None:
        try:
            request("shutdown")
        except Exception:
            proc.terminate()
        try:
            proc.wait(timeout=10)
        except subprocess.TimeoutExpired:
            proc.kill()


def connection():
    """One connection per process (a forked child opens its own)."""
    if _state["pid"] != os.getpid() or _state["conn"] is None:
        if _ENV not in os.environ:
            _start_server()
        _state["conn"] = Client(os.environ[_ENV], family="AF_UNIX", authkey=_AUTH)
        _state["pid"] = os.getpid()
    return _state["conn"]


def request(op, *args):
    conn = connection()
    conn.send((op,) + args)
    status, payload = conn.recv()
    if status == "need":            # the server lacks static arrays: send them, then retry
        for key in payload:
            arr = _by_key[key]
            conn.send(("put", key, str(arr.dtype), arr.shape))
            conn.send_bytes(memoryview(np.ascontiguousarray(arr)).cast("B"))
            st, _ = conn.recv()
            if st != "ok":
                raise RuntimeError("pb200 shim: upload failed")
        conn.send((op,) + args)
        status, payload = conn.recv()
    if status == "error":
        _by_key.clear()
        raise RuntimeError(f"pb200 shim server: {payload}")
    _by_key.clear()                 # the uploads (if any) are done: drop the references
    if status == "bytes":           # (dtype, shape) followed by the raw buffer
        dtype, shape = payload
        out = np.empty(shape, dtype)
        conn.recv_bytes_into(memoryview(out).cast("B"))
        return out
    return payload


_by_key = {}
CACHE_MIN_BYTES = 1 << 20


def static_key(arr):
    """Content key of a static argument.  Arrays below 1 MB are hashed on every call (an
    in-place change is honoured, as with the reference's C call); for larger ones the key is
    cached per array object (address, shape, dtype): line lists and Voigt tables are read-only
    in the reference."""
    arr = np.ascontiguousarray(arr)
    ident = (id(arr), arr.ctypes.data, arr.shape, str(arr.dtype))
    key = _keys.get(ident) if arr.nbytes >= CACHE_MIN_BYTES else None
    if key is None:
        raw = memoryview(arr).cast("B")
        if arr.nbytes <= FULL_HASH_BYTES:
            digest = zlib.crc32(raw)
        else:   # sampled: head, tail and every 4096th 64-byte block
            digest = zlib.crc32(raw[:1 << 20])
            digest = zlib.crc32(raw[-(1 << 20):], digest)
            flat = np.frombuffer(raw, np.uint8)
            blocks = flat[:flat.size // 64 * 64].reshape(-1, 64)[::4096]
            digest = zlib.crc32(np.ascontiguousarray(blocks), digest)
        key = f"{arr.dtype}{arr.shape}:{arr.nbytes}:{digest:08x}"
        if arr.nbytes >= CACHE_MIN_BYTES:
            _keys[ident] = key
    _by_key[key] = arr          # kept until the request that names it has been served
    return key
