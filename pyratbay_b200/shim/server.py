"""GPU server of the shim (see __init__): owns the engines, serves the forked workers of the
reference over a Unix socket.  `python -m pyratbay_b200.shim.server <socket path> <owner pid>`.
"""
import os
import sys
import threading
import time
from multiprocessing.connection import Listener

import numpy as np

from . import client

_AUTH = b"pb200-shim"


class Server:
    def __init__(self, device=0):
        self.device = device
        self.arrays = {}            # content key -> numpy array
        self.engines = {}           # tuple of static keys (+ cutoff) -> Engine
        self.tables = {}            # etable key -> (device tensor, DeviceTable)
        self.lock = threading.Lock()
        self.stop = threading.Event()
        self.calls = 0

    # -- engines -----------------------------------------------------------------------
    def engine_for(self, static, cutoff):
        sig = tuple(sorted(static.items())) + (("cutoff", cutoff),)
        eng = self.engines.get(sig)
        if eng is None:
            from ..engine import Engine
            a = {name: self.arrays[key] for name, key in static.items()}
            eng = Engine(self.device)
            eng.set_grid(a["wn"], a["own"], a["divisors"])
            eng.set_species(a["molrad"], a["molmass"], a["isoimol"], a["isomass"], a["isoratio"])
            eng.set_lines(a["lwn"], a["elow"], a["gf"], a["lid"])
            eng.set_voigt(a["lorentz"], a["doppler"], a["psize"], a["pindex"], a["profile"], cutoff)
            self.engines[sig] = eng
            # the engine holds device copies: drop the large host copies (a later engine that
            # needs one asks the client again)
            for key in static.values():
                if self.arrays[key].nbytes > (16 << 20):
                    del self.arrays[key]
        return eng

    def extinction(self, static, unit):
        sig = tuple(sorted(static.items())) + (("cutoff", unit["cutoff"]),)
        if sig not in self.engines:
            missing = [key for key in static.values() if key not in self.arrays]
            if missing:
                return "need", missing
        eng = self.engine_for(static, unit["cutoff"])
        out = eng.extinction_batch(
            [unit["temp"]], unit["moldensity"][None, :], unit["isoz"][None, :], unit["isoiext"],
            unit["nextinct"], unit["ethresh"], unit["add"], unit["resolution"])
        return "bytes", np.ascontiguousarray(out[0])

    # -- Voigt grid --------------------------------------------------------------------
    def grid_sizes(self, state, lorentz, doppler, dwn, psize, profile_len):
        from .. import engine
        psize = np.ascontiguousarray(psize, np.int64).copy()
        index = np.zeros_like(psize)
        profile = np.zeros(int(profile_len), np.float64)
        engine.voigt_grid(profile, psize, index, lorentz, doppler, dwn, device=self.device)
        state["profile"] = profile
        # the caller will pass this very table back with every extinction call
        self.arrays[client.static_key(profile)] = profile
        return "ok", (psize, index)

    # -- table interpolation -----------------------------------------------------------
    def interp(self, per_mol, key, ttable, temps, density, lay1, lay2, ext):
        import torch
        if key not in self.arrays:
            return "need", [key]
        entry = self.tables.get(key)
        if entry is None:
            from ..engine import DeviceTable
            dev = torch.device("cuda", self.device)
            table = torch.from_numpy(self.arrays[key]).to(dev)
            entry = self.tables[key] = (table, DeviceTable(table, ttable, self.device))
        table, handle = entry
        out = torch.from_numpy(np.ascontiguousarray(ext)).to(table.device)
        handle.interp(temps, density, lay1, lay2, per_mol, out.data_ptr(), overwrite=False,
                      stream=torch.cuda.current_stream(table.device).cuda_stream, sync=True)
        return "bytes", out.cpu().numpy()

    # -- connection loop ---------------------------------------------------------------
    def serve(self, conn):
        state = {}
        try:
            while not self.stop.is_set():
                try:
                    msg = conn.recv()
                except (EOFError, ConnectionResetError):
                    return
                op, args = msg[0], msg[1:]
                try:
                    if op == "put":
                        key, dtype, shape = args
                        arr = np.empty(shape, dtype)
                        conn.recv_bytes_into(memoryview(arr).cast("B"))
                        self.arrays[key] = arr
                        reply = ("ok", None)
                    elif op == "shutdown":
                        conn.send(("ok", None))
                        self.stop.set()
                        return
                    else:
                        with self.lock:
                            self.calls += 1
                            if op == "extinction":
                                reply = self.extinction(*args)
                            elif op == "grid_sizes":
                                reply = self.grid_sizes(state, *args)
                            elif op == "grid_profile":
                                reply = ("bytes", state.pop("profile"))
                            elif op in ("interp_ec", "interp_ec_per_mol"):
                                reply = self.interp(op == "interp_ec_per_mol", *args)
                            elif op == "stats":
                                reply = ("ok", {"calls": self.calls, "engines": len(self.engines),
                                                "arrays": len(self.arrays)})
                            else:
                                reply = ("error", f"unknown request {op!r}")
                except Exception as exc:   # report to the caller, keep serving
                    reply = ("error", f"{type(exc).__name__}: {exc}")
                if reply[0] == "bytes":
                    arr = reply[1]
                    conn.send(("bytes", (str(arr.dtype), arr.shape)))
                    conn.send_bytes(memoryview(arr).cast("B"))
                else:
                    conn.send(reply)
        finally:
            conn.close()


def main():
    addr, owner = sys.argv[1], int(sys.argv[2])
    from .. import _lib
    _lib.load()
    _lib.require_device()          # fail loudly before anyone connects: no CPU fallback
    server = Server(int(os.environ.get("PB200_SHIM_DEVICE", "0")))
    listener = Listener(addr, family="AF_UNIX", authkey=_AUTH)

    def watchdog():                # leave when the process that started us is gone
        while not server.stop.is_set():
            try:
                os.kill(owner, 0)
            except OSError:
                break
            time.sleep(1.0)
        server.stop.set()
        try:                       # unblock accept()
            from multiprocessing.connection import Client
            Client(addr, family="AF_UNIX", authkey=_AUTH).close()
        except Exception:
            pass
    threading.Thread(target=watchdog, daemon=True).start()
    while not server.stop.is_set():
        try:
            conn = listener.accept()
        except Exception:
            continue
        if server.stop.is_set():
            break
        threading.Thread(target=server.serve, args=(conn,), daemon=True).start()
    listener.close()
    os._exit(0)


if __name__ == "__main__":
    main()
