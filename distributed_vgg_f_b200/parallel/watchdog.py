"""Failure propagation between ranks.

The reference has none: if one rank dies, the others sit in a gloo collective until the process-group
timeout (30 minutes) and the job never reports an error.  Here every rank runs a small daemon
thread that polls the rendezvous store (the TCPStore behind ``-iu``) once a second:

* a rank that fails sets ``b200/abort`` (with the reason) on its way out -> every other rank prints
  the reason and exits with status 75 within about a second, even if its main thread is blocked
  inside a collective;
* if the store itself disappears (rank 0 was killed), the survivors exit with status 76.

Device-side hangs are covered separately: every cross-GPU spin in csrc/allreduce.cu is bounded and
traps with a diagnostic (profiles/r1_watchdog_example.txt).
"""
from __future__ import annotations

import os
import sys
import threading
from typing import Optional

import torch.distributed as dist

ABORT_KEY = "b200/abort"
EXIT_PEER_FAILED = 75
EXIT_STORE_LOST = 76


def _default_store():
    try:
        from torch.distributed.distributed_c10d import _get_default_store

        return _get_default_store()
    except Exception:       # private API moved or no process group: run without the watchdog
        return None


class AbortWatch:
    def __init__(self, rank: int, interval: float = 1.0, store=None) -> None:
        self.rank, self.interval = rank, interval
        self.store = store if store is not None else _default_store()
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None

    def start(self) -> "AbortWatch":
        if self.store is not None and self._thread is None:
            self._thread = threading.Thread(target=self._run, name="b200-abort-watch", daemon=True)
            self._thread.start()
        return self

    def stop(self) -> None:
        self._stop.set()

    def signal(self, reason: str) -> None:
        """Tell the other ranks that this one is going down."""
        if self.store is None:
            return
        try:
            self.store.set(ABORT_KEY, "rank %d: %s" % (self.rank, reason[:500]))
        except Exception:
            pass

    def _run(self) -> None:
        misses = 0
        while not self._stop.wait(self.interval):
            try:
                hit = self.store.check([ABORT_KEY])
                misses = 0
            except Exception:
                misses += 1
                if misses >= 3 and not self._stop.is_set():
                    self._die(EXIT_STORE_LOST, "lost the rendezvous store (rank 0 gone?)")
                continue
            if hit and not self._stop.is_set():
                try:
                    reason = self.store.get(ABORT_KEY).decode("utf-8", "replace")
                except Exception:
                    reason = "unknown"
                self._die(EXIT_PEER_FAILED, "a peer failed -- " + reason)

    def _die(self, code: int, why: str) -> None:
        print("[Error] rank %d aborting: %s" % (self.rank, why), file=sys.stderr, flush=True)
        os._exit(code)          # the main thread may be blocked inside a collective: no clean unwinding


def start_abort_watch(rank: int) -> Optional[AbortWatch]:
    if not (dist.is_available() and dist.is_initialized()):
        return None
    if os.environ.get("B200_ABORT_WATCH", "1") == "0":
        return None
    return AbortWatch(rank).start()
