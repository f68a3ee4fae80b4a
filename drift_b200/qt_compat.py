"""QThread / pyqtSignal / QMutex when PyQt5 is installed (inside the DRIFT GUI), light stand-ins
otherwise (headless).  The reference's ``SimulationThread`` derives from ``QThread`` and exposes six
class-level ``pyqtSignal`` objects (lib/simulation_thread.py:19-26); the drop-in keeps that shape."""
from __future__ import annotations

import threading

try:  # pragma: no cover - PyQt5 is not part of the GPU image
    from PyQt5.QtCore import QMutex, QMutexLocker, QThread, pyqtSignal  # type: ignore

    HAVE_QT = True
except Exception:  # ImportError, or a broken Qt install
    HAVE_QT = False

    class _Bound:
        def __init__(self):
            self._slots = []

        def connect(self, fn):
            self._slots.append(fn)

        def disconnect(self, fn=None):
            if fn is None:
                self._slots = []
            else:
                self._slots = [s for s in self._slots if s != fn]

        def emit(self, *args):
            for fn in tuple(self._slots):
                fn(*args)

    class pyqtSignal:  # noqa: N801 - mirrors the Qt name
        def __init__(self, *arg_types):
            self._attr = None

        def __set_name__(self, owner, name):
            self._attr = "__signal_" + name

        def __get__(self, obj, owner=None):
            if obj is None:
                return self
            bound = obj.__dict__.get(self._attr)
            if bound is None:
                bound = _Bound()
                obj.__dict__[self._attr] = bound
            return bound

    class QThread:  # noqa: D401 - minimal threading-based equivalent
        def __init__(self, *a, **k):
            self._thread = None

        def start(self):
            self._thread = threading.Thread(target=self.run, daemon=True)
            self._thread.start()

        def run(self):
            pass

        def wait(self, msecs=None):
            if self._thread is not None:
                self._thread.join(None if msecs is None else msecs / 1000.0)
                return not self._thread.is_alive()
            return True

        def isRunning(self):  # noqa: N802
            return self._thread is not None and self._thread.is_alive()

        def msleep(self, ms):
            threading.Event().wait(ms / 1000.0)

        def deleteLater(self):  # noqa: N802
            pass

    class QMutex:
        def __init__(self):
            self._l = threading.Lock()

        def lock(self):
            self._l.acquire()

        def unlock(self):
            self._l.release()

    class QMutexLocker:
        def __init__(self, m):
            self._m = m

        def __enter__(self):
            self._m.lock()
            return self

        def __exit__(self, *a):
            self._m.unlock()
            return False
