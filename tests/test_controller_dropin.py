"""Drop-in proof at the controller level: the reference's UNMODIFIED ``lib/simulation_controller.py`` builds,
configures, wires and starts the drift_b200 ``SimulationThread`` once its module-level name is rebound
(the controller binds the class at import, lib/simulation_controller.py:18).

``reference``-marked: needs the reference tree (build container).  PyQt5 is replaced by the stand-in of
``oracle/ref_harness.py`` extended with the widgets the controller module imports.  With a CUDA device the run
goes all the way (all six signals fire); without one, ``initialize_simulation`` reports the missing device through
``log_message`` -- the reference's own error convention (lib/simulation_thread.py:281-283) -- and nothing else
differs."""
import sys
import types
from unittest import mock

import pytest

pytestmark = pytest.mark.reference


def _install_widgets_shim():
    from oracle import ref_harness

    ref_harness.install_pyqt_shim()
    core = sys.modules["PyQt5.QtCore"]

    class QTimer:
        def __init__(self, *a):
            self.timeout = ref_harness._BoundSignal()
            self.active = False
            self.interval = None

        def start(self, ms=None):
            self.active, self.interval = True, ms

        def stop(self):
            self.active = False

        def isActive(self):  # noqa: N802
            return self.active

    core.QTimer = getattr(core, "QTimer", QTimer)
    core.Qt = getattr(core, "Qt", mock.MagicMock())
    for name, attrs in (("PyQt5.QtGui", ("QIcon", "QPainter", "QPen", "QColor", "QFont", "QBrush")),
                        ("PyQt5.QtWidgets", ("QWidget", "QLabel", "QVBoxLayout"))):
        if name not in sys.modules:
            m = types.ModuleType(name)
            for a in attrs:
                setattr(m, a, type(a, (), {"__init__": lambda self, *a, **k: None}))
            sys.modules[name] = m
            setattr(sys.modules["PyQt5"], name.split(".")[1], m)


def test_reference_controller_drives_the_dropin_thread(tmp_path, monkeypatch):
    import random

    import numpy as np

    from oracle import ref_harness

    _install_widgets_shim()
    ref = ref_harness.import_reference()
    import importlib

    sc = importlib.import_module("lib.simulation_controller")
    from drift_b200.simulation_thread import SimulationThread as DropIn

    assert sc.SimulationThread is ref.simulation_thread.SimulationThread  # what the controller would use as shipped
    monkeypatch.setattr(sc, "SimulationThread", DropIn)                    # the one-line integration
    monkeypatch.chdir(tmp_path)
    random.seed(0)
    np.random.seed(0)
    import os

    G = ref_harness.load_graph(os.path.join(ref_harness.REFERENCE_ROOT, "example-data", "graphml", "small",
                                            "g.40.16.graphml"), seed=0)
    assert G is not None and len(G) == 40

    win = mock.MagicMock()
    win.graph = G
    win.agents_spinbox.value.return_value = 60
    win.duration_spinbox.value.return_value = 1
    win.speed_spinbox.value.return_value = 100          # tick = 0.1 s x 100 = 10 s: 360 ticks
    win.selection_combo.currentText.return_value = "random"
    win.simulation_widget.st_selector = None
    win.get_current_settings.return_value = {"bpr_alpha": 0.2, "bpr_beta": 3.0}
    logs, fired = [], {k: 0 for k in ("agents", "trip", "status", "finished", "selector")}
    win.add_log_message.side_effect = logs.append
    win.update_agents.side_effect = lambda *_: fired.__setitem__("agents", fired["agents"] + 1)
    win.add_completed_trip.side_effect = lambda *_: fired.__setitem__("trip", fired["trip"] + 1)
    win.update_status.side_effect = lambda *_: fired.__setitem__("status", fired["status"] + 1)
    win.simulation_finished.side_effect = lambda *_: fired.__setitem__("finished", fired["finished"] + 1)
    win.handle_st_selector_ready.side_effect = lambda *_: fired.__setitem__("selector", fired["selector"] + 1)

    ctl = sc.SimulationController(win)
    with ref_harness._quiet():
        ctl.start_simulation()   # :42-80 -> _start_simulation_with_parameters :172-196 -> _setup_and_start_simulation :256-285
    th = ctl.simulation_thread
    assert isinstance(th, DropIn)
    assert th.time_acceleration == 100 and th.num_agents == 60 and th.selection_mode == "random"
    assert th.settings.get("bpr_alpha") == 0.2 and th.settings.get("bpr_beta") == 3.0   # apply_settings reached it
    assert ctl.simulation_timer.isActive()
    # the six signals carry the controller's slots (lib/simulation_controller.py:266-271)
    for sig, slot in ((th.log_message, win.add_log_message), (th.agents_updated, win.update_agents),
                      (th.trip_completed, win.add_completed_trip), (th.status_updated, win.update_status),
                      (th.simulation_finished, win.simulation_finished),
                      (th.st_selector_ready, win.handle_st_selector_ready)):
        slots = getattr(sig, "_slots", None)
        assert slots is not None and slot in slots
    th.wait()
    try:
        import torch

        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:  # the whole run: every signal fired, trips were recorded, the CSV exists
        assert fired["selector"] == 1 and fired["finished"] == 1 and fired["agents"] > 0 and fired["status"] > 0
        assert fired["trip"] > 0 and th.data_logger.trips_data
    else:         # no device: reported the reference's way, through log_message, and the thread ended cleanly
        assert any("CUDA" in m or "cuda" in m for m in logs), logs
        assert fired["finished"] == 0 and not th.isRunning()
    # the controller's stop path works on the drop-in too (:287-345: stop(), wait(), deleteLater())
    with ref_harness._quiet():
        ctl.stop_simulation()
    assert ctl.simulation_thread is None
