"""ctypes binding of the C ABI declared in ``include/drift_b200.h``.

The shared library is built in-tree by ``drift_b200.build`` (``libdrift_b200.so``).  There is
no CPU implementation behind it: without the library, or without a CUDA device, every call
fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdrift_b200.so")

DRIFT_OK = 0
ROUTE_OK, ROUTE_TRIVIAL, ROUTE_NOPATH = 0, 1, 2
ROUTER_NX_EXACT, ROUTER_BATCHED = 0, 1

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)


class GraphDesc(C.Structure):
    _fields_ = [
        ("V", C.c_int32), ("L", C.c_int32), ("Ef", C.c_int32), ("Eb", C.c_int32),
        ("fwd_rowptr", c_i32p), ("fwd_col", c_i32p), ("fwd_w", c_f64p), ("fwd_link", c_i32p),
        ("bwd_rowptr", c_i32p), ("bwd_col", c_i32p), ("bwd_w", c_f64p), ("bwd_link", c_i32p),
        ("link_len", c_f64p), ("link_v0", c_f64p), ("link_t0", c_f64p), ("link_cap", c_i32p),
        ("link_u", c_i32p), ("link_v", c_i32p),
    ]


class StatsOut(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "moving_agents", "total_agents_on_roads", "total_capacity", "links_with_traffic",
        "arrivals_total", "departures_total", "route_arena_used", "events_last_tick", "arena_compactions", "fallback_routes",
        "tie_fallbacks")]


class HierDesc(C.Structure):
    _fields_ = [
        ("V", C.c_int32), ("max_level", C.c_int32), ("max_label", C.c_int32), ("pad0", C.c_int32),
        ("mean_label", C.c_double), ("n_up", C.c_int64), ("n_down", C.c_int64),
        ("level", c_i32p),
        ("up_rowptr", c_i32p), ("up_other", c_i32p), ("up_w", c_f64p), ("up_a", c_i32p), ("up_b", c_i32p),
        ("up_hops", c_i32p),
        ("down_rowptr", c_i32p), ("down_other", c_i32p), ("down_w", c_f64p), ("down_a", c_i32p), ("down_b", c_i32p),
        ("down_hops", c_i32p),
    ]


HIER = C.c_void_p
# name -> (restype, argtypes); must list every symbol of include/drift_b200.h
ENGINE = C.c_void_p
SIGNATURES = {
    "drift_last_error": (C.c_char_p, []),
    "drift_abi_version": (C.c_int, []),
    "drift_create": (C.c_int, [C.POINTER(ENGINE), C.c_int, C.POINTER(GraphDesc)]),
    "drift_destroy": (C.c_int, [ENGINE]),
    "drift_sync": (C.c_int, [ENGINE]),
    "drift_set_bpr": (C.c_int, [ENGINE, C.c_double, C.c_double, c_f64p, c_i64p, C.c_int32, c_i32p]),
    "drift_add_agents": (C.c_int, [ENGINE, C.c_int64, C.c_int64, C.c_int64]),
    "drift_route_od": (C.c_int, [ENGINE, C.c_int32, C.c_int64, c_i32p, c_i32p, C.c_int32, c_i32p, c_i32p]),
    "drift_fetch_routes": (C.c_int, [ENGINE, c_i32p, C.c_int64]),
    "drift_fetch_route_costs": (C.c_int, [ENGINE, c_f64p, C.c_int64]),
    "drift_commit_departures": (C.c_int, [ENGINE, C.c_int64, c_i32p]),
    "drift_submit_departures": (C.c_int, [ENGINE, C.c_int64, c_i32p, c_i64p, c_i32p]),
    "drift_step": (C.c_int, [ENGINE, C.c_int32, C.c_double]),
    "drift_set_demand": (C.c_int, [ENGINE, C.c_uint64, c_f64p, c_i32p, C.c_int32, C.c_int32]),
    "drift_push_trips": (C.c_int, [ENGINE, c_i32p, C.c_int32]),
    "drift_store_trip_pool": (C.c_int, [ENGINE, c_i32p, C.c_int32, C.c_int32]),
    "drift_refill_from_pool": (C.c_int, [ENGINE, C.c_int32]),
    "drift_run": (C.c_int, [ENGINE, C.c_int32, C.c_double, C.c_double]),
    "drift_update_travel_times": (C.c_int, [ENGINE]),
    "drift_update_travel_times_with": (C.c_int, [ENGINE, C.c_void_p]),
    "drift_reroute_moving": (C.c_int, [ENGINE, C.c_int32, c_i64p]),
    "drift_invalidate_prefetch": (C.c_int, [ENGINE, C.c_double, C.c_double, C.c_int32]),
    "drift_read_occupancy": (C.c_int, [ENGINE, c_i32p]),
    "drift_recount_occupancy": (C.c_int, [ENGINE, c_i32p]),
    "drift_recount_device": (C.c_int, [ENGINE]),
    "drift_read_travel_times": (C.c_int, [ENGINE, c_f64p]),
    "drift_read_agents": (C.c_int, [ENGINE, c_u8p, c_f64p, c_f64p, c_f64p, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p]),
    "drift_read_agent_route": (C.c_int, [ENGINE, C.c_int64, c_i32p, C.c_int64, c_i64p]),
    "drift_export_routes": (C.c_int, [ENGINE, C.c_int64, c_i32p, c_i32p, c_f64p, c_i32p, C.c_int64]),
    "drift_drain_arrivals": (C.c_int, [ENGINE, c_i32p, c_i32p, C.c_int64, c_i64p]),
    "drift_drain_departures": (C.c_int, [ENGINE, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, C.c_int64, c_i64p]),
    "drift_drain_departures_routes": (C.c_int, [ENGINE, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, c_i64p, c_i32p, c_i32p,
                                                C.c_int64, c_i64p]),
    "drift_export_logged_routes": (C.c_int, [ENGINE, C.c_int64, c_i64p, c_i32p, c_i32p, c_f64p, c_i32p, C.c_int64]),
    "drift_enable_entry_trace": (C.c_int, [ENGINE, C.c_int64]),
    "drift_drain_entries": (C.c_int, [ENGINE, c_i32p, c_i32p, c_i32p, c_f64p, C.c_int64, c_i64p]),
    "drift_stats": (C.c_int, [ENGINE, C.POINTER(StatsOut)]),
    "drift_tick_index": (C.c_int, [ENGINE, c_i64p]),
    "drift_tick_begin": (C.c_int, [ENGINE, C.c_double, c_i64p]),
    "drift_tick_begin_demand": (C.c_int, [ENGINE, C.c_double, C.c_double, c_i64p]),
    "drift_events_ptr": (C.c_int, [ENGINE, C.POINTER(C.c_void_p), c_i64p]),
    "drift_tick_finish": (C.c_int, [ENGINE, C.c_void_p, C.c_int64]),
    "drift_set_stream": (C.c_int, [ENGINE, C.c_void_p]),
    "drift_tick_begin_async": (C.c_int, [ENGINE, C.c_double, C.c_double, C.c_void_p, C.c_int32]),
    "drift_tick_finish_gathered": (C.c_int, [ENGINE, C.c_void_p, C.c_int32, C.c_int32]),
    "drift_max_events": (C.c_int, [ENGINE, c_i64p, C.c_int32]),
    "drift_p2p_export": (C.c_int, [ENGINE, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_void_p]),
    "drift_p2p_connect": (C.c_int, [ENGINE, C.c_void_p]),
    "drift_kernel_times": (C.c_int, [ENGINE, c_f64p, c_i64p, c_f64p, c_i64p]),
    "drift_set_profiling": (C.c_int, [ENGINE, C.c_int32]),
    "drift_phase_times": (C.c_int, [ENGINE, c_f64p, c_f64p, c_f64p, c_f64p, c_i64p, C.c_int32]),
    "drift_p2p_phase_detail": (C.c_int, [ENGINE, c_f64p, C.c_int32]),
    "drift_tree_stats": (C.c_int, [ENGINE, c_f64p, c_i64p, c_i64p, c_i64p, c_i64p, c_i64p, c_i64p, c_i64p, C.c_int32]),
    "drift_time_passes": (C.c_int, [ENGINE, C.c_int32, c_f64p, c_f64p]),
    "drift_set_landmarks": (C.c_int, [ENGINE, C.c_int32, c_f64p, c_f64p]),
    "drift_hier_build": (C.c_int, [C.c_int32, C.c_int32, c_i32p, c_i32p, c_f64p, c_i32p, C.c_int32, C.c_double,
                                   C.POINTER(HIER)]),
    "drift_hier_info": (C.c_int, [HIER, c_i64p, c_i64p, c_i64p, c_i32p, c_i32p, c_f64p, c_f64p]),
    "drift_hier_export": (C.c_int, [HIER, C.POINTER(HierDesc)]),
    "drift_hier_free": (C.c_int, [HIER]),
    "drift_set_hierarchy": (C.c_int, [ENGINE, C.POINTER(HierDesc)]),
    "drift_label_stats": (C.c_int, [ENGINE, c_i64p, c_f64p, c_f64p, c_i64p, c_i64p, c_i64p, C.c_int32]),
    "drift_host_pow": (C.c_int, [c_f64p, C.c_double, c_f64p, C.c_int64]),
    "drift_set_option": (C.c_int, [ENGINE, C.c_char_p, C.c_double]),
    "drift_device_ptr": (C.c_int, [ENGINE, C.c_char_p, C.POINTER(C.c_void_p), c_i64p]),
}

_lib = None


class DriftError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("drift_b200 error %d: %s" % (code, msg))
        self.code = code


def load():
    """Load ``libdrift_b200.so`` and declare every prototype.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "drift_b200: %s is missing -- build it with `python -m drift_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != DRIFT_OK:
        msg = load().drift_last_error()
        raise DriftError(rc, msg.decode() if msg else "?")
