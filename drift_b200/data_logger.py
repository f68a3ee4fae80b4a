"""Trip records: CSV / JSON output identical to the reference's ``SimulationDataLogger``
(lib/data_logger.py:8-306): same ten CSV columns, same ids (``T%06d`` assigned when the trip
starts), ``HH:MM:SS`` stamps counted from the simulation start, ``route_taken`` as ``str(list)``,
distances from the minimum ``length`` per hop, and the same summary statistics.  Records are held
as a list of dicts (``trips_data``) because the reference's ExportManager and signal payloads
read that list directly (lib/managers/export_manager.py:42-84)."""
from __future__ import annotations

import ast
import csv
import datetime
import json
import math
import os
import sys

FIELDNAMES = ["trip_id", "agent_id", "start_time", "end_time", "origin_node", "destination_node",
              "route_taken", "trip_distance_km", "trip_duration_min", "average_speed_kmh"]


class SimulationDataLogger:
    def __init__(self, output_dir="data", simulation_start_time=None):
        self.output_dir = output_dir
        self.trips_data = []
        self.trip_counter = 0
        os.makedirs(output_dir, exist_ok=True)
        self.simulation_start_time = simulation_start_time or datetime.datetime.now()

    # ---- record lifecycle -------------------------------------------------------------------
    def _get_timestamp(self, simulation_time):
        return (self.simulation_start_time + datetime.timedelta(seconds=simulation_time)).strftime("%H:%M:%S")

    def start_trip(self, agent, simulation_time):
        self.trip_counter += 1
        path = getattr(agent, "path", None)
        rec = {
            "trip_id": "T%06d" % self.trip_counter,
            "agent_id": "A%06d" % id(agent),
            "start_time": self._get_timestamp(simulation_time),
            "end_time": None,
            "origin_node": agent.source,
            "destination_node": agent.target,
            "route_taken": [agent.source],
            "trip_distance_km": 0.0,
            "trip_duration_min": 0.0,
            "average_speed_kmh": 0.0,
            "simulation_start_time": simulation_time,
            "path_nodes": list(path) if path is not None and hasattr(agent, "path") else [],
        }
        agent.current_trip_data = rec
        return rec

    def update_trip(self, agent, simulation_time, current_node=None):
        rec = getattr(agent, "current_trip_data", None)
        if rec is None:
            return
        if current_node and current_node not in rec["route_taken"]:
            rec["route_taken"].append(current_node)

    def end_trip(self, agent, simulation_time, traffic_manager=None, path_length_m=None):
        """``path_length_m`` lets the engine pass the hop-length sum it already knows; otherwise it is
        taken from ``agent.graph`` like the reference does (lib/data_logger.py:103-120)."""
        rec = getattr(agent, "current_trip_data", None)
        if rec is None:
            return None
        minutes = (simulation_time - rec["simulation_start_time"]) / 60.0
        rec["trip_duration_min"] = minutes
        rec["end_time"] = self._get_timestamp(simulation_time)
        if hasattr(agent, "path") and len(agent.path) > 1:  # TypeError on path None: reference quirk kept
            metres = path_length_m if path_length_m is not None else self._calculate_path_distance(agent.path, agent.graph)
            rec["trip_distance_km"] = metres / 1000.0
        else:
            rec["trip_distance_km"] = 0.0
        rec["average_speed_kmh"] = (rec["trip_distance_km"] / minutes) * 60 if minutes > 0 else 0.0
        if agent.target not in rec["route_taken"]:
            rec["route_taken"].append(agent.target)
        out = dict(rec)
        out["route_taken"] = str(out["route_taken"])
        del out["simulation_start_time"]
        del out["path_nodes"]
        self.trips_data.append(out)
        agent.current_trip_data = None
        return out

    def _calculate_path_distance(self, path, graph):
        total = 0.0
        for u, v in zip(path[:-1], path[1:]):
            try:
                parallel = graph.get_edge_data(u, v)
                if parallel:
                    best = min(parallel.values(), key=lambda a: a.get("length", float("inf")))
                    total += best.get("length", 1000.0)
            except Exception:
                total += 1000.0
        return total

    # ---- output -------------------------------------------------------------------------------
    def _write_csv(self, filepath):
        with open(filepath, "w", newline="", encoding="utf-8") as fh:
            w = csv.DictWriter(fh, fieldnames=FIELDNAMES)
            w.writeheader()
            for trip in self.trips_data:
                w.writerow({k: v for k, v in trip.items() if k in FIELDNAMES})

    def save_to_csv(self, filename=None):
        if not filename:
            filename = "simulation_trips_%s.csv" % datetime.datetime.now().strftime("%Y%m%d_%H%M%S")
        filepath = os.path.join(self.output_dir, filename)
        if not self.trips_data:
            print("No trip data to save.")
            return filepath
        self._write_csv(filepath)
        print("Trip data saved to: %s" % filepath)
        print("Total trips recorded: %d" % len(self.trips_data))
        return filepath

    def export_to_csv(self, filepath):
        if not self.trips_data:
            print("No trip data to export.")
            return filepath
        self._write_csv(filepath)
        return filepath

    def _json_payload(self):
        trips = []
        for trip in self.trips_data:
            route = trip.get("route_taken", [])
            if isinstance(route, str):
                try:
                    route = ast.literal_eval(route)
                except Exception:
                    route = []
            trips.append({k: (route if k == "route_taken" else trip.get(k, 0.0 if k.startswith(("trip_d", "average")) else None))
                          for k in FIELDNAMES})
        return {
            "metadata": {
                "export_timestamp": datetime.datetime.now().isoformat(),
                "simulation_start_time": self.simulation_start_time.isoformat(),
                "total_trips": len(self.trips_data),
            },
            "trips": trips,
        }

    def save_to_json(self, filename=None):
        if not filename:
            filename = "simulation_trips_%s.json" % datetime.datetime.now().strftime("%Y%m%d_%H%M%S")
        filepath = os.path.join(self.output_dir, filename)
        if not self.trips_data:
            print("No trip data to save.")
            return filepath
        return self.export_to_json(filepath)

    def export_to_json(self, filepath):
        if not self.trips_data:
            print("No trip data to export.")
            return filepath
        with open(filepath, "w", encoding="utf-8") as fh:
            json.dump(self._json_payload(), fh, indent=2, ensure_ascii=False)
        return filepath

    def export_to_format(self, filepath, format_type="csv"):
        kind = format_type.lower()
        if kind == "csv":
            return self.export_to_csv(filepath)
        if kind == "json":
            return self.export_to_json(filepath)
        raise ValueError("Unsupported format: %s" % format_type)

    def get_statistics(self):
        if not self.trips_data:
            return {}
        n = len(self.trips_data)
        km = sum(t["trip_distance_km"] for t in self.trips_data)
        minutes = sum(t["trip_duration_min"] for t in self.trips_data)
        speed = sum(t["average_speed_kmh"] for t in self.trips_data) / n
        return {
            "total_trips": n,
            "total_distance_km": km,
            "total_duration_hours": minutes / 60,
            "average_speed_kmh": speed,
            "average_trip_distance_km": km / n,
            "average_trip_duration_min": minutes / n,
        }


class _RunningSum:
    """Left-to-right float sum that ends on the value the builtin ``sum`` returns for the same sequence: plain
    additions before Python 3.12, Neumaier-compensated from 3.12 on (bltinmodule.c: the compensation is added once, at
    the end, if it is finite and non-zero)."""

    def __init__(self):
        self.total = 0.0
        self.comp = 0.0
        self.compensated = sys.version_info >= (3, 12)

    def add(self, x):
        x = float(x)
        if not self.compensated:
            self.total += x
            return
        t = self.total + x
        if abs(self.total) >= abs(x):
            self.comp += (self.total - t) + x
        else:
            self.comp += (x - t) + self.total
        self.total = t

    def value(self):
        if self.compensated and self.comp and math.isfinite(self.comp):
            return self.total + self.comp
        return self.total


class TripStatistics:
    """``SimulationDataLogger.get_statistics`` (lib/data_logger.py:154-171) for trips that are streamed out instead
    of kept: the same six figures from running sums, equal to the list-based ones bit for bit."""

    def __init__(self):
        self.n = 0
        self._km, self._min, self._kmh = _RunningSum(), _RunningSum(), _RunningSum()

    def add(self, distance_km, duration_min, speed_kmh):
        self.n += 1
        self._km.add(distance_km)
        self._min.add(duration_min)
        self._kmh.add(speed_kmh)

    def result(self):
        if not self.n:
            return {}
        km, minutes = self._km.value(), self._min.value()
        return {
            "total_trips": self.n,
            "total_distance_km": km,
            "total_duration_hours": minutes / 60,
            "average_speed_kmh": self._kmh.value() / self.n,
            "average_trip_distance_km": km / self.n,
            "average_trip_duration_min": minutes / self.n,
        }


class TripJsonStream:
    """The JSON export of lib/data_logger.py:260-306, written trip by trip for runs whose trips do not fit in host
    memory as a list of dicts (``block_runner``): same two keys, same per-trip fields, same conversions
    (``route_taken`` parsed back from its string form, 0.0 for missing numbers).  ``total_trips`` is only known at
    the end, so "metadata" follows "trips" in the file -- a JSON object has no order; ``json.load`` returns the same
    dict as for the reference's file."""

    def __init__(self, filepath, simulation_start_time=None):
        self.filepath = filepath
        self.simulation_start_time = simulation_start_time or datetime.datetime.now().replace(
            hour=6, minute=0, second=0, microsecond=0)
        self.total = 0
        self._fh = open(filepath, "w", encoding="utf-8")
        self._fh.write('{\n  "trips": [')

    @staticmethod
    def _trip(trip):
        route = trip.get("route_taken", [])
        if isinstance(route, str):
            try:
                route = ast.literal_eval(route)
            except Exception:
                route = []
        return {k: (route if k == "route_taken" else trip.get(k, 0.0 if k.startswith(("trip_d", "average")) else None))
                for k in FIELDNAMES}

    def write(self, rows):
        for trip in rows:
            text = json.dumps(self._trip(trip), indent=2, ensure_ascii=False)
            self._fh.write(("," if self.total else "") + "\n    " + text.replace("\n", "\n    "))
            self.total += 1

    def close(self):
        if self._fh is None:
            return self.filepath
        meta = {
            "export_timestamp": datetime.datetime.now().isoformat(),
            "simulation_start_time": self.simulation_start_time.isoformat(),
            "total_trips": self.total,
        }
        self._fh.write(("\n  " if self.total else "") + '],\n  "metadata": '
                       + json.dumps(meta, indent=2, ensure_ascii=False).replace("\n", "\n  ") + "\n}\n")
        self._fh.close()
        self._fh = None
        return self.filepath
