"""Project / Experiment / RunComputation / Computation: the user-facing shell of MDSuite that
the hot-path calculators hang off, with the same names, call signatures and result shapes.

Mirrors (API only; SQLAlchemy and h5py are not available, so storage is stdlib ``sqlite3`` +
JSON for results/metadata and ``TrajectoryStore`` for arrays):

  mdsuite/project/project.py:45-338          Project, add_experiment, run, experiments
  mdsuite/experiment/experiment.py:89-639    Experiment (metadata, add_data, species, run)
  mdsuite/experiment/run.py:58-242           RunComputation (one property per calculator)
  mdsuite/database/scheme.py:193-342         Computation (data_dict, [], keys, parameters)
  mdsuite/database/calculator_database.py:91-248  cache query / store protocol
"""
from __future__ import annotations

import json
import logging
import os
import sqlite3
from collections import OrderedDict
from dataclasses import dataclass, fields, is_dataclass
from datetime import datetime
from typing import Dict, List, Optional, Union

import numpy as np

from . import distributed as D
from .file_io import (BlockInput, LAMMPSFluxFile, LAMMPSTrajectoryFile, ScriptInput,
                      TrajectoryMetadata)
from .store import TrajectoryStore, join_path
from .units import Units, resolve_units

log = logging.getLogger("mdsuite_b200")


# ------------------------------------------------------------------------------------------
# results
# ------------------------------------------------------------------------------------------
def conv_to_db(val):
    """Serialise a calculator argument the way database/scheme.py / calculator_database.py
    do for the cache comparison: JSON-able values stay, everything else becomes ``str``."""
    if isinstance(val, (str, int, float, bool)) or val is None:
        return val
    if isinstance(val, (np.integer,)):
        return int(val)
    if isinstance(val, (np.floating,)):
        return float(val)
    if isinstance(val, (list, tuple)):
        return [conv_to_db(v) for v in val]
    if isinstance(val, np.ndarray):
        return val.tolist()
    if isinstance(val, dict):
        return {str(k): conv_to_db(v) for k, v in val.items()}
    return str(val)  # slices -> "slice(None, None, None)"


class Computation:
    """Result of one calculator run (database/scheme.py:193-342)."""

    def __init__(self, name: str, experiment_name: str, parameters: dict,
                 results: "OrderedDict[str, dict]", comp_id: Optional[int] = None,
                 metadata: Optional[dict] = None):
        self.name = name
        self.experiment_name = experiment_name
        self._parameters = parameters
        self._results = results
        self.id = comp_id
        # run diagnostics that are not part of the reference's data_dict (e.g. the RDF's
        # bin-edge tie report); data_dict keeps exactly the reference's keys
        self.metadata = dict(metadata or {})

    def __repr__(self):
        return f"Exp{self.experiment_name}_{self.name}_{self.id}"

    @property
    def data_dict(self) -> dict:
        return dict(self._results)

    def __getitem__(self, item):
        try:
            return self._results[item]
        except KeyError:
            raise KeyError(f"Could not find {item} - available keys are {self._results.keys()}")

    def keys(self) -> list:
        return list(self._results.keys())

    @property
    def computation_parameter(self) -> dict:
        return dict(self._parameters)

    @property
    def data_range(self) -> int:
        if "data_range" in self._parameters:
            return int(self._parameters["data_range"])


def subjects_key(subjects: List[str]) -> str:
    """scheme.py:225-268: species joined by '_', empty -> 'System'."""
    key = "_".join(s for s in subjects if s is not None)
    return key if key and key != "System" else "System"


# ------------------------------------------------------------------------------------------
# species
# ------------------------------------------------------------------------------------------
@dataclass
class Species:
    name: str
    n_particles: int
    charge: float = 0.0
    mass: Optional[float] = None


# ------------------------------------------------------------------------------------------
# Experiment
# ------------------------------------------------------------------------------------------
class Experiment:
    def __init__(self, project: "Project", name: str, time_step: float = None,
                 temperature: float = None, units: Union[str, Units] = None,
                 cluster_mode: bool = None):
        self.project = project
        self.name = name
        self.time_step = time_step
        self.temperature = temperature
        # experiment.py:188-191: units given here win, stored units come next, REAL is the
        # default only when nothing is stored (_load_metadata)
        self._units_given = units is not None
        self.units = resolve_units(units if units is not None else "real")
        self.cluster_mode = cluster_mode
        self.active = True
        self.experiment_path = os.path.join(project.storage_path, project.name, name)
        self.database_path = os.path.join(self.experiment_path, "database")
        if project.is_writer:
            os.makedirs(self.database_path, exist_ok=True)
        project.barrier()
        self.store = TrajectoryStore(self.database_path if project.persist else None,
                                     sharded=project.sharded)
        self.species: "OrderedDict[str, Species]" = OrderedDict()
        self.molecules: dict = {}
        self.box_array: Optional[list] = None
        self.number_of_configurations = 0
        self.number_of_atoms = 0
        self.sample_rate = 1
        self.version = 0
        self.read_files: List[str] = []
        self.property_groups: Dict[str, list] = {}
        self._load_metadata()

    # -- metadata persistence (experiment_database.py:46-433, JSON instead of SQL rows) ------
    def _meta_path(self):
        return os.path.join(self.experiment_path, "experiment.json")

    def _load_metadata(self):
        if self.project.persist and os.path.exists(self._meta_path()):
            with open(self._meta_path()) as fh:
                m = json.load(fh)
            self.time_step = m["time_step"] if self.time_step is None else self.time_step
            self.temperature = m["temperature"] if self.temperature is None else self.temperature
            self.box_array = m["box_array"]
            self.number_of_configurations = m["number_of_configurations"]
            self.number_of_atoms = m["number_of_atoms"]
            self.sample_rate = m["sample_rate"]
            self.version = m["version"]
            self.read_files = m["read_files"]
            self.species = OrderedDict((k, Species(**v)) for k, v in m["species"])
            if not self._units_given and m.get("units") is not None:
                self.units = Units(**m["units"])

    def _save_metadata(self):
        if not self.project.persist or not self.project.is_writer:
            return
        m = dict(time_step=self.time_step, temperature=self.temperature,
                 box_array=self.box_array, number_of_configurations=self.number_of_configurations,
                 number_of_atoms=self.number_of_atoms, sample_rate=self.sample_rate,
                 version=self.version, read_files=self.read_files,
                 units={f.name: getattr(self.units, f.name) for f in fields(self.units)},
                 species=[(k, v.__dict__) for k, v in self.species.items()])
        with open(self._meta_path(), "w") as fh:
            json.dump(m, fh)

    @property
    def volume(self) -> float:
        """experiment_database.py:430-433."""
        return float(np.prod(np.asarray(self.box_array, dtype=float)))

    @property
    def run(self) -> "RunComputation":
        return RunComputation(experiment=self)

    # -- ingest (experiment.py:459-552) ------------------------------------------------------------
    def add_data(self, simulation_data, force: bool = False, update_with_pubchempy: bool = False):
        items = simulation_data if isinstance(simulation_data, list) else [simulation_data]
        for item in items:
            proc = _get_processor(item)
            tag = getattr(proc, "file_path", None) or getattr(proc, "name", repr(proc))
            if str(tag) in self.read_files and not force:
                log.info("This file has already been read, skipping this now.")
                continue
            self._add_from_processor(proc)
            self.read_files.append(str(tag))
        self._save_metadata()
        self.project.barrier()     # shared files: every rank's rows have landed

    def _add_from_processor(self, proc):
        meta: TrajectoryMetadata = proc.metadata
        if all(sp.name == "Observables" for sp in meta.species_list):
            return self._add_observables(proc, meta)
        offset = self.number_of_configurations
        if offset == 0:
            self.box_array = [float(b) for b in meta.box_l]
            self.sample_rate = int(meta.sample_rate)
            for sp in meta.species_list:
                self.species[sp.name] = Species(sp.name, int(sp.n_particles),
                                                charge=float(sp.charge or 0), mass=sp.mass)
            self.number_of_atoms = sum(s.n_particles for s in self.species.values())
        total = offset + meta.n_configurations
        for sp in meta.species_list:
            for prop in sp.properties:
                path = join_path(sp.name, prop.name)
                if not self.store.check_existence(path):
                    self.store.add_dataset(path, (sp.n_particles, total, prop.n_dims))
                else:
                    self.store.resize_dataset(path, total)
        if isinstance(proc, BlockInput):
            for sp, prop, rows, arr in proc.blocks():
                self.store.add_data(join_path(sp, prop), arr, start=offset, rows=rows)
        elif isinstance(proc, ScriptInput):
            for sp, prop, arr, rows in proc.arrays():
                self.store.add_data(join_path(sp, prop), arr, start=offset, rows=rows)
        else:
            pos = offset
            for chunk in proc.get_configurations_generator():
                for sp, props in chunk.data.items():
                    for prop, arr in props.items():
                        self.store.add_data(join_path(sp, prop), arr, start=pos)
                pos += chunk.chunk_size
        self.number_of_configurations = total
        self.version += 1  # new data invalidates cached computations (calculator_database.py:139-150)

    def _add_observables(self, proc, meta: TrajectoryMetadata):
        """System observables from a flux / log table (lammps_flux_files.py): stored as
        ``Observables/<Property>`` (1, n_steps, n_dims).  Unlike upstream, "Observables" is not
        registered as a particle species.  An experiment that has no trajectory yet takes its
        length, box and sample rate from the table."""
        n = meta.n_configurations
        if self.number_of_configurations == 0:
            self.box_array = [float(b) for b in meta.box_l]
            self.sample_rate = int(meta.sample_rate)
            self.number_of_configurations = n
        pos = 0
        for prop in meta.species_list[0].properties:
            path = join_path("Observables", prop.name)
            if self.store.check_existence(path):
                self.store.remove(path)
            self.store.add_dataset(path, (1, n, prop.n_dims))
        for chunk in proc.get_configurations_generator():
            for prop, arr in chunk.data["Observables"].items():
                self.store.add_data(join_path("Observables", prop), arr, start=pos)
            pos += chunk.chunk_size
        self.version += 1

    # -- transformations (experiment.py:270-282) -----------------------------------------------------
    def cls_transformation_run(self, transformation, *args, **kwargs):
        transformation.experiment = self
        transformation.run_transformation(*args, **kwargs)

    def load_matrix(self, property_name: str = None, species: list = None,
                    select_slice=np.s_[:], path: list = None):
        """experiment.py:554-597: host float64 arrays (dict path -> array, or one array)."""
        if path is None:
            species = list(self.species) if species is None else species
            path = [join_path(s, property_name) for s in species]
        out = {p: self.store.load_data(p, select_slice) for p in path}
        return out[path[0]] if len(path) == 1 else out


def _get_processor(item):
    """experiment.py:62-86."""
    if isinstance(item, (str, os.PathLike)):
        p = str(item)
        if p.endswith(".lammpstraj"):
            return LAMMPSTrajectoryFile(p)
        raise ValueError(f"datafile ending '{os.path.splitext(p)[1]}' not recognized; "
                         "instantiate a file reader (LAMMPSTrajectoryFile, ScriptInput)")
    if hasattr(item, "metadata"):
        return item
    raise ValueError(f"simulation_data entry {item!r} is neither a path nor a file processor")


# ------------------------------------------------------------------------------------------
# Project
# ------------------------------------------------------------------------------------------
class Project:
    def __init__(self, name: str = None, storage_path: str = "./", persist: bool = True,
                 sharded: bool = None):
        """``sharded`` (default: True when a torch.distributed group with more than one rank is
        active): all ranks open the SAME project; the trajectory store is atom-sharded
        (store.py), a persistent project directory is shared and written by rank 0 only, and
        cache decisions are broadcast.  ``sharded=False`` gives every rank a private, complete
        project (bench.py's weak-scaling replicas)."""
        self.name = f"MDSuite_Project_{name}" if name is not None else "MDSuite_Project"
        self.storage_path = str(storage_path)
        self.persist = persist
        self.sharded = (D.world_size() > 1) if sharded is None else bool(sharded)
        self._shared_files = self.sharded and persist   # one directory for all ranks
        if self.is_writer:
            os.makedirs(os.path.join(self.storage_path, self.name), exist_ok=True)
        self.barrier()
        self._experiments: "OrderedDict[str, Experiment]" = OrderedDict()
        self._db = None
        names = []
        if self.is_writer:
            db_path = os.path.join(self.storage_path, self.name, "project.db") if persist \
                else ":memory:"
            self._db = sqlite3.connect(db_path)
            self._db.execute(
                "CREATE TABLE IF NOT EXISTS computations (id INTEGER PRIMARY KEY, name TEXT, "
                "experiment TEXT, parameters TEXT, results BLOB, metadata TEXT)")
            self._db.execute("CREATE TABLE IF NOT EXISTS experiments (name TEXT PRIMARY KEY)")
            cols = [r[1] for r in self._db.execute("PRAGMA table_info(computations)")]
            if "metadata" not in cols:      # database written before the metadata column
                self._db.execute("ALTER TABLE computations ADD COLUMN metadata TEXT")
            self._db.commit()
            names = [n for (n,) in self._db.execute("SELECT name FROM experiments").fetchall()]
        if self._shared_files:
            names = D.broadcast_object(names)
        for exp_name in names:
            self._experiments[exp_name] = Experiment(self, exp_name)

    @property
    def is_writer(self) -> bool:
        """False only on the non-zero ranks of a shared persistent project: they neither create
        files nor touch the database."""
        return not self._shared_files or D.is_root()

    def barrier(self):
        if self._shared_files:
            D.barrier()

    @property
    def experiments(self) -> Dict[str, Experiment]:
        return self._experiments

    @property
    def active_experiments(self) -> Dict[str, Experiment]:
        return {k: v for k, v in self._experiments.items() if v.active}

    @property
    def run(self) -> "RunComputation":
        return RunComputation(experiments=list(self.active_experiments.values()))

    def add_experiment(self, name: str = "__missing__", timestep: float = None,
                       temperature: float = None, units: Union[str, Units] = None,
                       cluster_mode: bool = None, active: bool = True, simulation_data=None,
                       update_with_pubchempy: bool = False) -> Experiment:
        if name == "__missing__":
            raise ValueError("Experiment name can not be empty! "
                             "Use None to automatically generate a unique name.")
        if name is None:
            name = f"Experiment_{datetime.now().strftime('%Y%m%d-%H%M%S')}"
        if name in self._experiments:
            log.info("This experiment already exists")
            return self._experiments[name]
        exp = Experiment(self, name, time_step=timestep, temperature=temperature, units=units,
                         cluster_mode=cluster_mode)
        exp.active = active
        self._experiments[name] = exp
        if self.is_writer:
            self._db.execute("INSERT OR IGNORE INTO experiments VALUES (?)", (name,))
            self._db.commit()
        if simulation_data is not None:
            exp.add_data(simulation_data, update_with_pubchempy=update_with_pubchempy)
        exp._save_metadata()
        return exp

    # -- computation cache ------------------------------------------------------------------------------
    def find_computation(self, name: str, experiment: str, parameters: dict) -> Optional[Computation]:
        """Cached result with identical arguments and experiment version, or None.  In a
        sharded project rank 0 decides (and, for a shared database, provides the record), so
        every rank takes the same hit / miss branch."""
        row = None
        if self.is_writer:
            want = json.dumps(parameters, sort_keys=True)
            row = self._db.execute(
                "SELECT id, results, metadata FROM computations WHERE name=? AND experiment=? "
                "AND parameters=?", (name, experiment, want)).fetchone()
            if row is not None:
                row = (row[0], bytes(row[1]), row[2])
        if self._shared_files:
            row = D.broadcast_object(row)
        elif self.sharded:
            hit = D.broadcast_object(row is not None)
            if hit != (row is not None):
                raise RuntimeError("ranks disagree about a cached computation: the per-rank "
                                   "projects of a sharded run have diverged")
        if row is None:
            return None
        return Computation(name, experiment, parameters, decode_results(row[1]), comp_id=row[0],
                           metadata=json.loads(row[2]) if row[2] else {})

    def store_computation(self, name: str, experiment: str, parameters: dict,
                          results: "OrderedDict[str, dict]", metadata: dict = None):
        if self.is_writer:
            self._db.execute(
                "INSERT INTO computations (name, experiment, parameters, results, metadata) "
                "VALUES (?,?,?,?,?)",
                (name, experiment, json.dumps(parameters, sort_keys=True),
                 sqlite3.Binary(encode_results(results)),
                 json.dumps(metadata) if metadata else None))
            self._db.commit()
        self.barrier()


# Result blobs: a JSON index in which every long numeric series is replaced by a reference into
# a packed float64 payload (the reference stores plain JSON text, database/scheme.py:270-333;
# series are ~50x cheaper packed).  Nothing read back from the database is ever executed.
_RESULT_MAGIC = b"MDKR0001"
_SERIES_MIN = 16


def _pack_value(val, payload: list, cursor: list):
    if isinstance(val, dict):
        return {str(k): _pack_value(v, payload, cursor) for k, v in val.items()}
    if isinstance(val, np.ndarray):
        if val.ndim == 1 and val.dtype.kind == "f" and len(val) >= _SERIES_MIN:
            # a float series goes to the payload as it is (no per-element Python objects)
            arr = np.ascontiguousarray(val, dtype=np.float64)
            ref = {"__f64__": [cursor[0], len(arr)]}
            payload.append(arr)
            cursor[0] += len(arr)
            return ref
        val = val.tolist()
    if isinstance(val, (np.floating, np.integer)):
        return val.item()
    if isinstance(val, (list, tuple)):
        if len(val) >= _SERIES_MIN and all(type(v) is float for v in val):
            arr = np.asarray(val, dtype=np.float64)
            ref = {"__f64__": [cursor[0], len(arr)]}
            payload.append(arr)
            cursor[0] += len(arr)
            return ref
        return [_pack_value(v, payload, cursor) for v in val]
    if isinstance(val, (str, int, float, bool)) or val is None:
        return val
    raise TypeError(f"result value of type {type(val)} cannot be stored")


def _unpack_value(val, payload: np.ndarray):
    if isinstance(val, dict):
        if set(val) == {"__f64__"}:
            off, n = val["__f64__"]
            return payload[off:off + n].tolist()
        return {k: _unpack_value(v, payload) for k, v in val.items()}
    if isinstance(val, list):
        return [_unpack_value(v, payload) for v in val]
    return val


def encode_results(results: "OrderedDict[str, dict]") -> bytes:
    payload, cursor = [], [0]
    index = [[k, _pack_value(v, payload, cursor)] for k, v in results.items()]
    text = json.dumps(index).encode()
    text += b" " * (-len(text) % 8)
    body = np.concatenate(payload).tobytes() if payload else b""
    return _RESULT_MAGIC + len(text).to_bytes(8, "little") + text + body


def decode_results(blob: bytes) -> "OrderedDict[str, dict]":
    if blob[:8] != _RESULT_MAGIC:
        raise ValueError("project.db: unknown result record format (not written by this version)")
    n = int.from_bytes(blob[8:16], "little")
    index = json.loads(blob[16:16 + n].decode())
    payload = np.frombuffer(blob, dtype=np.float64, offset=16 + n)
    return OrderedDict((k, _unpack_value(v, payload)) for k, v in index)


def args_to_parameters(args, version: int) -> dict:
    """The (name -> serialised value) record compared by the cache
    (calculator_database.py:130-153 + :174-194)."""
    assert is_dataclass(args)
    out = {f.name: conv_to_db(getattr(args, f.name)) for f in fields(args)}
    out["version"] = version
    return out


# ------------------------------------------------------------------------------------------
# RunComputation
# ------------------------------------------------------------------------------------------
class RunComputation:
    """experiment/run.py:58-242 (hot-path subset)."""

    def __init__(self, experiment: Experiment = None, experiments: List[Experiment] = None):
        self.experiment = experiment
        self.experiments = experiments
        self.kwargs = {"experiment": experiment, "experiments": experiments}

    def _transformation(self, cls):
        def wrapper(*args, **kwargs):
            exps = self.experiments if self.experiments is not None else [self.experiment]
            for exp in exps:
                exp.cls_transformation_run(cls(), *args, **kwargs)
        return wrapper

    # transformations
    @property
    def CoordinateUnwrapper(self):
        from .transformations import CoordinateUnwrapper
        return self._transformation(CoordinateUnwrapper)

    @property
    def UnwrapViaIndices(self):
        from .transformations import UnwrapViaIndices
        return self._transformation(UnwrapViaIndices)

    @property
    def VelocityFromPositions(self):
        from .transformations import VelocityFromPositions
        return self._transformation(VelocityFromPositions)

    @property
    def IonicCurrent(self):
        from .transformations import IonicCurrent
        return self._transformation(IonicCurrent)

    @property
    def TranslationalDipoleMoment(self):
        from .transformations import TranslationalDipoleMoment
        return self._transformation(TranslationalDipoleMoment)

    @property
    def IntegratedHeatCurrent(self):
        from .transformations import IntegratedHeatCurrent
        return self._transformation(IntegratedHeatCurrent)

    @property
    def ThermalFlux(self):
        from .transformations import ThermalFlux
        return self._transformation(ThermalFlux)

    @property
    def MomentumFlux(self):
        from .transformations import MomentumFlux
        return self._transformation(MomentumFlux)

    # calculators
    @property
    def EinsteinHelfandThermalConductivity(self):
        from .calculators import EinsteinHelfandThermalConductivity
        return EinsteinHelfandThermalConductivity(**self.kwargs)

    @property
    def GreenKuboThermalConductivity(self):
        from .calculators import GreenKuboThermalConductivity
        return GreenKuboThermalConductivity(**self.kwargs)

    @property
    def GreenKuboViscosityFlux(self):
        from .calculators import GreenKuboViscosityFlux
        return GreenKuboViscosityFlux(**self.kwargs)

    @property
    def GreenKuboViscosity(self):
        from .calculators import GreenKuboViscosity
        return GreenKuboViscosity(**self.kwargs)

    @property
    def EinsteinHelfandIonicConductivity(self):
        from .calculators import EinsteinHelfandIonicConductivity
        return EinsteinHelfandIonicConductivity(**self.kwargs)

    @property
    def PotentialOfMeanForce(self):
        from .calculators import PotentialOfMeanForce
        return PotentialOfMeanForce(**self.kwargs)

    @property
    def KirkwoodBuffIntegral(self):
        from .calculators import KirkwoodBuffIntegral
        return KirkwoodBuffIntegral(**self.kwargs)

    @property
    def RadialDistributionFunction(self):
        from .calculators import RadialDistributionFunction
        return RadialDistributionFunction(**self.kwargs)

    @property
    def AngularDistributionFunction(self):
        from .calculators import AngularDistributionFunction
        return AngularDistributionFunction(**self.kwargs)

    @property
    def CoordinationNumbers(self):
        from .calculators import CoordinationNumbers
        return CoordinationNumbers(**self.kwargs)

    @property
    def EinsteinDiffusionCoefficients(self):
        from .calculators import EinsteinDiffusionCoefficients
        return EinsteinDiffusionCoefficients(**self.kwargs)

    @property
    def GreenKuboDiffusionCoefficients(self):
        from .calculators import GreenKuboDiffusionCoefficients
        return GreenKuboDiffusionCoefficients(**self.kwargs)

    @property
    def GreenKuboIonicConductivity(self):
        from .calculators import GreenKuboIonicConductivity
        return GreenKuboIonicConductivity(**self.kwargs)
