"""Ingest: LAMMPS text dumps and in-memory arrays -> TrajectoryStore.

Restates the parts of mdsuite/file_io the hot path's inputs come from:
lammps_trajectory_files.py:39-243 (LAMMPSTrajectoryFile), tabular_text_files.py:57-220 and
script_input.py:8-45 (ScriptInput), simulation_database.py:43-227 (metadata dataclasses).
A writer for the minimal dump format is included for the synthetic configs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

# lammps_trajectory_files.py:39-66 -- column names -> property names
var_names = {
    "Positions": ["x", "y", "z"],
    "Scaled_Positions": ["xs", "ys", "zs"],
    "Unwrapped_Positions": ["xu", "yu", "zu"],
    "Scaled_Unwrapped_Positions": ["xsu", "ysu", "zsu"],
    "Velocities": ["vx", "vy", "vz"],
    "Forces": ["fx", "fy", "fz"],
    "Box_Images": ["ix", "iy", "iz"],
    "Dipole_Orientation_Magnitude": ["mux", "muy", "muz"],
    "Angular_Velocity_Spherical": ["omegax", "omegay", "omegaz"],
    "Angular_Velocity_Non_Spherical": ["angmomx", "angmomy", "angmomz"],
    "Torque": ["tqx", "tqy", "tqz"],
    "Charge": ["q"],
    "Kinetic_Energy": ["c_KE"],
    "Potential_Energy": ["c_PE"],
    "Stress": ["c_Stress[1]", "c_Stress[2]", "c_Stress[3]", "c_Stress[4]", "c_Stress[5]",
               "c_Stress[6]"],
}


@dataclass
class PropertyInfo:
    name: str
    n_dims: int


@dataclass
class SpeciesInfo:
    name: str
    n_particles: int
    properties: List[PropertyInfo]
    mass: float = None
    charge: float = 0


@dataclass
class TrajectoryMetadata:
    n_configurations: int
    species_list: List[SpeciesInfo]
    box_l: list = None
    sample_rate: int = 1
    temperature: float = None


@dataclass
class TrajectoryChunkData:
    """species -> property -> (n_particles, chunk_size, n_dims)."""

    chunk_size: int
    data: Dict[str, Dict[str, np.ndarray]] = field(default_factory=dict)


class ScriptInput:
    """In-memory trajectory (script_input.py:8-45): ``data`` maps species -> property ->
    array of shape (n_frames, n_particles, n_dims) (frame-major, as MD codes produce it) or,
    with ``atom_major=True``, (n_particles, n_frames, n_dims)."""

    def __init__(self, data: Dict[str, Dict[str, np.ndarray]], box_l, sample_rate: int = 1,
                 charges: Optional[Dict[str, float]] = None, atom_major: bool = False,
                 name: str = "script", rows: Optional[Dict[str, tuple]] = None,
                 n_particles: Optional[Dict[str, int]] = None):
        """``rows`` / ``n_particles`` (multi-rank ingest of an atom-sharded store): the arrays of
        species ``sp`` hold only the global atom rows ``rows[sp] = (lo, hi)`` of its
        ``n_particles[sp]`` atoms -- every rank passes the block it owns instead of the whole
        trajectory."""
        self.name = name
        self.box_l = [float(b) for b in box_l]
        self.sample_rate = int(sample_rate)
        self.charges = charges or {}
        self.atom_major = atom_major
        self.data = data
        self.rows = rows
        self.n_particles = n_particles
        if (rows is None) != (n_particles is None):
            raise ValueError("rows and n_particles go together")

    def arrays(self):
        """Yields (species, property, atom-major array, global row range or None)."""
        for sp, props in self.data.items():
            for prop, arr in props.items():
                arr = np.asarray(arr)
                if not self.atom_major:
                    arr = np.swapaxes(arr, 0, 1)  # simulation_database.py:364-368
                yield sp, prop, arr, (self.rows[sp] if self.rows is not None else None)

    @property
    def metadata(self) -> TrajectoryMetadata:
        species, n_cfg = [], None
        for sp, props in self.data.items():
            infos, n_part = [], None
            for prop, arr in props.items():
                shape = np.shape(arr)
                n_part = shape[0] if self.atom_major else shape[1]
                if self.n_particles is not None:
                    n_part = int(self.n_particles[sp])
                n_cfg = shape[1] if self.atom_major else shape[0]
                infos.append(PropertyInfo(prop, shape[2]))
            species.append(SpeciesInfo(sp, n_part, infos, charge=self.charges.get(sp, 0)))
        return TrajectoryMetadata(n_cfg, species, box_l=self.box_l, sample_rate=self.sample_rate)


class BlockInput:
    """Atom-block-wise ingest for trajectories that are produced (or read) a block of atoms at
    a time and are too large to hold twice: the metadata is given up front, ``blocks()`` then
    yields ``(species, property, (lo, hi), array[hi - lo][n_frames][n_dims])`` for global atom
    rows [lo, hi).  Every block is written straight into the store's (page-locked) dataset; on
    an atom-sharded store a rank keeps the rows it owns, so a producer may skip blocks of other
    ranks altogether."""

    def __init__(self, n_frames: int, species: Dict[str, tuple], box_l, blocks,
                 sample_rate: int = 1, charges: Optional[Dict[str, float]] = None,
                 name: str = "blocks"):
        """species: name -> (n_particles, {property: n_dims}); blocks: callable -> iterator."""
        self.name = name
        self.n_frames = int(n_frames)
        self.species = species
        self.box_l = [float(b) for b in box_l]
        self.sample_rate = int(sample_rate)
        self.charges = charges or {}
        self.blocks = blocks

    @property
    def metadata(self) -> TrajectoryMetadata:
        sl = [SpeciesInfo(sp, int(n), [PropertyInfo(p, int(d)) for p, d in props.items()],
                          charge=self.charges.get(sp, 0))
              for sp, (n, props) in self.species.items()]
        return TrajectoryMetadata(self.n_frames, sl, box_l=self.box_l,
                                  sample_rate=self.sample_rate)


class LAMMPSTrajectoryFile:
    """Reader for LAMMPS text dumps (``*.lammpstraj``): 9 header lines per frame, then one row
    per atom; rows are sorted by ``id`` per frame; species from the ``element`` column if
    present, else ``type`` (lammps_trajectory_files.py:100-243)."""

    n_header_lines = 9

    def __init__(self, file_path: str, trajectory_is_sorted_by_ids: bool = False,
                 native: bool = True):
        self.file_path = str(file_path)
        self.sorted_by_ids = trajectory_is_sorted_by_ids
        # native=True parses with the C++ tokenizer of libmdk (mdk_lammps_scan / mdk_lammps_read);
        # native=False keeps the pure-Python parser (used to cross-check the two in the tests)
        self.native = native
        self._meta = None

    def _read_header(self, fh):
        return [fh.readline() for _ in range(self.n_header_lines)]

    def _scan_native(self):
        import ctypes as C

        from . import _lib

        lib = _lib.load()
        n_atoms, n_frames = C.c_longlong(), C.c_longlong()
        steps = (C.c_longlong * 2)()
        box = (C.c_double * 6)()
        cols = C.create_string_buffer(4096)
        _lib.check(lib.mdk_lammps_scan(self.file_path.encode(), C.byref(n_atoms), C.byref(n_frames),
                                       steps, box, cols, 4096), "mdk_lammps_scan")
        columns = cols.value.decode().split()
        first = self._read_native(0, 1, int(n_atoms.value), columns)[0][0]  # id-sorted frame 0
        return (int(n_atoms.value), int(n_frames.value), columns,
                [box[2 * d + 1] - box[2 * d] for d in range(3)], int(steps[1] - steps[0]), first)

    def _read_native(self, offset: int, n: int, n_atoms: int, columns):
        import ctypes as C

        from . import _lib

        out = np.empty((n, n_atoms, len(columns)), dtype=np.float64)
        off = C.c_longlong(offset)
        _lib.check(_lib.load().mdk_lammps_read(self.file_path.encode(), n_atoms, len(columns),
                                               columns.index("id"), int(self.sorted_by_ids), n,
                                               C.byref(off), out.ctypes.data_as(C.c_void_p)),
                   "mdk_lammps_read")
        return out, int(off.value)

    def _species_names_of_first_frame(self, columns, sp_col, n_atoms):
        """Species labels (strings) of the id-sorted first frame."""
        with open(self.file_path) as fh:
            for _ in range(self.n_header_lines):
                fh.readline()
            rows = [fh.readline().split() for _ in range(n_atoms)]
        ids = np.array([float(r[columns.index("id")]) for r in rows])
        order = np.arange(n_atoms) if self.sorted_by_ids else np.argsort(ids, kind="stable")
        return [rows[i][columns.index(sp_col)] for i in order]

    def _scan(self):
        if self.native:
            n_atoms, n_cfg, columns, box_l, sample_rate, _first = self._scan_native()
            if "id" not in columns:
                raise ValueError("LAMMPS dump needs an 'id' column")
            sp_col = "element" if "element" in columns else ("type" if "type" in columns else None)
            if sp_col is None:
                raise ValueError("LAMMPS dump needs an 'element' or 'type' column")
            names = self._species_names_of_first_frame(columns, sp_col, n_atoms)
            self._finish_scan(columns, n_atoms, n_cfg, names, box_l, sample_rate)
            return
        with open(self.file_path) as fh:
            header = self._read_header(fh)
            n_atoms = int(header[3].split()[0])
            columns = header[8].split()[2:]
            if "id" not in columns:
                raise ValueError("LAMMPS dump needs an 'id' column")
            sp_col = "element" if "element" in columns else ("type" if "type" in columns else None)
            if sp_col is None:
                raise ValueError("LAMMPS dump needs an 'element' or 'type' column")
            box_l = [float(header[5 + d].split()[1]) - float(header[5 + d].split()[0])
                     for d in range(3)]
            first = [fh.readline().split() for _ in range(n_atoms)]
            step0 = int(header[1].split()[0])
            header2 = self._read_header(fh)
            sample_rate = 1
            if header2[1].strip():
                sample_rate = int(header2[1].split()[0]) - step0
            # count lines for the number of configurations (:124-127)
            fh.seek(0)
            n_lines = sum(1 for _ in fh)
        if n_lines % (n_atoms + self.n_header_lines) != 0:
            raise ValueError("line count is not a multiple of (n_atoms + 9): truncated dump?")
        n_cfg = n_lines // (n_atoms + self.n_header_lines)
        ids = np.array([float(r[columns.index("id")]) for r in first])
        order = np.arange(n_atoms) if self.sorted_by_ids else np.argsort(ids, kind="stable")
        names = [first[i][columns.index(sp_col)] for i in order]
        self._finish_scan(columns, n_atoms, n_cfg, names, box_l, sample_rate)

    def _finish_scan(self, columns, n_atoms, n_cfg, names, box_l, sample_rate):
        species_rows: Dict[str, List[int]] = {}
        for sorted_pos, nm in enumerate(names):
            species_rows.setdefault(nm, []).append(sorted_pos)
        props = {}
        for prop, cols in var_names.items():
            if all(c in columns for c in cols):
                props[prop] = [columns.index(c) for c in cols]
        self._columns, self._n_atoms, self._n_cfg = columns, n_atoms, n_cfg
        self._species_rows = {k: np.asarray(v) for k, v in species_rows.items()}
        self._props = props
        self._box_l, self._sample_rate = box_l, max(sample_rate, 1)

    @property
    def metadata(self) -> TrajectoryMetadata:
        if self._meta is None:
            self._scan()
            sl = [SpeciesInfo(sp, len(rows), [PropertyInfo(p, len(c)) for p, c in self._props.items()])
                  for sp, rows in self._species_rows.items()]
            self._meta = TrajectoryMetadata(self._n_cfg, sl, box_l=self._box_l,
                                            sample_rate=self._sample_rate)
        return self._meta

    def get_configurations_generator(self, batch_size: int = 64):
        """Yields TrajectoryChunkData (tabular_text_files.py:122-220)."""
        self.metadata
        id_col = self._columns.index("id")
        if self.native:
            done, offset = 0, 0
            while done < self._n_cfg:
                k = min(batch_size, self._n_cfg - done)
                block, offset = self._read_native(offset, k, self._n_atoms, self._columns)
                yield self._chunk_from_block(block, k)
                done += k
            return
        with open(self.file_path) as fh:
            done = 0
            while done < self._n_cfg:
                k = min(batch_size, self._n_cfg - done)
                block = np.empty((k, self._n_atoms, len(self._columns)), dtype=np.float64)
                for f in range(k):
                    for _ in range(self.n_header_lines):
                        fh.readline()
                    rows = [fh.readline().split() for _ in range(self._n_atoms)]
                    tab = np.array([[_to_float(v) for v in r] for r in rows])
                    if not self.sorted_by_ids:
                        tab = tab[np.argsort(tab[:, id_col], kind="stable")]
                    block[f] = tab
                done += k
                yield self._chunk_from_block(block, k)

    def _chunk_from_block(self, block, k):
        """(k, n_atoms, n_columns) id-sorted table -> per species / property arrays."""
        chunk = TrajectoryChunkData(k)
        for sp, rows in self._species_rows.items():
            chunk.data[sp] = {
                prop: np.swapaxes(block[:, rows][:, :, cols], 0, 1)
                for prop, cols in self._props.items()
            }
        return chunk


# lammps_flux_files.py:41-50 -- column names of a LAMMPS log / flux table -> property names
flux_var_names = {
    "Temperature": ["temp"],
    "Time": ["time"],
    "Thermal_Flux": ["c_flux_thermal[1]", "c_flux_thermal[2]", "c_flux_thermal[3]"],
    "Stress_Visc": ["pxy", "pxz", "pyz"],
}


class LAMMPSFluxFile:
    """Reader for LAMMPS flux / log tables (lammps_flux_files.py:53-156): ``n_header_lines``
    lines of header, the last of which names the columns, then one row of numbers per sampled
    step; reading stops at the first line whose column count differs (log files interleave data
    blocks with text).  The table carries no box or sample rate, so both are passed in.  Every
    recognised column group becomes a system observable ``Observables/<Property>`` of shape
    (1, n_steps, n_dims)."""

    def __init__(self, file_path: str, sample_rate: int, box_l: list, n_header_lines: int = 2,
                 custom_data_map: Optional[Dict[str, List[str]]] = None):
        self.file_path = str(file_path)
        self.sample_rate = int(sample_rate)
        self.box_l = [float(b) for b in box_l]
        self.n_header_lines = int(n_header_lines)
        self.column_names = dict(flux_var_names)
        if custom_data_map:
            self.column_names.update(custom_data_map)
        self._table = None

    def _read(self):
        if self._table is not None:
            return
        with open(self.file_path) as fh:
            header = [fh.readline() for _ in range(self.n_header_lines)]
            columns = header[-1].split()
            rows, n_columns = [], None
            for line in fh:
                tok = line.split()
                if n_columns is None:
                    n_columns = len(tok)
                if len(tok) != n_columns or not tok:
                    break
                rows.append([_to_float(t) for t in tok])
        if not rows:
            raise ValueError(f"{self.file_path}: no data rows after {self.n_header_lines} header lines")
        if len(columns) > n_columns and columns[0] == "#":
            columns = columns[1:]          # '# time temp ...' style header
        self._table = np.asarray(rows, dtype=np.float64)
        self._props = {prop: [columns.index(c) for c in cols]
                       for prop, cols in self.column_names.items()
                       if all(c in columns for c in cols)}

    @property
    def metadata(self) -> TrajectoryMetadata:
        self._read()
        props = [PropertyInfo(p, len(c)) for p, c in self._props.items()]
        return TrajectoryMetadata(len(self._table), [SpeciesInfo("Observables", 1, props)],
                                  box_l=self.box_l, sample_rate=self.sample_rate)

    def get_configurations_generator(self, batch_size: int = 4096):
        self._read()
        for t0 in range(0, len(self._table), batch_size):
            block = self._table[t0:t0 + batch_size]
            chunk = TrajectoryChunkData(len(block))
            chunk.data["Observables"] = {p: block[None][:, :, cols]
                                         for p, cols in self._props.items()}
            yield chunk


def _to_float(tok: str) -> float:
    try:
        return float(tok)
    except ValueError:
        return np.nan  # element symbols etc.


def write_lammps_dump(path: str, species_data: Dict[str, Dict[str, np.ndarray]], box_l,
                      step_stride: int = 1, charges: Optional[Dict[str, float]] = None,
                      fmt: str = "%.9g"):
    """Minimal LAMMPS dump from atom-major arrays species -> property ->
    (n_atoms, n_frames, n_dims).  Columns: id type element x y z [vx vy vz] [ix iy iz] [q].
    float32 values survive the ``%.9g`` round trip exactly."""
    names = list(species_data)
    props = list(next(iter(species_data.values())))
    cols = ["id", "type", "element"]
    for p in props:
        cols += var_names[p]
    if charges:
        cols.append("q")
    n_frames = next(iter(next(iter(species_data.values())).values())).shape[1]
    counts = [species_data[s][props[0]].shape[0] for s in names]
    n_atoms = sum(counts)
    with open(path, "w") as fh:
        for f in range(n_frames):
            fh.write("ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n%d\n" % (f * step_stride, n_atoms))
            fh.write("ITEM: BOX BOUNDS pp pp pp\n")
            for d in range(3):
                fh.write("0.0 %r\n" % float(box_l[d]))
            fh.write("ITEM: ATOMS " + " ".join(cols) + "\n")
            aid = 1
            for ti, s in enumerate(names):
                vals = np.concatenate([np.asarray(species_data[s][p][:, f, :]) for p in props], axis=1)
                for row in vals:
                    txt = " ".join(fmt % v for v in row)
                    q = (" " + repr(float(charges[s]))) if charges else ""
                    fh.write(f"{aid} {ti + 1} {s} {txt}{q}\n")
                    aid += 1
