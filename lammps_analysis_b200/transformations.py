"""Transformations on the hot path and the driver they share.

Mirrors mdsuite/transformations/transformations.py:
  :171-237  _save_output (float32 store, system observables get a leading axis of 1)
  :275-326  _prepare_database_entry  -- create the output dataset, or EXTEND it when the
            experiment has grown (resize + ``offset``: only the appended frames are computed)
  :328-350  find_property_per_config / find_property_single_val
  :352-388  get_prop_through_transformation (property -> transformation table, list = fallbacks)
  :390-433  input resolution: dataset per configuration -> species value -> experiment value ->
            recursive transformation -> CannotFindPropertyError
  :436-519  SingleSpeciesTrafo  (one output dataset per species, batch loop with carry-over)
  :522-619  MultiSpeciesTrafo   (one system observable ``Observables/<name>`` of shape (1, T, d))
and the ``transform_batch`` bodies whose arithmetic runs in libmdk: unwrap_coordinates.py:51-81,
unwrap_via_indices.py:49-57, ionic_current.py:48-58, translational_dipole_moment.py:52-62,
momentum_flux.py:45-55, thermal_flux.py:51-92, integrated_heat_current.py:49-60,
velocity_from_positions.py:62-77.

Differences that are deliberate:
* the reference sizes its time batches from host RAM; the results do not depend on that plan
  (the unwrap carry is an exact integer image count, observables are sums over atoms), so batches
  here are sized by HBM only and the carry is threaded through them the same way;
* extending after ``Experiment.add_data``: upstream the single-species loop returns early when
  the dataset exists (:466-473, leaving a stale, too-short output) and the multi-species path
  reads the ROW count as the old length (:300-309).  Here both follow the documented intent:
  the dataset is resized, frames [old length, new length) are computed, and the unwrap carry
  is rebuilt from the last stored frame so that the extended series is what a single run over
  the whole trajectory gives;
* atoms shard across ranks (store.py): a rank transforms / reduces the atom block it owns.
"""
from __future__ import annotations

import collections.abc
import logging
from typing import Dict, Iterable, NamedTuple, Optional

import numpy as np

from . import distributed as D
from . import kernels as K
from . import trace
from .store import join_path

log = logging.getLogger("mdsuite_b200")


class CannotFindPropertyError(Exception):
    """transformations.py:61-64."""


class CannotFindTransformationError(Exception):
    """transformations.py:55-58."""


class PropertyInfo(NamedTuple):
    """database/simulation_database.py:43-58."""
    name: str
    n_dims: int


class _Properties:
    """database/mdsuite_properties.py:34-87 (the names the hot path touches)."""
    positions = PropertyInfo("Positions", 3)
    unwrapped_positions = PropertyInfo("Unwrapped_Positions", 3)
    velocities = PropertyInfo("Velocities", 3)
    velocities_from_positions = PropertyInfo("Velocities_From_Positions", 3)
    box_images = PropertyInfo("Box_Images", 3)
    box_length = PropertyInfo("Box_Array", 3)
    charge = PropertyInfo("Charge", 1)
    ionic_current = PropertyInfo("Ionic_Current", 3)
    translational_dipole_moment = PropertyInfo("Translational_Dipole_Moment", 3)
    momentum_flux = PropertyInfo("Momentum_Flux", 3)
    thermal_flux = PropertyInfo("Thermal_Flux", 3)
    integrated_heat_current = PropertyInfo("Integrated_Heat_Current", 3)
    kinetic_energy = PropertyInfo("Kinetic_Energy", 1)
    potential_energy = PropertyInfo("Potential_Energy", 1)
    stress = PropertyInfo("Stress", 6)
    time_step = PropertyInfo("Time_Step", 1)
    sample_rate = PropertyInfo("Sample_Rate", 1)


mdsuite_properties = _Properties()


def _as_prop(p) -> PropertyInfo:
    if isinstance(p, PropertyInfo):
        return p
    if hasattr(p, "name") and hasattr(p, "n_dims"):
        return PropertyInfo(p.name, int(p.n_dims))
    for v in vars(_Properties).values():
        if isinstance(v, PropertyInfo) and v.name == p:
            return v
    return PropertyInfo(str(p), 3)


class Transformations:
    """Driver shared by all transformations (transformations.py:67-433)."""

    input_properties: list = []
    output_property = None
    # an extension run (new frames appended) works in time chunks of at most this many bytes;
    # a full run works in blocks of rows (all frames of some atoms) of at most block_bytes
    chunk_bytes: int = 8 << 30
    block_bytes: int = 1 << 30

    def __init__(self, input_properties: Iterable = None, output_property=None,
                 scale_function: dict = None, dtype=None):
        ins = input_properties if input_properties is not None else type(self).input_properties
        self.input_properties = [_as_prop(p) for p in ins]
        out = output_property if output_property is not None else type(self).output_property
        self.output_property = _as_prop(out) if out is not None else None
        self.scale_function = scale_function      # accepted for API parity (host-RAM plan)
        self.experiment = None
        self.offset = 0

    # -- :328-350 ----------------------------------------------------------------------------------
    def find_property_per_config(self, sp_name: str, prop: PropertyInfo) -> Optional[str]:
        path = join_path(sp_name, prop.name)
        return path if self.experiment.store.check_existence(path) else None

    def find_property_single_val(self, sp_name: str, prop: PropertyInfo):
        species_info = self.experiment.species.get(sp_name)
        val = getattr(species_info, prop.name.lower(), None) if species_info is not None else None
        if val is not None:
            return val
        return getattr(self.experiment, prop.name.lower(), None)

    # -- :352-388 ----------------------------------------------------------------------------------
    def get_prop_through_transformation(self, sp_name: str, prop: PropertyInfo) -> Optional[str]:
        candidates = property_to_transformation_dict().get(prop.name)
        if candidates is None:
            raise CannotFindTransformationError(
                f"was asked to get '{prop.name}' for '{sp_name}', but there is no transformation "
                "to get that property")
        if not isinstance(candidates, (list, tuple)):
            self.experiment.cls_transformation_run(candidates(), species=[sp_name])
        else:
            for cls in candidates:          # go through the list until one works
                try:
                    self.experiment.cls_transformation_run(cls(), species=[sp_name])
                except CannotFindPropertyError:
                    continue
                break
            else:
                raise CannotFindTransformationError(
                    f"was asked to get '{prop.name}' for '{sp_name}'. There are transformations "
                    f"to get this property ({candidates}), but none of them have the required "
                    "data")
        return self.find_property_per_config(sp_name, prop)

    # -- :390-433 ----------------------------------------------------------------------------------
    def resolve_inputs(self, species_names):
        """({species: {property: dataset path}}, {species: {property: constant float64 array}})
        following the reference's order: per-configuration dataset, species value, experiment
        value, then a transformation that produces the property."""
        paths: Dict[str, Dict[str, str]] = {}
        consts: Dict[str, Dict[str, np.ndarray]] = {}
        for sp in species_names:
            paths[sp], consts[sp] = {}, {}
            for prop in self.input_properties:
                path = self.find_property_per_config(sp, prop)
                if path is not None:
                    paths[sp][prop.name] = path
                    continue
                val = self.find_property_single_val(sp, prop)
                if val is not None:
                    if not isinstance(val, collections.abc.Iterable):
                        val = [val]
                    consts[sp][prop.name] = np.asarray(val, dtype=np.float64)
                    continue
                try:
                    path = self.get_prop_through_transformation(sp, prop)
                except CannotFindTransformationError:
                    path = None
                if path is None:
                    raise CannotFindPropertyError(
                        f"While performing transformation '{self.output_property.name}': "
                        f"Property '{prop.name}' for species '{sp}' cannot be found in the "
                        "simulation database nor in the simulation metadata, nor can it be "
                        "obtained by a transformation")
                paths[sp][prop.name] = path
        return paths, consts

    # -- :275-326 ----------------------------------------------------------------------------------
    def _output_path(self, species: str, system_tensor: bool = False):
        if system_tensor:
            return 1, join_path("Observables", self.output_property.name)
        return (self.experiment.species[species].n_particles,
                join_path(species, self.output_property.name))

    def _is_complete(self, species: str, system_tensor: bool = False) -> bool:
        _, path = self._output_path(species, system_tensor)
        store = self.experiment.store
        return (store.check_existence(path)
                and store.shape(path)[1] >= self.experiment.number_of_configurations)

    def _prepare_database_entry(self, species: str, system_tensor: bool = False):
        """Create the output dataset, or extend it when the experiment has grown since it was
        written (:300-311).  Returns (path, offset): ``offset`` is the first frame that still
        has to be computed.  Called after the inputs have been resolved, so that a
        transformation that cannot find its inputs leaves no empty dataset behind."""
        exp, store = self.experiment, self.experiment.store
        n_rows, path = self._output_path(species, system_tensor)
        n_cfg = exp.number_of_configurations
        if store.check_existence(path):
            self.offset = store.shape(path)[1]
            store.resize_dataset(path, n_cfg)
        else:
            store.add_dataset(path, (n_rows, n_cfg, self.output_property.n_dims))
            self.offset = 0
        return path, self.offset

    def run_transformation(self, species: Iterable[str] = None):
        raise NotImplementedError

    def transform_batch(self, batch: dict, carryover=None):
        raise NotImplementedError("transformation of a batch must be implemented")


class SingleSpeciesTrafo(Transformations):
    """transformations.py:436-519: the transformation is applied to each species separately.

    ``transform_batch(batch, carryover)`` receives {property name: CUDA float32 tensor
    [atoms][frames][dims] | constant float64 array} for the atom block this rank owns and
    returns the CUDA tensor [atoms][frames][out dims] (or [1][frames][dims], broadcast over the
    atoms), optionally with a carry-over for the next batch."""

    def initial_carry(self, species: str, paths: dict, offset: int):
        """Carry-over with which an EXTENDING run starts at frame ``offset`` > 0."""
        return None

    def run_transformation(self, species: Iterable[str] = None):
        import torch

        exp = self.experiment
        store = exp.store
        species = list(exp.species) if species is None else list(species)
        for sp in species:
            if self._is_complete(sp):
                log.info("%s already exists for %s, skipping transformation",
                         self.output_property.name, sp)
                continue  # :466-473
            paths, consts = self.resolve_inputs([sp])
            paths, consts = paths[sp], consts[sp]
            out_path, offset = self._prepare_database_entry(sp)
            n_atoms, n_frames, _ = store.shape(out_path)
            # every atom is an independent series: a rank transforms the atom block it owns
            # (store.py) and writes back only those rows -- no communication
            lo, hi = store.owned_rows(out_path)
            if hi <= lo:
                continue
            trace.mark(f"{type(self).__name__}[{sp}] inputs resolved, output dataset ready")
            if offset == 0:
                self._run_row_blocks(sp, paths, consts, out_path, lo, hi, n_frames)
                trace.mark(f"{type(self).__name__}[{sp}] row blocks enqueued")
            else:
                self._run_appended_frames(sp, paths, consts, out_path, lo, hi, offset, n_frames)

    def _run_row_blocks(self, sp, paths, consts, out_path, lo, hi, n_frames):
        """The whole time range of this rank's rows, a block of rows at a time.  Rows are
        contiguous in the store, so a block is one DMA each way; blocks are pipelined over three
        streams -- upload of block k + 1, kernel on block k, write-back of block k - 1 -- which
        keeps both directions of the host link busy.  The result stays resident in HBM for the
        calculator that follows when it fits."""
        import torch

        store = self.experiment.store
        names = list(paths)
        dims_in = sum(p.n_dims for p in self.input_properties if p.name in paths)
        row_bytes = n_frames * 4 * (dims_in + self.output_property.n_dims)
        rows_per = int(max(1, min(hi - lo, self.block_bytes // max(row_bytes, 1))))
        out_bytes = (hi - lo) * n_frames * 4 * self.output_property.n_dims
        keep = out_bytes < 0.35 * torch.cuda.mem_get_info()[0]
        whole = torch.empty(hi - lo, n_frames, self.output_property.n_dims, dtype=torch.float32,
                            device="cuda") if keep else None
        resident = {n: store.device(paths[n]) for n in names if store.is_resident(paths[n])}
        pinned = {n: store.pinned_tensor(paths[n]) for n in names if n not in resident}
        if any(t is None for t in pinned.values()) or rows_per >= hi - lo:
            # file-backed (or small) inputs: one upload per dataset through the store's cache
            batch = dict(consts)
            for n in names:
                batch[n] = store.device(paths[n])
            ret = self.transform_batch(batch, carryover=None)
            out_dev = ret[0] if isinstance(ret, tuple) else ret
            if out_dev.shape[0] == 1 and hi - lo != 1:
                out_dev = out_dev.expand(hi - lo, -1, -1).contiguous()
            store.write_from_device(out_path, out_dev, 0)
            store.invalidate(out_path)
            store.adopt_device(out_path, out_dev)
            return
        trace.mark("row blocks: device buffers allocated")
        compute = torch.cuda.current_stream()
        upload = store.upload_stream()      # shared: the species queued first is complete first
        fence = torch.cuda.Event()
        fence.record()
        upload.wait_event(fence)            # staging buffers may be recycled memory
        stage = [{n: torch.empty((rows_per,) + tuple(pinned[n].shape[1:]), dtype=torch.float32,
                                 device="cuda") for n in pinned} for _ in range(2)]
        free_ev = [None, None]          # compute is done with staging buffer i
        done = []                       # (r0, r1, event): rows of `whole` written by the kernel
        row0 = store.owned_rows(paths[names[0]])[0]
        for k, r0 in enumerate(range(lo, hi, rows_per)):
            r1 = min(hi, r0 + rows_per)
            buf = stage[k & 1]
            with torch.cuda.stream(upload):
                if free_ev[k & 1] is not None:
                    upload.wait_event(free_ev[k & 1])
                for n, pin in pinned.items():
                    buf[n][:r1 - r0].copy_(pin[r0 - row0:r1 - row0], non_blocking=True)
                    store.h2d_bytes += (r1 - r0) * int(np.prod(pin.shape[1:])) * 4
                up_ev = torch.cuda.Event()
                up_ev.record()
                trace.event(f"{sp} upload block {k} done")
            compute.wait_event(up_ev)
            batch = dict(consts)
            for n in names:
                batch[n] = resident[n][r0 - lo:r1 - lo] if n in resident else buf[n][:r1 - r0]
            if whole is not None:
                batch["__out__"] = whole[r0 - lo:r1 - lo]   # a kernel may write its result here
            ret = self.transform_batch(batch, carryover=None)
            out_dev = ret[0] if isinstance(ret, tuple) else ret
            if out_dev.shape[0] == 1 and r1 - r0 != 1:
                out_dev = out_dev.expand(r1 - r0, -1, -1).contiguous()
            if whole is not None and out_dev.data_ptr() != whole[r0 - lo:r1 - lo].data_ptr():
                whole[r0 - lo:r1 - lo].copy_(out_dev)
                out_dev = whole[r0 - lo:r1 - lo]
            ev = torch.cuda.Event()
            ev.record()
            trace.event(f"{sp} unwrap block {k} done")
            free_ev[k & 1] = ev
            done.append((r0 - lo, r1 - lo, ev))
            store.write_from_device(out_path, out_dev, 0, row0=r0)   # asynchronous, side stream
        store.invalidate(out_path)
        if whole is not None:
            # stays resident for the calculator that follows, which may start on the first row
            # blocks while the later ones are still being uploaded and transformed
            store.adopt_device(out_path, whole, blocks=done)

    def _run_appended_frames(self, sp, paths, consts, out_path, lo, hi, offset, n_frames):
        """Extension after ``Experiment.add_data``: frames [offset, n_frames) only, in time
        chunks, with the carry-over rebuilt from the last stored frame."""
        import torch

        store = self.experiment.store
        dims_in = sum(p.n_dims for p in self.input_properties if p.name in paths)
        per_frame = (hi - lo) * 4 * (dims_in + self.output_property.n_dims)
        frames_per_chunk = max(1, min(n_frames, self.chunk_bytes // max(per_frame, 1)))
        carry = self.initial_carry(sp, paths, offset)
        for t0 in range(offset, n_frames, frames_per_chunk):
            t1 = min(n_frames, t0 + frames_per_chunk)
            batch = dict(consts)
            for name, path in paths.items():
                batch[name] = torch.from_numpy(
                    np.ascontiguousarray(store.host(path)[:, t0:t1])).cuda()
            ret = self.transform_batch(batch, carryover=carry)
            out_dev, carry = ret if isinstance(ret, tuple) else (ret, carry)
            if out_dev.shape[0] == 1 and hi - lo != 1:
                out_dev = out_dev.expand(hi - lo, -1, -1).contiguous()
            store.write_from_device(out_path, out_dev, t0)
        store.invalidate(out_path)


class CoordinateUnwrapper(SingleSpeciesTrafo):
    """Box-jump unwrapping (unwrap_coordinates.py:51-81)."""

    input_properties = [mdsuite_properties.positions, mdsuite_properties.box_length]
    output_property = mdsuite_properties.unwrapped_positions

    def _box(self, batch):
        return np.asarray(batch["Box_Array"], dtype=np.float64).reshape(3)

    def transform_batch(self, batch, carryover=None):
        import torch

        pos = batch["Positions"]
        carry = carryover
        if carry is None:
            carry = {
                "last_pos": torch.zeros(pos.shape[0], 3, dtype=torch.float32, device=pos.device),
                "last_image_box": torch.zeros(pos.shape[0], 3, dtype=torch.float64,
                                              device=pos.device),
                "have": False,
            }
        out = batch.get("__out__")
        if out is None:
            out = torch.empty_like(pos)
        K.unwrap(pos, self._box(batch), carry["last_pos"], carry["last_image_box"], carry["have"],
                 out)
        carry["have"] = True
        return out, carry

    def initial_carry(self, species, paths, offset):
        """Rebuild (last wrapped position, image count) at frame offset - 1 from what is
        stored: the image is the integer rint((unwrapped - wrapped) / L)."""
        import torch

        store = self.experiment.store
        box = np.asarray(self.experiment.box_array, dtype=np.float64)
        wrapped = np.asarray(store.host(paths["Positions"])[:, offset - 1], dtype=np.float64)
        out_path = join_path(species, self.output_property.name)
        unwrapped = np.asarray(store.host(out_path)[:, offset - 1], dtype=np.float64)
        image = np.rint((unwrapped - wrapped) / box)
        return {"last_pos": torch.from_numpy(wrapped.astype(np.float32)).cuda(),
                "last_image_box": torch.from_numpy(image).cuda(), "have": True}


class UnwrapViaIndices(SingleSpeciesTrafo):
    """pos + box_images * L (unwrap_via_indices.py:49-57)."""

    input_properties = [mdsuite_properties.positions, mdsuite_properties.box_images,
                        mdsuite_properties.box_length]
    output_property = mdsuite_properties.unwrapped_positions

    def transform_batch(self, batch, carryover=None):
        import torch

        pos, img = batch["Positions"], batch["Box_Images"]
        if not isinstance(img, torch.Tensor):
            # a constant species / experiment value is not an image-flag series
            raise CannotFindPropertyError("UnwrapViaIndices needs per-configuration Box_Images")
        out = torch.empty_like(pos)
        K.unwrap_indices(pos, img, np.asarray(batch["Box_Array"], dtype=np.float64).reshape(3),
                         out)
        return out

    def find_property_single_val(self, sp_name, prop):
        if prop.name == "Box_Images":
            return None      # only a stored series will do: lets the unwrap fallback kick in
        return super().find_property_single_val(sp_name, prop)

    def get_prop_through_transformation(self, sp_name, prop):
        if prop.name == "Box_Images":
            raise CannotFindTransformationError("no transformation produces Box_Images")
        return super().get_prop_through_transformation(sp_name, prop)


class VelocityFromPositions(SingleSpeciesTrafo):
    """v(t) = (x(t + dt) - x(t)) / dt on unwrapped positions (velocity_from_positions.py:45-77)."""

    input_properties = [mdsuite_properties.unwrapped_positions, mdsuite_properties.time_step,
                        mdsuite_properties.sample_rate]
    output_property = mdsuite_properties.velocities_from_positions

    def transform_batch(self, batch, carryover=None):
        import torch

        pos = batch["Unwrapped_Positions"]
        dt = np.float32(np.asarray(batch["Time_Step"]).ravel()[0]) * \
            np.float32(np.asarray(batch["Sample_Rate"]).ravel()[0])
        out = torch.empty_like(pos)
        K.velocity_from_positions(pos, float(dt), out)
        return out


class MultiSpeciesTrafo(Transformations):
    """transformations.py:522-619: information of several species is combined into one system
    observable ``Observables/<name>`` of shape (1, n_frames, n_dims).

    ``transform_batch(batch, carryover)`` receives {species: {property: tensor | constant}} for
    the atom blocks this rank owns and returns a CUDA tensor [1][frames][dims] (float32 or
    float64).  ``reduces_over_atoms``: the result is a sum over atoms, so the ranks' partial
    results are added with one all-reduce; a transformation that is not such a sum can only run
    on one rank."""

    reduces_over_atoms = False

    def run_transformation(self, species: Iterable[str] = None):
        exp = self.experiment
        store = exp.store
        species = list(exp.species) if species is None else list(species)
        if self._is_complete(self.output_property.name, system_tensor=True):
            log.info("%s already exists for this experiment, skipping transformation",
                     self.output_property.name)
            return  # :572-579
        if D.world_size() > 1 and not self.reduces_over_atoms:
            raise NotImplementedError(f"{type(self).__name__} is not a sum over atoms and cannot "
                                      "run on an atom-sharded store")
        paths, consts = self.resolve_inputs(species)
        out_path, offset = self._prepare_database_entry(self.output_property.name,
                                                        system_tensor=True)
        batch = {}
        for sp in species:
            batch[sp] = dict(consts[sp])
            for name, path in paths[sp].items():
                dev = store.device(path)              # the atom block this rank owns
                batch[sp][name] = dev if offset == 0 else dev[:, offset:].contiguous()
        ret = self.transform_batch(batch, carryover=None)
        out_dev = ret[0] if isinstance(ret, tuple) else ret
        if out_dev.dim() == 2:
            out_dev = out_dev[None]
        if self.reduces_over_atoms:
            D.all_reduce_sum_([out_dev])
        # float64 -> float32 store rounding (:171-237)
        store.add_data(out_path, out_dev.cpu().numpy(), start=offset)


class _AtomSumObservable(MultiSpeciesTrafo):
    """Observables that are plain sums over atoms and species, accumulated in fp64 by an
    HBM-bound reduction kernel (12 B per atom-frame)."""

    reduces_over_atoms = True

    def _accumulate(self, props: dict, J):
        raise NotImplementedError

    def transform_batch(self, batch, carryover=None):
        import torch

        n_frames = None
        for props in batch.values():
            for v in props.values():
                if isinstance(v, torch.Tensor):
                    n_frames = v.shape[1]
        J = torch.zeros(n_frames, 3, dtype=torch.float64, device="cuda")
        for props in batch.values():
            first = next(v for v in props.values() if isinstance(v, torch.Tensor))
            if first.shape[0] > 0:
                self._accumulate(props, J)
        return J[None]


def _per_atom_frame(t):
    """[A][T][1] dataset -> contiguous [A][T]."""
    return t.reshape(t.shape[0], -1).contiguous()


class IonicCurrent(_AtomSumObservable):
    """J(t) = sum_species sum_atoms q v (ionic_current.py:48-58), stored as
    ``Observables/Ionic_Current`` with shape (1, n_frames, 3) (transformations.py:204-207,
    289-291).  The charge is a per-configuration dataset or the species constant."""

    input_properties = [mdsuite_properties.velocities, mdsuite_properties.charge]
    output_property = mdsuite_properties.ionic_current
    vector_property = "Velocities"

    def _accumulate(self, props, J):
        import torch

        q = props["Charge"]
        if isinstance(q, torch.Tensor):
            q = _per_atom_frame(q)                       # per-atom-frame charge dataset (A, T, 1)
        else:
            q = float(np.asarray(q).ravel()[0])          # species constant (:335-350)
        K.ionic_current(props[self.vector_property], q, J)


class TranslationalDipoleMoment(IonicCurrent):
    """M(t) = sum_species sum_atoms q r_unwrapped (translational_dipole_moment.py:52-62): the
    same charge-weighted atom reduction as the ionic current, applied to unwrapped positions
    (which are produced first when missing, transformations.py:352-388)."""

    input_properties = [mdsuite_properties.unwrapped_positions, mdsuite_properties.charge]
    output_property = mdsuite_properties.translational_dipole_moment
    vector_property = "Unwrapped_Positions"


class MomentumFlux(_AtomSumObservable):
    """Sum over atoms of the off-diagonal stress components xy, xz, yz
    (momentum_flux.py:45-55)."""

    input_properties = [mdsuite_properties.stress]
    output_property = mdsuite_properties.momentum_flux

    def _accumulate(self, props, J):
        K.flux_sum(props["Stress"], J, comp0=3)


class IntegratedHeatCurrent(_AtomSumObservable):
    """sum_a r_a (KE_a + PE_a) on unwrapped positions (integrated_heat_current.py:49-60)."""

    input_properties = [mdsuite_properties.unwrapped_positions, mdsuite_properties.kinetic_energy,
                        mdsuite_properties.potential_energy]
    output_property = mdsuite_properties.integrated_heat_current

    def _accumulate(self, props, J):
        K.flux_sum(props["Unwrapped_Positions"], J, comp0=0,
                   w1=_per_atom_frame(props["Kinetic_Energy"]),
                   w2=_per_atom_frame(props["Potential_Energy"]))


class ThermalFlux(_AtomSumObservable):
    """sum_a (KE_a + PE_a) v_a - S_a v_a (thermal_flux.py:51-92)."""

    input_properties = [mdsuite_properties.stress, mdsuite_properties.velocities,
                        mdsuite_properties.kinetic_energy, mdsuite_properties.potential_energy]
    output_property = mdsuite_properties.thermal_flux

    def _accumulate(self, props, J):
        K.thermal_flux(props["Stress"], props["Velocities"],
                       _per_atom_frame(props["Kinetic_Energy"]),
                       _per_atom_frame(props["Potential_Energy"]), J)


def property_to_transformation_dict() -> dict:
    """transformations/transformation_dict.py:46-63 (hot-path subset): which transformation
    produces which property; a list is tried in order until one finds its inputs."""
    return {
        "Integrated_Heat_Current": IntegratedHeatCurrent,
        "Ionic_Current": IonicCurrent,
        "Momentum_Flux": MomentumFlux,
        "Thermal_Flux": ThermalFlux,
        "Translational_Dipole_Moment": TranslationalDipoleMoment,
        "Unwrapped_Positions": [UnwrapViaIndices, CoordinateUnwrapper],
        "Velocities_From_Positions": VelocityFromPositions,
    }
