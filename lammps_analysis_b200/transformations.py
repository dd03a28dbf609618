"""Transformations on the hot path: CoordinateUnwrapper, UnwrapViaIndices, IonicCurrent.

Mirrors mdsuite/transformations/transformations.py:171-619 (driver: skip-if-exists, dataset
creation, input resolution, batch loop with carry-over, float32 save) and the three
``transform_batch`` bodies (unwrap_coordinates.py:51-81, unwrap_via_indices.py:49-57,
ionic_current.py:48-58), whose arithmetic runs in ``mdk_unwrap`` / ``mdk_unwrap_indices`` /
``mdk_ionic_current``.

The reference sizes its time batches from host RAM; the results do not depend on that plan
(the unwrap carry is an exact integer image count, the current is a sum over atoms), so batches
here are sized by HBM only and the carry is threaded through them the same way.
"""
from __future__ import annotations

import logging

import numpy as np

from . import distributed as D
from . import kernels as K
from .store import join_path

log = logging.getLogger("mdsuite_b200")


class CannotFindPropertyError(Exception):
    """transformations.py:61-64."""


class Transformations:
    input_properties: list = []
    output_property: str = None
    output_dims: int = 3
    # frames per launch are bounded so that input + output of a chunk stay within this budget
    chunk_bytes: int = 8 << 30

    def __init__(self):
        self.experiment = None

    def _require(self, species: str, prop: str):
        path = join_path(species, prop)
        if not self.experiment.store.check_existence(path):
            raise CannotFindPropertyError(
                f"While performing transformation '{type(self).__name__}': cannot find "
                f"'{prop}' for species '{species}' in the database, as a species value or as "
                f"an experiment value")
        return path


class _SingleSpeciesTrafo(Transformations):
    """transformations.py:436-519: one output dataset per species."""

    def run_transformation(self, species: list = None):
        import torch

        exp = self.experiment
        species = list(exp.species) if species is None else species
        for sp in species:
            out_path = join_path(sp, self.output_property)
            if exp.store.check_existence(out_path):
                log.info("%s already exists for %s, skipping", self.output_property, sp)
                continue  # transformations.py:466-473
            paths = [self._require(sp, p) for p in self.input_properties]
            n_atoms, n_frames, _ = exp.store.shape(paths[0])
            exp.store.add_dataset(out_path, (n_atoms, n_frames, self.output_dims))
            # every atom is an independent series: a rank transforms the atom block it owns
            # (store.py) and writes back only those rows -- no communication
            lo, hi = exp.store.owned_rows(out_path)
            if hi <= lo:
                continue
            per_frame = (hi - lo) * 12 * (len(paths) + 1)
            frames_per_chunk = max(1, min(n_frames, self.chunk_bytes // max(per_frame, 1)))
            carry = None
            whole = None
            for t0 in range(0, n_frames, frames_per_chunk):
                t1 = min(n_frames, t0 + frames_per_chunk)
                if t0 == 0 and t1 == n_frames:
                    inputs = [exp.store.device(p) for p in paths]
                else:
                    inputs = [torch.from_numpy(np.ascontiguousarray(exp.store.host(p)[:, t0:t1]))
                              .cuda() for p in paths]
                out_dev = torch.empty_like(inputs[0])
                carry = self.transform_batch(inputs, out_dev, carry)
                exp.store.write_from_device(out_path, out_dev, t0)
                whole = out_dev if (t0 == 0 and t1 == n_frames) else None
            exp.store.invalidate(out_path)
            if whole is not None:
                exp.store.adopt_device(out_path, whole)  # stays resident for the calculator


class CoordinateUnwrapper(_SingleSpeciesTrafo):
    """Box-jump unwrapping (unwrap_coordinates.py:51-81)."""

    input_properties = ["Positions"]
    output_property = "Unwrapped_Positions"

    def transform_batch(self, inputs, out_dev, carry):
        import torch

        (pos,) = inputs
        box = np.asarray(self.experiment.box_array, dtype=np.float64)
        if carry is None:
            carry = {
                "last_pos": torch.zeros(pos.shape[0], 3, dtype=torch.float32, device=pos.device),
                "last_image_box": torch.zeros(pos.shape[0], 3, dtype=torch.float64,
                                              device=pos.device),
                "have": False,
            }
        K.unwrap(pos, box, carry["last_pos"], carry["last_image_box"], carry["have"], out_dev)
        carry["have"] = True
        return carry


class UnwrapViaIndices(_SingleSpeciesTrafo):
    """pos + box_images * L (unwrap_via_indices.py:49-57)."""

    input_properties = ["Positions", "Box_Images"]
    output_property = "Unwrapped_Positions"

    def transform_batch(self, inputs, out_dev, carry):
        pos, img = inputs
        K.unwrap_indices(pos, img, np.asarray(self.experiment.box_array, dtype=np.float64),
                         out_dev)
        return None


class IonicCurrent(Transformations):
    """J(t) = sum_species sum_atoms q v (ionic_current.py:48-58), stored as
    ``Observables/Ionic_Current`` with shape (1, n_frames, 3) (transformations.py:204-207,
    289-291).  Atoms shard across ranks; the partial currents are summed with one all-reduce."""

    input_properties = ["Velocities", "Charge"]
    output_property = "Ionic_Current"
    vector_property = "Velocities"

    def _ensure_inputs(self, species):
        pass

    def run_transformation(self, species: list = None):
        import torch

        exp = self.experiment
        out_path = join_path("Observables", self.output_property)
        if exp.store.check_existence(out_path):
            log.info("%s already exists, skipping", self.output_property)
            return  # transformations.py:572-579
        species = list(exp.species) if species is None else species
        self._ensure_inputs(species)
        n_frames = exp.number_of_configurations
        J = torch.zeros(n_frames, 3, dtype=torch.float64, device="cuda")
        for sp in species:
            vpath = self._require(sp, self.vector_property)
            lo, hi = exp.store.owned_rows(vpath)
            if hi <= lo:
                continue
            vel = exp.store.device(vpath, rows=(lo, hi))
            qpath = join_path(sp, "Charge")
            if exp.store.check_existence(qpath):
                # per-atom-frame charge dataset (A, T, 1)
                q = exp.store.device(qpath, rows=(lo, hi)).reshape(hi - lo, n_frames).contiguous()
            else:
                # species constant (transformations.py:335-350 find_property_single_val)
                q = float(exp.species[sp].charge)
            K.ionic_current(vel, q, J)
        D.all_reduce_sum_([J])
        exp.store.put(out_path, J.cpu().numpy()[None])  # float64 -> float32 store rounding


class TranslationalDipoleMoment(IonicCurrent):
    """M(t) = sum_species sum_atoms q r_unwrapped (translational_dipole_moment.py:52-62): the
    same charge-weighted atom reduction as the ionic current, applied to unwrapped positions
    (which are produced first when missing, transformations.py:352-388)."""

    input_properties = ["Unwrapped_Positions", "Charge"]
    output_property = "Translational_Dipole_Moment"
    vector_property = "Unwrapped_Positions"

    def _ensure_inputs(self, species):
        exp = self.experiment
        missing = [sp for sp in species
                   if not exp.store.check_existence(join_path(sp, "Unwrapped_Positions"))]
        if missing:
            first = next(iter(exp.species))
            if exp.store.check_existence(join_path(first, "Box_Images")):
                exp.run.UnwrapViaIndices(species=missing)
            else:
                exp.run.CoordinateUnwrapper(species=missing)


class _AtomSumObservable(Transformations):
    """MultiSpeciesTrafo pattern (transformations.py:522-619) for observables that are plain
    sums over atoms and species: ``Observables/{output_property}`` of shape (1, n_frames, 3).
    Atoms shard across ranks; the partial sums meet in one all-reduce (as IonicCurrent)."""

    def _ensure_inputs(self, species):
        pass

    def _accumulate(self, sp, lo, hi, J):
        raise NotImplementedError

    def run_transformation(self, species: list = None):
        import torch

        exp = self.experiment
        out_path = join_path("Observables", self.output_property)
        if exp.store.check_existence(out_path):
            log.info("%s already exists, skipping", self.output_property)
            return  # transformations.py:572-579
        species = list(exp.species) if species is None else species
        self._ensure_inputs(species)
        n_frames = exp.number_of_configurations
        J = torch.zeros(n_frames, 3, dtype=torch.float64, device="cuda")
        for sp in species:
            paths = [self._require(sp, p) for p in self.input_properties]
            lo, hi = exp.store.owned_rows(paths[0])
            if hi > lo:
                self._accumulate(sp, lo, hi, J)
        D.all_reduce_sum_([J])
        exp.store.put(out_path, J.cpu().numpy()[None])  # float64 -> float32 store rounding

    def _dev(self, sp, prop, lo, hi):
        return self.experiment.store.device(join_path(sp, prop), rows=(lo, hi))


class MomentumFlux(_AtomSumObservable):
    """Sum over atoms of the off-diagonal stress components xy, xz, yz
    (momentum_flux.py:45-55)."""

    input_properties = ["Stress"]
    output_property = "Momentum_Flux"

    def _accumulate(self, sp, lo, hi, J):
        K.flux_sum(self._dev(sp, "Stress", lo, hi), J, comp0=3)


class IntegratedHeatCurrent(_AtomSumObservable):
    """sum_a r_a (KE_a + PE_a) on unwrapped positions (integrated_heat_current.py:49-60)."""

    input_properties = ["Unwrapped_Positions", "Kinetic_Energy", "Potential_Energy"]
    output_property = "Integrated_Heat_Current"

    def _ensure_inputs(self, species):
        TranslationalDipoleMoment._ensure_inputs(self, species)

    def _accumulate(self, sp, lo, hi, J):
        ke = self._dev(sp, "Kinetic_Energy", lo, hi)
        pe = self._dev(sp, "Potential_Energy", lo, hi)
        K.flux_sum(self._dev(sp, "Unwrapped_Positions", lo, hi), J, comp0=0,
                   w1=ke.reshape(hi - lo, -1).contiguous(), w2=pe.reshape(hi - lo, -1).contiguous())


class ThermalFlux(_AtomSumObservable):
    """sum_a (KE_a + PE_a) v_a - S_a v_a (thermal_flux.py:51-92)."""

    input_properties = ["Stress", "Velocities", "Kinetic_Energy", "Potential_Energy"]
    output_property = "Thermal_Flux"

    def _accumulate(self, sp, lo, hi, J):
        ke = self._dev(sp, "Kinetic_Energy", lo, hi).reshape(hi - lo, -1).contiguous()
        pe = self._dev(sp, "Potential_Energy", lo, hi).reshape(hi - lo, -1).contiguous()
        K.thermal_flux(self._dev(sp, "Stress", lo, hi), self._dev(sp, "Velocities", lo, hi), ke,
                       pe, J)
