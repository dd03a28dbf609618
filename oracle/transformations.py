"""
Oracle restatement of the transformations on the hot path (TEST INFRASTRUCTURE).

Follows:
  mdsuite/transformations/unwrap_coordinates.py:51-81    (CoordinateUnwrapper.transform_batch)
  mdsuite/transformations/unwrap_via_indices.py:49-57    (UnwrapViaIndices.transform_batch)
  mdsuite/transformations/ionic_current.py:48-58         (IonicCurrent.transform_batch)
  mdsuite/transformations/transformations.py:446-519, 553-619 (batch loop with carry-over)
  mdsuite/database/simulation_database.py:491-497, 626   (fp32 storage, fp64 compute)

All arithmetic is float64 (transformations.py:95); ``store_f32`` applies the
float32 rounding that happens when the result is written back to the HDF5 store.
"""
from __future__ import annotations

import numpy as np


def unwrap_transform_batch(pos: np.ndarray, box_l: np.ndarray, carryover=None):
    """pos: (A, T_b, 3) float64; box_l broadcastable (1, 1, 3) float64."""
    pos = np.asarray(pos, dtype=np.float64)
    box_l = np.asarray(box_l, dtype=np.float64).reshape(1, 1, 3)
    if carryover is None:
        last_pos = pos[:, 0, :]
        last_image_box = np.zeros_like(last_pos)
    else:
        last_pos = np.asarray(carryover["last_pos"], dtype=np.float64)
        last_image_box = np.asarray(carryover["last_image_box"], dtype=np.float64)
    image_box = np.concatenate([last_pos[:, None, :], pos], axis=1)
    image_box = np.diff(image_box, axis=1)
    image_box = np.round(image_box / box_l)  # tf.math.round: half to even == np.round
    image_box = -np.cumsum(image_box, axis=1)
    image_box = image_box + last_image_box[:, None, :]
    unwrapped_pos = pos + image_box * box_l
    carry = {"last_pos": pos[:, -1, :], "last_image_box": image_box[:, -1, :]}
    return unwrapped_pos, carry


def unwrap_via_indices_transform_batch(pos, box_im, box_l):
    pos = np.asarray(pos, dtype=np.float64)
    box_im = np.asarray(box_im, dtype=np.float64)
    box_l = np.asarray(box_l, dtype=np.float64).reshape(1, 1, 3)
    return pos + box_im * box_l


def ionic_current_transform_batch(batch: dict):
    """batch: {species: {"Velocities": (A, T_b, 3), "Charge": (1,1,1) | (A, T_b, 1)}}."""
    currents = []
    for properties in batch.values():
        vel = np.asarray(properties["Velocities"], dtype=np.float64)
        charge = np.asarray(properties["Charge"], dtype=np.float64)
        currents.append(np.sum(charge * vel, axis=0))
    out = currents[0]
    for c in currents[1:]:  # tf.add_n
        out = out + c
    return out


def store_f32(x):
    """Round trip through the float32 HDF5 dataset (simulation_database.py:491-497, :626)."""
    return np.asarray(x, dtype=np.float32).astype(np.float64)


def run_unwrap(pos_f32: np.ndarray, box, batch_size: int):
    """SingleSpeciesTrafo.run_transformation for one species: batch loop with carry.

    pos_f32: (A, T, 3) stored (float32) positions.  Returns the stored float32
    unwrapped positions.
    """
    A, T, _ = pos_f32.shape
    out = np.empty((A, T, 3), dtype=np.float32)
    carry = None
    n_batches, remainder = divmod(T, batch_size)
    bounds = [(b * batch_size, (b + 1) * batch_size) for b in range(n_batches)]
    if remainder:
        bounds.append((n_batches * batch_size, T))
    for lo, hi in bounds:
        res, carry = unwrap_transform_batch(
            pos_f32[:, lo:hi].astype(np.float64), np.asarray(box, dtype=np.float64), carry
        )
        out[:, lo:hi] = res.astype(np.float32)
    return out


def run_ionic_current(vel_by_species: dict, charge_by_species: dict):
    """MultiSpeciesTrafo.run_transformation, single batch; returns float32 (1, T, 3)."""
    batch = {
        sp: {
            "Velocities": np.asarray(v, dtype=np.float32).astype(np.float64),
            "Charge": np.asarray(charge_by_species[sp], dtype=np.float64).reshape(1, 1, -1)
            if np.ndim(charge_by_species[sp]) < 3
            else np.asarray(charge_by_species[sp], dtype=np.float64),
        }
        for sp, v in vel_by_species.items()
    }
    J = ionic_current_transform_batch(batch)  # (T, 3)
    return J[np.newaxis].astype(np.float32)  # transformations.py:204-207


def dipole_moment_transform_batch(batch: dict):
    """translational_dipole_moment.py:52-62: sum_species sum_atoms charge * unwrapped_pos."""
    dipms = []
    for properties in batch.values():
        pos = np.asarray(properties["Unwrapped_Positions"], dtype=np.float64)
        charge = np.asarray(properties["Charge"], dtype=np.float64)
        dipms.append(np.sum(charge * pos, axis=0))
    out = dipms[0]
    for d in dipms[1:]:
        out = out + d
    return out


def momentum_flux_transform_batch(batch: dict):
    """momentum_flux.py:45-55: sum over atoms (and species) of stress components 3, 4, 5."""
    fluxes = []
    for properties in batch.values():
        stress = np.asarray(properties["Stress"], dtype=np.float64)
        phi = np.stack([stress[:, :, 3], stress[:, :, 4], stress[:, :, 5]], axis=2)
        fluxes.append(np.sum(phi, axis=0))
    out = fluxes[0]
    for f in fluxes[1:]:
        out = out + f
    return out


def integrated_heat_current_transform_batch(batch: dict):
    """integrated_heat_current.py:49-60: sum_a pos * (KE + PE)."""
    currents = []
    for properties in batch.values():
        pos = np.asarray(properties["Unwrapped_Positions"], dtype=np.float64)
        ke = np.asarray(properties["Kinetic_Energy"], dtype=np.float64)
        pe = np.asarray(properties["Potential_Energy"], dtype=np.float64)
        currents.append(np.sum(pos * (ke + pe), axis=0))
    out = currents[0]
    for c in currents[1:]:
        out = out + c
    return out


def thermal_flux_transform_batch(batch: dict):
    """thermal_flux.py:51-92: sum_a (KE + PE) v - S v with the 6-component stress."""
    fluxes = []
    for properties in batch.values():
        stress = np.asarray(properties["Stress"], dtype=np.float64)
        vel = np.asarray(properties["Velocities"], dtype=np.float64)
        ke = np.asarray(properties["Kinetic_Energy"], dtype=np.float64)
        pe = np.asarray(properties["Potential_Energy"], dtype=np.float64)
        phi_x = (stress[:, :, 0] * vel[:, :, 0] + stress[:, :, 3] * vel[:, :, 1]
                 + stress[:, :, 4] * vel[:, :, 2])
        phi_y = (stress[:, :, 3] * vel[:, :, 0] + stress[:, :, 1] * vel[:, :, 1]
                 + stress[:, :, 5] * vel[:, :, 2])
        phi_z = (stress[:, :, 4] * vel[:, :, 0] + stress[:, :, 5] * vel[:, :, 1]
                 + stress[:, :, 2] * vel[:, :, 2])
        phi = np.dstack([phi_x, phi_y, phi_z])
        phi_sum_atoms = phi.sum(axis=0)
        energy = ke + pe
        energy_velocity_atoms = np.sum(energy * vel, axis=0)
        fluxes.append(energy_velocity_atoms - phi_sum_atoms)
    out = fluxes[0]
    for f in fluxes[1:]:
        out = out + f
    return out
