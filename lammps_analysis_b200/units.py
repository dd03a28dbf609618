"""Unit systems and physical constants (mirrors mdsuite/utils/units.py:29-98)."""
from __future__ import annotations

from dataclasses import dataclass

avogadro_constant = 6.02214076e23
elementary_charge = 1.602176634e-19
boltzmann_constant = 1.380649e-23
golden_ratio = 1.618033988749895


@dataclass(frozen=True)
class Units:
    time: float
    length: float
    energy: float
    NkTV2p: float
    boltzmann: float
    temperature: float
    pressure: float
    avogadro: float = avogadro_constant
    elementary_charge: float = elementary_charge

    @property
    def volume(self) -> float:
        return self.length**3


REAL = Units(time=1e-15, length=1e-10, energy=4184 / 6.02214076e23, NkTV2p=68568.415,
             boltzmann=0.0019872067, temperature=1, pressure=101325.0)
METAL = Units(time=1e-12, length=1e-10, energy=1.6022e-19, NkTV2p=1.6021765e6,
              boltzmann=8.617343e-5, temperature=1, pressure=100000)
SI = Units(time=1, length=1, energy=1, NkTV2p=1.380649e-23, boltzmann=1.386049e-23,
           temperature=1, pressure=1)
units_dict = {"real": REAL, "metal": METAL, "si": SI}


def resolve_units(units) -> Units:
    """str | Units -> Units (experiment.py:89-231 accepts either)."""
    if isinstance(units, Units):
        return units
    try:
        return units_dict[str(units).lower()]
    except KeyError:
        raise KeyError(f"unknown unit system {units!r}; available: {sorted(units_dict)}")
