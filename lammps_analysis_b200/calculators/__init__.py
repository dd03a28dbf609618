"""Hot-path calculators behind the MDSuite names (mdsuite/calculators/__init__.py:29-87)."""
from .angular_distribution_function import AngularDistributionFunction
from .coordination_number_calculation import CoordinationNumbers
from .einstein_diffusion_coefficients import EinsteinDiffusionCoefficients
from .einstein_helfand_ionic_conductivity import EinsteinHelfandIonicConductivity
from .einstein_helfand_thermal_conductivity import EinsteinHelfandThermalConductivity
from .green_kubo_flux_calculators import (GreenKuboThermalConductivity, GreenKuboViscosity,
                                          GreenKuboViscosityFlux)
from .green_kubo_ionic_conductivity import GreenKuboIonicConductivity
from .green_kubo_self_diffusion_coefficients import GreenKuboDiffusionCoefficients
from .kirkwood_buff_integrals import KirkwoodBuffIntegral
from .potential_of_mean_force import PotentialOfMeanForce
from .radial_distribution_function import RadialDistributionFunction

__all__ = [
    "RadialDistributionFunction",
    "CoordinationNumbers",
    "AngularDistributionFunction",
    "EinsteinDiffusionCoefficients",
    "GreenKuboDiffusionCoefficients",
    "GreenKuboIonicConductivity",
    "EinsteinHelfandIonicConductivity",
    "EinsteinHelfandThermalConductivity",
    "GreenKuboThermalConductivity",
    "GreenKuboViscosity",
    "GreenKuboViscosityFlux",
    "PotentialOfMeanForce",
    "KirkwoodBuffIntegral",
]
