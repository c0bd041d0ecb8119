"""B200-native N-body emulator forward pass (drop-in for ``jax_nbody_emulator``'s hot path).

Same public names as the reference package (``__init__.py:30-95``); the compute path is
hand-written sm_100a CUDA (tcgen05 / TMEM / TMA) behind the C ABI in ``include/nbe.h``.
"""
from .nbody_emulator import (NBodyEmulator, create_emulator, load_default_parameters,
                             modulate_emulator_parameters, modulate_emulator_parameters_vel)
from .subbox import SubboxConfig, SubboxProcessor
from .cosmology import growth_factor, hubble_rate, growth_rate, dlogH_dloga, vel_norm, acc_norm
from .models import (StyleNBodyEmulatorCore, StyleNBodyEmulatorVelCore, NBodyEmulatorCore,
                     NBodyEmulatorVelCore, init_params)
from .density import (get_delta_from_psi, deconvolve_mas_kernel, power_spectrum, mas_name_from_worder,
                      za_displacement_from_delta)
from ._lib import NBEError

__version__ = "0.1.0"

__all__ = [
    "create_emulator", "NBodyEmulator", "SubboxConfig", "SubboxProcessor", "load_default_parameters",
    "modulate_emulator_parameters", "modulate_emulator_parameters_vel",
    "growth_factor", "hubble_rate", "growth_rate", "dlogH_dloga", "vel_norm", "acc_norm",
    "StyleNBodyEmulatorCore", "StyleNBodyEmulatorVelCore", "NBodyEmulatorCore", "NBodyEmulatorVelCore",
    "get_delta_from_psi", "deconvolve_mas_kernel", "power_spectrum", "mas_name_from_worder",
    "za_displacement_from_delta",
]
