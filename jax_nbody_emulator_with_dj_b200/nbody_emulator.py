"""Factory / bundle, drop-in for the reference ``nbody_emulator.py``.

create_emulator (:268-384), NBodyEmulator (:23-112), load_default_parameters (:115-129),
modulate_emulator_parameters (:150-187) and modulate_emulator_parameters_vel (:221-266) keep
their names, arguments, defaults and error messages.  The premodulation math runs in the
fused CUDA modulation kernel (nbe_modulate) and is read back as fp32.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Any

import numpy as np

from ._engine import LAYERS, Engine
from .cosmology import growth_factor
from .subbox import SubboxConfig, SubboxProcessor


@dataclass
class NBodyEmulator:
    """Container for emulator components with convenient access methods."""
    model: Any
    params: dict | None
    processor: SubboxProcessor | None
    premodulate: bool = False
    compute_vel: bool = True
    dtype: Any = np.float32

    def apply(self, x, z, Om):
        """Apply the model directly to x (B, C, D, H, W); returns displacement or
        (displacement, velocity)."""
        if self.params is None:
            raise ValueError("No parameters loaded. Use load_params=True in create_emulator.")
        from .cosmology import growth_factor, vel_norm
        z = np.atleast_1d(np.asarray(z, dtype=np.float32))
        Om = np.atleast_1d(np.asarray(Om, dtype=np.float32))
        Dz = growth_factor(z, Om)
        if self.compute_vel:
            vel_fac = vel_norm(z, Om)
        x = _astype(x, self.dtype)
        if self.premodulate:
            if self.compute_vel:
                return self.model.apply(self.params, x, Dz, vel_fac)
            return self.model.apply(self.params, x, Dz)
        if self.compute_vel:
            return self.model.apply(self.params, x, Om, Dz, vel_fac)
        return self.model.apply(self.params, x, Om, Dz)

    def process_box(self, input_box, z, Om, desc="Processing subboxes", show_progress=True, **kw):
        if self.processor is None:
            raise ValueError("No processor created. Use create_processor=True in create_emulator.")
        return self.processor.process_box(input_box, z, Om, desc=desc, show_progress=show_progress, **kw)

    def __call__(self, x, z, Om):
        return self.apply(x, z, Om)


def _astype(x, dtype):
    if hasattr(x, "detach"):
        import torch
        m = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}
        name = dtype if isinstance(dtype, str) else (str(dtype).split(".")[-1] if isinstance(dtype, torch.dtype)
                                                     else np.dtype(dtype).name)
        return x.to(m[name])
    return np.asarray(x).astype(dtype)


PARAMS_ENV = "NBE_PARAMS"
DEFAULT_PARAMS_PATH = Path(__file__).parent / "model_parameters" / "nbody_emulator_params.npz"


def load_default_parameters(path=None) -> dict:
    """Load the pretrained parameters (reference: nbody_emulator.py:115-129).

    The file is the reference's ``nbody_emulator_params.npz``: one pickled object array ``params``
    holding the tree ``{block: {layer: {weight, bias, style_weight, style_bias}}}``.  Looked up in
    this order: the ``path`` argument, the ``NBE_PARAMS`` environment variable, then
    ``model_parameters/nbody_emulator_params.npz`` inside this package (the reference's location;
    the blob itself is not redistributed with either checkout)."""
    import os
    params_path = Path(path) if path is not None else Path(os.environ.get(PARAMS_ENV) or DEFAULT_PARAMS_PATH)
    if not params_path.exists():
        raise FileNotFoundError(
            f"{params_path} not found: copy the reference package's model_parameters/nbody_emulator_params.npz "
            f"there, or point {PARAMS_ENV} (or the path argument) at it")
    with np.load(params_path, allow_pickle=True) as f:
        params = f['params'].item()
    return {'params': params}


def _modulate_tree(params, z, Om, vel, eps):
    """Both premodulation entry points: the math runs in the fused CUDA modulation kernel for all 33
    layers at once and is read back as fp32.  Like the reference (nbody_emulator.py:166-185, :238-264)
    the decision is per layer: layers with ``style_weight`` are modulated, anything else is passed
    through with a 'skipping' note.  The kernel needs the complete styled net, so a tree in which only
    some of the 33 conv layers carry style parameters is rejected instead of being silently returned
    un-modulated."""
    Dz = np.float32(growth_factor(z, Om))
    P = params['params']
    idx = {(b, l): i for i, (b, l, *_r) in enumerate(LAYERS)}
    styled = [(b, l) for (b, l) in idx if b in P and l in P[b] and 'style_weight' in P[b][l]]
    eng = None
    if styled:
        if len(styled) != len(idx):
            missing = sorted(set(idx) - set(styled))
            raise ValueError(f"cannot modulate a partially styled tree: {len(missing)} of the 33 conv layers have no "
                             f"'style_weight' (first: {missing[0][0]}/{missing[0][1]})")
        eng = Engine.get()
        eng.set_params(params, False, vel, eps)
        eng.modulate(np.float32(Om), Dz)
    out = {'params': {}}
    for bname, bp in P.items():
        out['params'][bname] = {}
        for lname, lp in bp.items():
            if eng is not None and 'style_weight' in lp and (bname, lname) in idx:
                w, dw = eng.get_modulated(idx[(bname, lname)], 0, want_dw=vel)
                ent = {'weight': w, 'bias': lp['bias']}
                if vel:
                    ent = {'weight': w, 'dweight': dw, 'bias': lp['bias']}
                out['params'][bname][lname] = ent
            else:
                print(f'skipping {bname} {lname}')
                out['params'][bname][lname] = lp
    if eng is not None:
        eng.invalidate()
    return out


def modulate_emulator_parameters(params, z, Om, eps=1.e-8):
    """Premodulate all layers for fixed (z, Om): {'weight', 'bias'} per layer."""
    return _modulate_tree(params, z, Om, False, eps)


def modulate_emulator_parameters_vel(params, z, Om, eps=1.e-8):
    """Premodulate all layers for fixed (z, Om): {'weight', 'dweight', 'bias'} per layer; the
    two layers fed by the raw field (conv_l00/{conv_0,skip}) carry the extra weight/Dz term."""
    return _modulate_tree(params, z, Om, True, eps)


def create_emulator(premodulate: bool = False, compute_vel: bool = True, load_params: bool = True,
                    processor_config: SubboxConfig | None = None, premodulate_z: float | None = None,
                    premodulate_Om: float | None = None, dtype=None, **model_kwargs) -> NBodyEmulator:
    """Factory: model (+ params, + processor).  Same arguments and errors as the reference.
    Extra keyword (not in the reference): ``params_path`` -- file for load_default_parameters."""
    params_path = model_kwargs.pop("params_path", None)
    from .models import (NBodyEmulatorCore, NBodyEmulatorVelCore, StyleNBodyEmulatorCore,
                         StyleNBodyEmulatorVelCore)
    precision = model_kwargs.pop("precision", None)
    if premodulate:
        model = NBodyEmulatorVelCore(**model_kwargs) if compute_vel else NBodyEmulatorCore(**model_kwargs)
    else:
        model = StyleNBodyEmulatorVelCore(**model_kwargs) if compute_vel else StyleNBodyEmulatorCore(**model_kwargs)
    if precision is not None:
        model.precision = precision

    params = None
    if load_params:
        params = load_default_parameters(params_path)
        if premodulate:
            if premodulate_z is None or premodulate_Om is None:
                raise ValueError("premodulate_z and premodulate_Om are required "
                                 "when premodulate=True and load_params=True")
            if compute_vel:
                params = modulate_emulator_parameters_vel(params, premodulate_z, premodulate_Om)
            else:
                params = modulate_emulator_parameters(params, premodulate_z, premodulate_Om)

    processor = None
    if processor_config is not None:
        processor = SubboxProcessor(model, params, processor_config)
    if processor_config is not None:
        dtype = processor_config.dtype
    elif dtype is None:
        dtype = np.float32
    return NBodyEmulator(model=model, params=params, processor=processor, premodulate=premodulate,
                         compute_vel=compute_vel, dtype=dtype)
