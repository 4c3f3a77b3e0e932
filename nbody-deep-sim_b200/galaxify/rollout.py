"""Leapfrog rollout with the force slot left open: the integrator the surrogate-model trainers wrap around a model.

Mirrors (reference, read-only) trainer.py:217-226 `Trainer.step` and gnn.py:223-253 `GraphModel.step` / `.rollout`:

    vel_ = vel + 0.5*dt*acc ; pos_ = pos + dt*vel_ ; acc_ = model.predict(pos_, cat(vel_, m)) ; vel_ += 0.5*dt*acc_

i.e. LeapFrogSimulator.step (src/galaxify/simulation.py:164-170) with `model.predict` in place of
compute_accelerations. The kick/drift arithmetic runs in two fused sm_100a kernels (nbody_kick_drift_f32,
nbody_kick_f32: include/nbody_b200.h) with the reference's separately rounded multiply/add; the force is whatever the
caller passes as `predict(pos, feats) -> (n,3)`. `DirectSumForce` plugs this engine's own all-pairs kernel into that
slot, which turns the rollout into LeapFrogSimulator (bit for bit) and is how the op is parity-tested.

Same argument names and order as the reference: step(pos, vel, m, acc, dt) -> (pos_, vel_, acc_); new tensors are
returned, inputs are not modified. Tensors must be float32 CUDA tensors (no CPU path).
"""

from __future__ import annotations

import ctypes

import torch

from . import _native
from .simulation import _ptr


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(name: str, t: torch.Tensor, n: int | None = None) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be a float32 CUDA tensor: galaxify (B200 engine) has no CPU path")
    if t.dim() != 2 or t.shape[1] != 3 or (n is not None and t.shape[0] != n):
        raise ValueError(f"{name} must have shape (n, 3), got {tuple(t.shape)}")
    return t.contiguous()


def kick_drift(pos: torch.Tensor, vel: torch.Tensor, acc: torch.Tensor, dt: float):
    """(pos_, vel_) with vel_ = vel + 0.5*dt*acc and pos_ = pos + dt*vel_   (trainer.py:219-221)."""
    _native.lib()
    pos = _check("pos", pos)
    n = pos.shape[0]
    vel, acc = _check("vel", vel, n), _check("acc", acc, n)
    pos_, vel_ = torch.empty_like(pos), torch.empty_like(vel)
    with torch.cuda.device(pos.device):
        _native.call("nbody_kick_drift_f32", _ptr(pos), _ptr(vel), _ptr(acc), _ptr(pos_), _ptr(vel_), n,
                     _native.f32(dt), _native.f32(0.5 * dt), _stream())
    return pos_, vel_


def kick_(vel: torch.Tensor, acc: torch.Tensor, dt: float) -> torch.Tensor:
    """In place vel += 0.5*dt*acc   (trainer.py:225); returns vel."""
    _native.lib()
    vel_c = _check("vel", vel)
    if vel_c.data_ptr() != vel.data_ptr():
        raise ValueError("vel must be contiguous for the in-place kick")
    acc = _check("acc", acc, vel.shape[0])
    with torch.cuda.device(vel.device):
        _native.call("nbody_kick_f32", _ptr(vel), _ptr(acc), _ptr(vel), vel.shape[0], _native.f32(0.5 * dt), _stream())
    return vel


def step(predict, pos, vel, m, acc, dt):
    """One rollout step, trainer.py:217-226 with `predict` in the place of `self.model.predict`.

    predict(pos_, feats) receives the drifted positions and feats = cat([vel_, m], dim=-1), as the reference passes
    them (trainer.py:223), and returns the (n,3) accelerations."""
    pos_, vel_ = kick_drift(pos, vel, acc, dt)
    acc_ = predict(pos_, torch.cat([vel_, m], dim=-1))
    kick_(vel_, acc_, dt)
    return pos_, vel_, acc_


def rollout(predict, pos, vel, m, steps: int, dt: float):
    """gnn.py:234-253: {"pos": [...], "vel": [...], "acc": [...]} with steps+1 entries each (initial state first).

    The reference appends the INITIAL `vel` at every step (gnn.py:250 appends `vel`, not `vel_`); here the stepped
    velocities are recorded, which is what the variable name promises. The initial acceleration is
    predict(pos, cat(pos, vel, m)) as at gnn.py:247."""
    if m.dim() == 1:
        m = m.unsqueeze(-1)
    pos_, vel_ = pos, vel
    memory = {"pos": [pos_], "vel": [vel_]}
    acc = predict(pos_, torch.cat((pos_, vel_, m), dim=-1))
    memory["acc"] = [acc]
    for _ in range(steps):
        pos_, vel_, acc = step(predict, pos_, vel_, m, acc, dt)
        memory["pos"].append(pos_)
        memory["vel"].append(vel_)
        memory["acc"].append(acc)
    return memory


class DirectSumForce:
    """`predict` callable backed by this engine's all-pairs kernel (simulation.py:71-89): the exact force in the
    model's slot. feats' last column is the mass, as in both reference call sites."""

    def __init__(self, g_const: float = 1.0, softening: float = 0.1):
        self.g_const, self.softening = g_const, softening
        self._ws = None

    def __call__(self, pos: torch.Tensor, feats: torch.Tensor) -> torch.Tensor:
        lib = _native.lib()
        pos = _check("pos", pos)
        n = pos.shape[0]
        mass = feats[:, -1].contiguous()
        acc = torch.empty_like(pos)
        need = lib.nbody_workspace_bytes(n, n)
        if self._ws is None or self._ws.numel() < need or self._ws.device != pos.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=pos.device)
        with torch.cuda.device(pos.device):
            _native.call("nbody_accel_f32", _ptr(pos), _ptr(mass), _ptr(acc), n, _native.f32(self.g_const),
                         _native.f32(self.softening**2), _ptr(self._ws), self._ws.numel(), _stream())
        return acc

    predict = __call__
