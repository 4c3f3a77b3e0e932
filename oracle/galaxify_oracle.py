"""CPU oracle for the galaxify hot path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module, and
only as the checker or the timed CPU baseline — never as part of the product path (nbody-deep-sim_b200/ does not
import it and has no CPU fallback).

It restates, function by function, the reference's algorithm (bikuta6/nbody-deep-sim, src/galaxify/simulation.py)
with the same torch CPU operators in the same order, so that on the same machine it reproduces the reference's
FP32 results bit for bit when evaluated unchunked; row-chunking (needed beyond N ~ 16k, where the reference's
(N,N,3) temporaries no longer fit) keeps every per-row operation identical and only changes how torch splits the
reduction, which moves results at the 1e-7 level at most.

Parity pin: the reference ships no tests, golden vectors or fixtures for this path, so the pin is
tests/golden/*.npz — outputs of the UNMODIFIED reference imported from /root/reference in the build container by
tests/golden/make_golden.py (committed next to the vectors). tests/test_oracle.py checks this module against them.
"""

from __future__ import annotations

import numpy as np
import torch


def _f32(x, device="cpu") -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.float32).contiguous()
    return torch.as_tensor(np.asarray(x), dtype=torch.float32).contiguous().to(device)


def accelerations(pos, mass, g_const: float, softening: float, rows: slice | None = None, chunk: int | None = None,
                  device="cpu"):
    """FP32 accelerations of the bodies in `rows` (default: all) — simulation.py:71-89.

    diff[i,j] = r_j - r_i (:80); dist_sq = sum(diff^2) + softening^2 (:82, the Python double is cast to FP32 by the
    add); inv = dist_sq^-1.5 (:83); diagonal forced to 0 (:85); acc = G * sum_j diff * inv * m_j (:86-88).
    `chunk` bounds the number of i-rows materialised at once. `device="cuda"` runs the same operators on the GPU,
    i.e. the reference's own `device="cuda"` path (simulation.py:46-51), used as the on-box oracle where the CPU is
    too slow (tools/validate_config2.py).
    """
    pos = _f32(pos, device)
    mass = _f32(mass, device)
    n = pos.shape[0]
    lo, hi, _ = (rows or slice(None)).indices(n)
    if chunk is None:
        chunk = max(1, min(hi - lo, (256 << 20) // max(1, 12 * n)))  # ~256 MB per (rows,N,3) temporary
    out = torch.empty((hi - lo, 3), dtype=torch.float32, device=pos.device)
    for a in range(lo, hi, chunk):
        b = min(hi, a + chunk)
        diff = pos.unsqueeze(0) - pos[a:b].unsqueeze(1)  # (rows, N, 3): r_j - r_i
        dist_sq = (diff**2).sum(dim=2) + softening**2
        inv_dist_cube = dist_sq.pow(-1.5)
        idx = torch.arange(a, b, device=pos.device)
        inv_dist_cube[idx - a, idx] = 0  # the rows' share of fill_diagonal_(0)
        out[a - lo : b - lo] = g_const * (diff * inv_dist_cube.unsqueeze(2) * mass.unsqueeze(0).unsqueeze(2)).sum(dim=1)
    return out


def energies(pos, vel, mass, g_const: float, softening: float, chunk: int | None = None, device="cpu"):
    """(u_energy, k_energy) — simulation.py:91-115. Softening enters as |r| + eps (:105).

    Unchunked (the default up to N = 8192) it is the reference's operator sequence; chunked, the upper-triangle sum
    is accumulated row block by row block in FP64.
    """
    pos, vel, mass = _f32(pos, device), _f32(vel, device), _f32(mass, device)
    n = pos.shape[0]
    kinetic = 0.5 * mass * (vel**2).sum(dim=1)
    k_energy = kinetic.sum().item()
    if chunk is None and n <= 8192:
        diff = pos.unsqueeze(0) - pos.unsqueeze(1)
        dist = (diff**2).sum(dim=2).sqrt() + softening
        dist.masked_fill_(torch.eye(n, dtype=torch.bool, device=pos.device), float("inf"))
        potential = -g_const * (mass.unsqueeze(0) * mass.unsqueeze(1)) / dist
        return potential.triu(1).sum().item(), k_energy
    chunk = chunk or max(1, (256 << 20) // max(1, 12 * n))
    u = 0.0
    cols = torch.arange(n, device=pos.device)
    for a in range(0, n, chunk):
        b = min(n, a + chunk)
        diff = pos.unsqueeze(0) - pos[a:b].unsqueeze(1)
        dist = (diff**2).sum(dim=2).sqrt() + softening
        potential = -g_const * (mass.unsqueeze(0) * mass[a:b].unsqueeze(1)) / dist
        upper = cols.unsqueeze(0) > torch.arange(a, b, device=pos.device).unsqueeze(1)  # j > i
        u += torch.where(upper, potential, torch.zeros((), device=pos.device)).sum(dtype=torch.float64).item()
    return u, k_energy


class State:
    """Mutable (positions, velocities, accelerations, masses) in FP32, as BaseSimulator holds them (:58-69)."""

    def __init__(self, pos, vel, mass, g_const=1.0, softening=0.1, dt=0.01, device="cpu", chunk=None):
        self.pos, self.vel, self.mass = _f32(pos, device).clone(), _f32(vel, device).clone(), _f32(mass, device).clone()
        self.g_const, self.softening, self.dt, self.device, self.chunk = g_const, softening, dt, device, chunk
        self.acc = self._force()  # :69

    def _force(self):
        return accelerations(self.pos, self.mass, self.g_const, self.softening, chunk=self.chunk, device=self.device)

    def leapfrog_step(self):
        """Kick-drift-kick, simulation.py:164-170: each update is a rounded multiply followed by a rounded add."""
        self.vel += 0.5 * self.dt * self.acc
        self.pos += self.dt * self.vel
        self.acc = self._force()
        self.vel += 0.5 * self.dt * self.acc

    def euler_step(self):
        """simulation.py:183-187: force at the old positions, then v, then x with the new v."""
        self.acc = self._force()
        self.vel += self.dt * self.acc
        self.pos += self.dt * self.vel

    def energies(self):
        return energies(self.pos, self.vel, self.mass, self.g_const, self.softening, device=self.device)


def run(pos, vel, mass, *, integrator: str, steps: int, g_const=1.0, softening=0.1, dt=0.01, calc_energy=False,
        keep=None, device="cpu", chunk=None):
    """The loop of BaseSimulator.run (simulation.py:117-146) over a State.

    Returns {step: dict(pos, vel, acc[, u, k])} for the 0-based steps in `keep` (default: all), plus the State.
    """
    st = State(pos, vel, mass, g_const, softening, dt, device=device, chunk=chunk)
    advance = {"leapfrog": st.leapfrog_step, "euler": st.euler_step}[integrator]
    keep = set(range(steps)) if keep is None else set(keep)
    out = {}
    for s in range(steps):
        advance()
        if s in keep:
            rec = dict(pos=st.pos.cpu().clone().numpy(), vel=st.vel.cpu().clone().numpy(),
                       acc=st.acc.cpu().clone().numpy())
            if calc_energy:
                rec["u"], rec["k"] = st.energies()
            out[s] = rec
    return out, st


def accelerations_f64(pos, mass, g_const: float, softening: float, rows: slice | None = None, chunk: int = 512):
    """The same formula evaluated in FP64 from the FP32-rounded inputs: the 'exact' answer the FP32 results
    scatter around (the reference itself sits 3-4e-7 from it). Used where the FP32 restatement
    cannot reach (N >= 65k, on a sample of rows). numpy, chunked."""
    p = np.asarray(_f32(pos).numpy(), dtype=np.float64)
    m = np.asarray(_f32(mass).numpy(), dtype=np.float64)
    n = p.shape[0]
    lo, hi, _ = (rows or slice(None)).indices(n)
    eps2 = float(np.float32(softening**2))
    g = float(np.float32(g_const))
    out = np.empty((hi - lo, 3))
    for a in range(lo, hi, chunk):
        b = min(hi, a + chunk)
        diff = p[None, :, :] - p[a:b, None, :]
        d2 = (diff * diff).sum(axis=2) + eps2
        with np.errstate(divide="ignore"):  # softening = 0 puts 0**-1.5 on the (masked) diagonal
            inv = d2**-1.5
        inv[np.arange(b - a), np.arange(a, b)] = 0.0
        out[a - lo : b - lo] = g * np.einsum("ijk,ij,j->ik", diff, inv, m)
    return out
