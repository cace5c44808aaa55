"""Deterministic synthetic clouds for the five BASELINE.json configs.

numpy only; shared by tests/ and bench.py.  Nothing here touches the oracle or the device.
"""
from __future__ import annotations

import numpy as np


def rotation(axis, deg: float) -> np.ndarray:
    u = np.asarray(axis, np.float64)
    u = u / np.linalg.norm(u)
    th = np.deg2rad(deg)
    K = np.array([[0, -u[2], u[1]], [u[2], 0, -u[0]], [-u[1], u[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * (K @ K)


def rigid_about(R: np.ndarray, t, center) -> np.ndarray:
    """4x4 of x -> R (x - c) + c + t"""
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = np.asarray(center) - R @ np.asarray(center) + np.asarray(t)
    return T


def extent(xyz: np.ndarray) -> float:
    return float(np.linalg.norm(xyz.max(axis=0).astype(np.float64) - xyz.min(axis=0).astype(np.float64)))


def bumpy_ellipsoid(n: int, seed: int = 1234) -> np.ndarray:
    """Config 3/4/5 surface: bumpy ellipsoid, extent ~2.74."""
    rng = np.random.default_rng(seed)
    d = rng.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    th = np.arccos(np.clip(d[:, 2], -1, 1))
    ph = np.arctan2(d[:, 1], d[:, 0])
    r = (1 + 0.15 * np.sin(3 * th) * np.cos(2 * ph) + 0.10 * np.sin(2 * th + 1) * np.sin(3 * ph + 0.5)
         + 0.05 * np.cos(5 * th) * np.cos(ph))
    return (d * r[:, None] * np.array([1.0, 0.7, 0.5])).astype(np.float32)


def perturbed_pair(src: np.ndarray, axis, deg: float, shift_frac: float, noise_frac: float = 0.0,
                   noise_seed: int = 4321):
    """target = R (src - c) + c + shift_frac*extent*x^ (+ N(0,(noise_frac*extent)^2)). Returns tgt, T_true(4x4)."""
    s64 = src.astype(np.float64)
    c = s64.mean(axis=0)
    ext = extent(src)
    T = rigid_about(rotation(axis, deg), (shift_frac * ext, 0.0, 0.0), c)
    tgt = s64 @ T[:3, :3].T + T[:3, 3]
    if noise_frac > 0:
        tgt = tgt + np.random.default_rng(noise_seed).normal(0.0, noise_frac * ext, tgt.shape)
    return tgt.astype(np.float32), T


def config_bunny(bunny: np.ndarray):
    """Config 2: bunny fixture vs itself, 5 deg about (1,1,1), 0.02*extent shift, sigma 1e-3*extent, rng 20191."""
    return perturbed_pair(bunny, (1, 1, 1), 5.0, 0.02, 1e-3, noise_seed=20191)


def config_synth(n: int = 1_000_000, seed: int = 1234, noise_frac: float = 1e-4, deg: float = 30.0,
                 shift_frac: float = 0.1):
    """Config 3 (n = 1e6) / config 4 (n = 16*2**20): 30 deg about (1,2,3), 0.1*extent shift."""
    src = bumpy_ellipsoid(n, seed)
    tgt, T = perturbed_pair(src, (1, 2, 3), deg, shift_frac, noise_frac)
    return src, tgt, T


def config_batch_pair(b: int, n: int = 32768):
    """Config 5, pair b: seed 1234+b, 10 deg about a per-pair random axis, 0.05*extent shift."""
    src = bumpy_ellipsoid(n, 1234 + b)
    axis = np.random.default_rng(900000 + b).standard_normal(3)
    tgt, T = perturbed_pair(src, axis, 10.0, 0.05, 1e-4, noise_seed=4321 + b)
    return src, tgt, T


def rotation_angle(Ra: np.ndarray, Rb: np.ndarray) -> float:
    """geodesic angle (rad) between two rotations"""
    R = np.asarray(Ra, np.float64)[:3, :3].T @ np.asarray(Rb, np.float64)[:3, :3]
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.arctan2(0.5 * np.linalg.norm(v), 0.5 * (np.trace(R) - 1)))


def shard_range(n: int, rank: int, world: int):
    """contiguous [lo, hi) slice of n units for `rank` of `world` (query sharding / pair batches)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
