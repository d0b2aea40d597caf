"""Deterministic synthetic LiDAR scans for tests and bench.

The reference ships no data and no generator (its only workload is the sphere demo,
/root/reference/src/chad/main.cpp:7-38). This module makes the five BASELINE.json
workloads reproducible on ANY machine: only IEEE-exact operations (+ - * / on float64,
comparisons, the final float32 rounding) and numpy's PCG64 raw stream are used -- no libm
calls (sin/cos/log differ by an ulp between CPUs), so the same seed yields bit-identical
points here, on the GPU box, and for the committed golden hashes in tests/golden/.

Geometry: a spinning multi-beam sensor at `pos`, `beams` elevation rings from -25 deg to
+15 deg, `azimuth_steps` (2048) columns, ray-cast against an axis-aligned scene, range noise
from an Irwin-Hall(12) approximation of N(0, sigma). Noise is mandatory: exactly co-planar
neighbourhoods make the reference's plane fit produce NaN.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# (cos, sin) of the per-step rotation, as exact double literals (see module docstring)
_AZ_STEP = (float.fromhex("0x1.ffff621621d02p-1"), float.fromhex("0x1.921f8becca4bap-9"))  # 2*pi/2048
_EL_STEP = {
    16: (float.fromhex("0x1.ff7210493a3dfp-1"), float.fromhex("0x1.7d22a6ca6a97ap-5")),
    32: (float.fromhex("0x1.ffdec380a9b22p-1"), float.fromhex("0x1.70f1533b5233ep-6")),
    64: (float.fromhex("0x1.fff7f3cd2ddc7p-1"), float.fromhex("0x1.6b1c15963a600p-7")),
    128: (float.fromhex("0x1.fffe05068a065p-1"), float.fromhex("0x1.684191d7a9157p-8")),
}  # 40 deg / (beams - 1)
_EL_START = (float.fromhex("0x1.d0079302dd767p-1"), float.fromhex("-0x1.b0c2d77379853p-2"))  # -25 deg
AZIMUTH_STEPS = 2048


def _rotation_table(start, step, n):
    c, s = start
    cd, sd = step
    out = np.empty((n, 2), dtype=np.float64)
    for i in range(n):
        out[i, 0] = c
        out[i, 1] = s
        c, s = c * cd - s * sd, s * cd + c * sd
    return out


_DIR_CACHE: dict[int, np.ndarray] = {}


def beam_directions(beams: int) -> np.ndarray:
    """Unit ray directions, shape (beams * 2048, 3), beam-major (sensor firing order)."""
    if beams not in _EL_STEP:
        raise ValueError(f"beams must be one of {sorted(_EL_STEP)}")
    if beams not in _DIR_CACHE:
        az = _rotation_table((1.0, 0.0), _AZ_STEP, AZIMUTH_STEPS)
        el = _rotation_table(_EL_START, _EL_STEP[beams], beams)
        d = np.empty((beams, AZIMUTH_STEPS, 3), dtype=np.float64)
        d[:, :, 0] = el[:, 0:1] * az[None, :, 0]
        d[:, :, 1] = el[:, 0:1] * az[None, :, 1]
        d[:, :, 2] = el[:, 1:2] * np.ones((1, AZIMUTH_STEPS))
        _DIR_CACHE[beams] = d.reshape(-1, 3)
    return _DIR_CACHE[beams]


@dataclass(frozen=True)
class Scene:
    """Axis-aligned scene. `half_xy` = wall distance from the world origin (walls at x,y = +-half_xy),
    `ground_z`/`ceil_z` horizontal planes (ceil_z None = open sky), `recess` > 0 adds the urban facade
    pattern: the y-walls sit `recess` metres further out wherever floor(x / 10 m) is odd."""
    half_x: float
    half_y: float
    ground_z: float
    ceil_z: float | None = None
    recess: float = 0.0
    max_range: float = 100.0


BOX_ROOM = Scene(half_x=20.0, half_y=20.0, ground_z=-1.8)
INDOOR = Scene(half_x=4.0, half_y=4.0, ground_z=-1.5, ceil_z=1.5)
URBAN = Scene(half_x=1.0e9, half_y=10.0, ground_z=-1.8, recess=2.0)


def _plane_hit(origin, direction, plane):
    """Ray parameter of the hit with an axis plane, +inf when the ray points away / is parallel."""
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (plane - origin) / direction
    return np.where((direction != 0.0) & (t > 0.0), t, np.inf)


def _irwin_hall(rng: np.random.Generator, n: int) -> np.ndarray:
    acc = rng.random(n)
    for _ in range(11):
        acc = acc + rng.random(n)
    return acc - 6.0


def lidar_scan(pos, beams: int = 64, scene: Scene = BOX_ROOM, seed: int = 1234, sigma: float = 0.01) -> np.ndarray:
    """One scan: float32 points, shape (n, 3), n <= beams * 2048 (rays beyond max_range are dropped)."""
    d = beam_directions(beams)
    px, py, pz = (float(pos[0]), float(pos[1]), float(pos[2]))
    dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
    t = _plane_hit(pz, dz, scene.ground_z)
    if scene.ceil_z is not None:
        t = np.minimum(t, _plane_hit(pz, dz, scene.ceil_z))
    t = np.minimum(t, _plane_hit(px, dx, scene.half_x))
    t = np.minimum(t, _plane_hit(px, dx, -scene.half_x))
    for sgn in (1.0, -1.0):
        t_near = _plane_hit(py, dy, sgn * scene.half_y)
        if scene.recess > 0.0:
            t_far = _plane_hit(py, dy, sgn * (scene.half_y + scene.recess))
            x_hit = px + dx * np.where(np.isfinite(t_near), t_near, 0.0)
            odd = (np.floor(x_hit / 10.0) % 2.0) == 1.0
            t_wall = np.where(odd, t_far, t_near)
        else:
            t_wall = t_near
        t = np.minimum(t, t_wall)
    rng = np.random.Generator(np.random.PCG64(seed))
    noise = _irwin_hall(rng, d.shape[0]) * sigma
    keep = t < scene.max_range
    t = np.where(keep, t, 0.0) + noise
    pts = np.empty((d.shape[0], 3), dtype=np.float64)
    pts[:, 0] = px + dx * t
    pts[:, 1] = py + dy * t
    pts[:, 2] = pz + dz * t
    return np.ascontiguousarray(pts[keep].astype(np.float32))


@dataclass(frozen=True)
class Workload:
    """One BASELINE.json config: scene, sensor, trajectory and map parameters."""
    name: str
    scene: Scene
    beams: int
    scans: int
    start_x: float
    step_x: float
    sdf_res: float
    sdf_trunc: float
    seed: int = 1234

    def pose(self, s: int) -> np.ndarray:
        return np.array([self.start_x + self.step_x * s, 0.0, 0.0], dtype=np.float32)

    def scan(self, s: int) -> tuple[np.ndarray, np.ndarray]:
        pos = self.pose(s)
        return lidar_scan(pos, self.beams, self.scene, self.seed + s), pos

    def truncated(self, scans: int) -> "Workload":
        return Workload(self.name, self.scene, self.beams, scans, self.start_x, self.step_x, self.sdf_res, self.sdf_trunc, self.seed)


# BASELINE.json `configs`, in order. configs[1] is the bench workload.
WORKLOADS = {
    "cfg0_single_64beam": Workload("cfg0_single_64beam", BOX_ROOM, 64, 1, 0.0, 0.0, 0.05, 0.10),
    "cfg1_traj100_128beam": Workload("cfg1_traj100_128beam", BOX_ROOM, 128, 100, -12.5, 0.25, 0.05, 0.10),
    "cfg2_fine_indoor": Workload("cfg2_fine_indoor", INDOOR, 64, 20, -1.0, 0.1, 0.02, 0.06),
    "cfg3_urban_5km": Workload("cfg3_urban_5km", URBAN, 128, 5000, 0.0, 1.0, 0.10, 0.20),
    "cfg4_traj1000_128beam": Workload("cfg4_traj1000_128beam", BOX_ROOM, 128, 1000, -12.5, 0.025, 0.05, 0.10),
}


def sphere_demo_points(n: int) -> np.ndarray:
    """The first `n` points of the reference demo's workload shape (main.cpp:7-38: points on a sphere of
    radius 5 m around the origin), regenerated with exact arithmetic from PCG64 instead of the demo's
    libstdc++ mt19937/uniform_real_distribution pair (whose stream is implementation-defined)."""
    rng = np.random.Generator(np.random.PCG64(420))
    v = rng.random((n, 3)) * 2.0 - 1.0
    norm2 = (v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1]) + v[:, 2] * v[:, 2]
    v = v * (1.0 / np.sqrt(norm2))[:, None] * 5.0
    return np.ascontiguousarray(v.astype(np.float32))
