"""Synthetic inputs: the reference's own test fixtures restated, and the benchmark clouds.

Reference fixtures (Tests/RendererTests/TestUtils.swift):
  makeProjectionMatrix :25-71, makeCameraParams :74-94, generateGridGaussians :144-186,
  generateVisibleGaussians :190-231 (both seeded with srand48/drand48), makePackedBuffers :236-276;
  the 1000-Gaussian scene of DepthFirstUnitTests.swift:21-117.
Benchmark clouds: SURVEY.md section 8(d) recipe (C1..C5 of BASELINE.md section 5).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

PACKED_F32_DTYPE = np.dtype(
    [("px", "<f4"), ("py", "<f4"), ("pz", "<f4"), ("opacity", "<f4"), ("sx", "<f4"), ("sy", "<f4"),
     ("sz", "<f4"), ("_pad0", "<f4"), ("rot", "<f4", (4,))]
)  # PackedWorldGaussian, BridgingTypes.h:58-64
PACKED_F16_DTYPE = np.dtype(
    [("px", "<f4"), ("py", "<f4"), ("pz", "<f4"), ("opacity", "<f2"), ("sx", "<f2"), ("sy", "<f2"),
     ("sz", "<f2"), ("rx", "<f2"), ("ry", "<f2"), ("rz", "<f2"), ("rw", "<f2"), ("_pad0", "<f2"),
     ("_pad1", "<f2")]
)  # PackedWorldGaussianHalf, BridgingTypes.h:67-73
assert PACKED_F32_DTYPE.itemsize == 48 and PACKED_F16_DTYPE.itemsize == 32


class Drand48:
    """POSIX srand48/drand48 (48-bit LCG), as used by the reference fixtures."""

    A, Cc, M = 0x5DEECE66D, 0xB, 1 << 48

    def __init__(self, seed: int):
        self.x = ((seed & 0xFFFFFFFF) << 16) | 0x330E

    def __call__(self) -> float:
        self.x = (self.A * self.x + self.Cc) % self.M
        return self.x / float(self.M)


def make_projection_matrix(width, height, near=0.1, far=10.0, fov_degrees=60.0, convention="openCV"):
    """TestUtils.swift:37-71. Returns m[col][row] float32 (flatten() is column-major)."""
    aspect = np.float32(width) / np.float32(height)
    fov = np.float32(fov_degrees) * np.float32(math.pi) / np.float32(180.0)
    f = np.float32(1.0) / np.float32(math.tan(float(fov) / 2.0))
    near, far = np.float32(near), np.float32(far)
    m = np.zeros((4, 4), np.float32)
    m[0] = (f / aspect, 0, 0, 0)
    m[1] = (0, f, 0, 0)
    if convention == "openCV":
        m[2] = (0, 0, far / (far - near), 1)
        m[3] = (0, 0, -(far * near) / (far - near), 0)
    else:
        m[2] = (0, 0, far / (near - far), -1)
        m[3] = (0, 0, (far * near) / (near - far), 0)
    return m


def focal_lengths(width, height, fov_degrees=60.0):
    """TestUtils.swift:82-93."""
    aspect = width / height
    f = 1.0 / math.tan(math.radians(fov_degrees) / 2.0)
    return width * f / (2 * aspect), height * f / 2


def look_at_opencv(eye, target, up=(0.0, -1.0, 0.0)):
    """World->camera matrix, OpenCV convention (+X right, +Y down, +Z forward), m[col][row]."""
    eye = np.asarray(eye, np.float64)
    fwd = np.asarray(target, np.float64) - eye
    fwd /= np.linalg.norm(fwd)
    upv = np.asarray(up, np.float64)
    right = np.cross(-upv, fwd)  # +Y down => world "up" is -Y of the camera
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    R = np.stack([right, down, fwd])  # rows
    t = -R @ eye
    m = np.zeros((4, 4), np.float32)
    m[0, :3] = R[:, 0]
    m[1, :3] = R[:, 1]
    m[2, :3] = R[:, 2]
    m[3, :3] = t
    m[3, 3] = 1.0
    return m


@dataclass
class Cloud:
    positions: np.ndarray  # (N,3) f32
    scales: np.ndarray     # (N,3) f32
    rotations: np.ndarray  # (N,4) f32, (x,y,z,w)
    opacities: np.ndarray  # (N,) f32
    harmonics: np.ndarray  # (N, 3*k) f32, planar [R0..Rk-1, G.., B..] (PLYLoader.swift:700-719)
    sh_components: int

    @property
    def count(self) -> int:
        return int(self.positions.shape[0])

    def pack(self, precision: str):
        """-> (gaussian records, harmonics) as numpy arrays in the layout of `precision`."""
        n = self.count
        if precision == "float32":
            g = np.zeros(n, PACKED_F32_DTYPE)
            g["px"], g["py"], g["pz"] = self.positions.T
            g["opacity"] = self.opacities
            g["sx"], g["sy"], g["sz"] = self.scales.T
            g["rot"] = self.rotations
            return g, np.ascontiguousarray(self.harmonics, np.float32)
        g = np.zeros(n, PACKED_F16_DTYPE)
        g["px"], g["py"], g["pz"] = self.positions.T
        g["opacity"] = self.opacities.astype(np.float16)
        g["sx"], g["sy"], g["sz"] = self.scales.astype(np.float16).T
        r = self.rotations.astype(np.float16)
        g["rx"], g["ry"], g["rz"], g["rw"] = r.T
        return g, np.ascontiguousarray(self.harmonics.astype(np.float16))


def _from_lists(pos, scl, rot, opa, col) -> Cloud:
    return Cloud(np.asarray(pos, np.float32), np.asarray(scl, np.float32), np.asarray(rot, np.float32),
                 np.asarray(opa, np.float32), np.asarray(col, np.float32), 0)


def generate_grid_gaussians(count: int, seed: int = 42) -> Cloud:
    """TestUtils.swift:144-186."""
    r = Drand48(seed)
    pos, scl, rot, opa, col = [], [], [], [], []
    f32 = np.float32
    for i in range(count):
        grid = int(math.sqrt(float(count))) + 1
        x = f32(i % grid) / f32(grid) * f32(4) - f32(2)
        y = f32(i // grid) / f32(grid) * f32(4) - f32(2)
        z = f32(r() * 3 + 2)
        pos.append((x, y, z))
        s = f32(r() * 0.1 + 0.05)
        scl.append((s, s, s))
        rot.append((0, 0, 0, 1))
        opa.append(f32(r() * 0.5 + 0.5))
        col.append((f32(r() * 0.5), f32(r() * 0.5), f32(r() * 0.5)))
    return _from_lists(pos, scl, rot, opa, col)


def generate_visible_gaussians(count: int, seed: int = 42) -> Cloud:
    """TestUtils.swift:190-231."""
    r = Drand48(seed)
    pos, scl, rot, opa, col = [], [], [], [], []
    f32 = np.float32
    for _ in range(count):
        z = f32(r() * 8 + 1.5)
        spread = z * f32(0.6)
        x = f32(r() * 2 - 1) * spread
        y = f32(r() * 2 - 1) * spread
        pos.append((x, y, z))
        s = f32(r() * 0.15 + 0.08)
        scl.append((s, s, s))
        rot.append((0, 0, 0, 1))
        opa.append(f32(r() * 0.5 + 0.5))
        col.append((f32(r()), f32(r()), f32(r())))
    return _from_lists(pos, scl, rot, opa, col)


def pipeline_stages_scene() -> Cloud:
    """The 1000-Gaussian scene of DepthFirstUnitTests.swift:26-48 (640x480, shComponents=1)."""
    f32 = np.float32
    pos, scl, rot, opa, col = [], [], [], [], []
    for i in range(1000):
        row, c = i // 32, i % 32
        pos.append((f32(c) * f32(0.1) - f32(1.6), f32(row) * f32(0.1) - f32(1.6), f32(2.0) + f32(i) * f32(0.001)))
        scl.append((0.01, 0.01, 0.01))
        rot.append((0, 0, 0, 1))
        opa.append(0.8)
        col.append((f32(i % 10) / f32(10), f32((i // 10) % 10) / f32(10), f32((i // 100) % 10) / f32(10)))
    cl = _from_lists(pos, scl, rot, opa, col)
    cl.sh_components = 1
    return cl


SH_COEFFS = {0: 1, 1: 4, 2: 9, 3: 16}


def synthetic_cloud(n: int, sh_degree: int = 3, seed: int = 42, z_range=(2.0, 20.0), aspect=16.0 / 9.0,
                    scale_median=0.02, scale_sigma=0.6, lateral=0.7) -> Cloud:
    """Benchmark cloud: camera at the origin looking down +Z (OpenCV), fov 60 deg.

    positions uniform in the frustum slab z in z_range with |x|,|y| <= lateral*z*tan(30deg)*aspect;
    log-normal per-axis scales; uniform random unit quaternions; opacity = sigmoid(N(0.5, 2^2));
    SH DC ~ U(-1,1)/C0-scaled, higher orders N(0, 0.1^2). Counter-based Philox stream keyed by seed.
    """
    rng = np.random.Generator(np.random.Philox(key=seed))
    z = rng.uniform(z_range[0], z_range[1], n)
    ext = lateral * z * math.tan(math.radians(30.0)) * aspect
    x = rng.uniform(-1.0, 1.0, n) * ext
    y = rng.uniform(-1.0, 1.0, n) * ext
    pos = np.stack([x, y, z], 1).astype(np.float32)
    scl = np.exp(rng.normal(math.log(scale_median), scale_sigma, (n, 3))).astype(np.float32)
    q = rng.normal(0.0, 1.0, (n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opa = (1.0 / (1.0 + np.exp(-rng.normal(0.5, 2.0, n)))).astype(np.float32)
    k = SH_COEFFS[sh_degree]
    sh = rng.normal(0.0, 0.1, (n, 3, k))
    sh[:, :, 0] = rng.uniform(-1.0, 1.0, (n, 3)) / 0.28209479177387814 * 0.5
    return Cloud(pos, scl, q.astype(np.float32), opa, sh.reshape(n, 3 * k).astype(np.float32), k)


def orbit_cameras(n_views: int, center=(0.0, 0.0, 11.0), radius=11.0, seed: int = 7):
    """Seeded orbit of poses looking at the cloud centre (C5). Yields (view m[col][row], position)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    out = []
    for _ in range(n_views):
        az = rng.uniform(-0.6, 0.6)
        el = rng.uniform(-0.25, 0.25)
        rr = radius * rng.uniform(0.9, 1.1)
        eye = np.array(center) + rr * np.array([math.sin(az) * math.cos(el), math.sin(el),
                                                -math.cos(az) * math.cos(el)])
        out.append((look_at_opencv(eye, center), eye.astype(np.float32)))
    return out
