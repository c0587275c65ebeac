"""The five BASELINE.json configurations as (name, scene, frame constants, limits, HiZ pyramid):
one definition shared by bench.py (`--config K`), the full-size `-m gpu` parity tests and tools/bench_configs.py."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import api, scenes as S


@dataclass
class Workload:
    key: int
    name: str
    scene: S.Scene
    frame_constants: np.ndarray  # FRAME_CONSTANTS_DTYPE[2] (current, last)
    config: api.Config
    hiz: tuple | None = None  # (pyramid, size, mips) when the config culls against a far-HiZ


HEADLINE = "icosphere subdiv8 (1310720 base tris, 20480 clusters), 2048^2 noise displacement, PN on, 1X+2X transient on, camera 1.5r, 0.75 px/segment @3840x2160"

NAMES = {
    1: "config 1: plane 256x256 (131072 base tris, 2048 clusters), 512^2 noise displacement, oblique camera, factors 1-11",
    2: HEADLINE,
    3: "config 3: 1024 instances (32x32 grid) x icosphere subdiv6 (81920 tris), shared 1024^2 noise, frustum + HiZ instance culling, 4 px/segment",
    4: "config 4: icosphere subdiv8, every edge factor in (11, 24] -> triangle_split on every triangle, ~600M output tris",
    5: "config 5: 64 instances x icosphere subdiv7 (327680 tris) far field, factors <= 2: 1X + 2X transient builds",
}


def make(key: int, small: bool = False) -> Workload:
    """`small`: the same shape at a size the CPU oracle finishes in seconds (debugging / CPU-side tests)."""
    if key == 1:
        s, f = S.config_plane(64 if small else 256, tex_size=128 if small else 512)
        cfg = api.Config(numVisibleClusterBits=12, numPartTriangleBits=18, numSplitTriangleBits=12, numGeneratedVerticesBits=24, numGeneratedClusterMegs=1024)
        return Workload(1, NAMES[1], s, f, cfg)
    if key == 2:
        subdiv, tex = (5, 256) if small else (8, 2048)
        s, f = S.config_icosphere(subdiv, tex_size=tex, distance=1.5, tess_rate_pixels=0.75)
        cfg = api.Config(numVisibleClusterBits=12 if small else 15, numPartTriangleBits=16 if small else 22, numSplitTriangleBits=14 if small else 20,
                         numGeneratedVerticesBits=22 if small else 27, numGeneratedClusterMegs=4095)
        return Workload(2, NAMES[2] if not small else "small icosphere (debug)", s, f, cfg)
    if key == 3:
        n, subdiv, tex = (64, 4, 128) if small else (1024, 6, 1024)
        s, f, pyr, size, mips = S.config_instances(n, subdiv=subdiv, tex_size=tex, tess_rate_pixels=4.0)
        cfg = api.Config(flags=api.FLAG_DEFAULT | api.FLAG_CULLING, numVisibleClusterBits=14 if small else 21, numPartTriangleBits=20 if small else 23,
                         numSplitTriangleBits=16 if small else 21, numGeneratedVerticesBits=25 if small else 28, numGeneratedClusterMegs=16000)
        return Workload(3, NAMES[3], s, f, cfg, (pyr, size, mips))
    if key == 4:
        s, f = S.config_split_stress(5 if small else 8, 256 if small else 2048)
        cfg = api.Config(numVisibleClusterBits=15, numPartTriangleBits=18 if small else 24, numSplitTriangleBits=15 if small else 21,
                         numGeneratedVerticesBits=24 if small else 29, numGeneratedClusterMegs=16000)
        return Workload(4, NAMES[4], s, f, cfg)
    if key == 5:
        s, f = S.config_far_field(8 if small else 64, subdiv=5 if small else 7, tex_size=256 if small else 2048)
        cfg = api.Config(numVisibleClusterBits=19, numPartTriangleBits=20 if small else 23, numSplitTriangleBits=12, numGeneratedVerticesBits=24 if small else 28,
                         numGeneratedClusterMegs=16000)
        return Workload(5, NAMES[5], s, f, cfg)
    raise ValueError(f"BASELINE config {key} does not exist (1..5)")


def place_on_ring(w: Workload, rank: int, world: int) -> Workload:
    """Weak scaling of a single-instance workload: rank r gets its own copy of the instance, rotated about the eye so
    that every rank sees a statistically identical frame (the tess metric depends on eye distance and edge length only)."""
    if world > 1:
        eye = w.frame_constants[0]["viewPos"][:3].astype(np.float64)
        ang = 2 * np.pi * rank / world
        c, s = np.cos(ang), np.sin(ang)
        rot = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        centre = eye + rot @ (-eye)  # instance 0 is at the origin
        w.scene.instances[0]["worldMatrix"] = S.translation(centre).T.reshape(16)
    return w


def setup(w: Workload, table, device: int = 0):
    """A ready context for the workload."""
    w.config.device = device
    gpu = api.TessClusters(w.config)
    gpu.set_tess_table(table)
    gpu.set_scene(w.scene)
    if w.hiz is not None:
        gpu.set_hiz(*w.hiz)
    return gpu
