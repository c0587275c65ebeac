"""Small seeded cases shared by the oracle tests (CPU) and the parity tests (GPU)."""
import numpy as np

from vk_tessellated_clusters_b200 import api, scenes as S

SMALL = dict(numVisibleClusterBits=12, numPartTriangleBits=16, numSplitTriangleBits=12, numGeneratedVerticesBits=22)


def case(name):
    """-> (scene, frame constants pair, Config, hiz tuple or None)"""
    F = api
    if name == "plane":  # factors sweep up to 11, no split
        s, f = S.config_plane(32, tex_size=64)
        return s, f, F.Config(**SMALL), None
    if name == "plane_ragged":  # 37x37 quads: ragged clusters (fewer than 64 triangles, odd vertex counts)
        s, f = S.config_plane(37, tex_size=64)
        return s, f, F.Config(**SMALL), None
    if name == "split":  # factors up to 40: one split level
        s, f = S.config_plane(32, tex_size=64, max_factor=40.0)
        return s, f, F.Config(**SMALL), None
    if name == "deep_split":  # factors up to 1500: three split levels
        s, f = S.config_plane(8, tex_size=64, max_factor=1500.0)
        return s, f, F.Config(numVisibleClusterBits=8, numPartTriangleBits=20, numSplitTriangleBits=18, numGeneratedVerticesBits=26, numGeneratedClusterMegs=4095), None
    if name == "mini":  # factors <= 2: 1X subsets + 2X mini batches + full clusters
        s, f = S.config_plane(32, tex_size=64, max_factor=2.3)
        return s, f, F.Config(**SMALL), None
    if name == "full":  # almost everything factor 1
        s, f = S.config_plane(32, tex_size=64, max_factor=1.2)
        return s, f, F.Config(**SMALL), None
    if name == "undisplaced":
        s, f = S.config_plane(32, tex_size=64, displaced=False)
        return s, f, F.Config(**SMALL), None
    if name == "linear_no_transient":  # TESS_USE_PN 0, 1X/2X off: plain partTriangleCounter path
        s, f = S.config_plane(32, tex_size=64, max_factor=14.0)
        return s, f, F.Config(flags=0, **SMALL), None
    if name == "only_1x":
        s, f = S.config_plane(32, tex_size=64, max_factor=2.3)
        return s, f, F.Config(flags=F.FLAG_PN_DISPLACEMENT | F.FLAG_TRANSIENT_1X, **SMALL), None
    if name == "only_2x":
        s, f = S.config_plane(32, tex_size=64, max_factor=2.3)
        return s, f, F.Config(flags=F.FLAG_PN_DISPLACEMENT | F.FLAG_TRANSIENT_2X, **SMALL), None
    if name == "animation":
        s, f = S.config_plane(16, tex_size=64)
        f[0]["animationState"] = 0.37
        return s, f, F.Config(flags=F.FLAG_DEFAULT | F.FLAG_ANIMATION, **SMALL), None
    if name == "icosphere":
        s, f = S.config_icosphere(4, tex_size=128)
        return s, f, F.Config(**SMALL), None
    if name == "far_field":  # 9 instances, factors in {1,2}
        s, f = S.config_far_field(9, subdiv=4, tex_size=128)
        return s, f, F.Config(**SMALL), None
    if name == "culling":  # instance grid with frustum + HiZ culling: hidden instances are emitted untessellated
        s, f, pyr, size, mips = S.config_instances(36, subdiv=3, tex_size=64, tess_rate_pixels=1.0)
        return s, f, F.Config(flags=F.FLAG_DEFAULT | F.FLAG_CULLING, **SMALL), (pyr, size, mips)
    if name == "split_factor_4":
        s, f = S.config_plane(16, tex_size=64, max_factor=90.0)
        return s, f, F.Config(splitFactor=4, **SMALL), None
    if name == "overflow_parts":  # part list too small: drop + counters keep counting
        s, f = S.config_plane(32, tex_size=64)
        return s, f, F.Config(numVisibleClusterBits=12, numPartTriangleBits=10, numSplitTriangleBits=12, numGeneratedVerticesBits=22), None
    if name == "overflow_vertices":
        s, f = S.config_plane(32, tex_size=64)
        return s, f, F.Config(numVisibleClusterBits=12, numPartTriangleBits=16, numSplitTriangleBits=12, numGeneratedVerticesBits=14), None
    if name == "overflow_split":
        s, f = S.config_plane(32, tex_size=64, max_factor=40.0)
        return s, f, F.Config(numVisibleClusterBits=12, numPartTriangleBits=16, numSplitTriangleBits=8, numGeneratedVerticesBits=22), None
    if name == "overflow_clas_data":
        s, f = S.config_plane(32, tex_size=64)
        return s, f, F.Config(numGeneratedClusterMegs=1, **SMALL), None
    if name == "overflow_visible":
        s, f = S.config_plane(64, tex_size=64, max_factor=3.0)
        return s, f, F.Config(numVisibleClusterBits=6, numPartTriangleBits=16, numSplitTriangleBits=12, numGeneratedVerticesBits=22), None
    if name == "overflow_transient":  # dual-ended part list collides: front parts vs transient meta from the back
        s, f = S.config_plane(32, tex_size=64, max_factor=2.6)
        return s, f, F.Config(numVisibleClusterBits=12, numPartTriangleBits=9, numSplitTriangleBits=12, numGeneratedVerticesBits=22), None
    raise KeyError(name)


ALL_CASES = ["plane", "plane_ragged", "split", "deep_split", "mini", "full", "undisplaced", "linear_no_transient", "only_1x", "only_2x", "animation",
             "icosphere", "far_field", "culling", "split_factor_4", "overflow_parts", "overflow_vertices", "overflow_split", "overflow_clas_data",
             "overflow_visible", "overflow_transient"]
