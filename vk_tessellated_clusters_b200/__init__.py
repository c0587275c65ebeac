"""B200-native per-frame adaptive tessellation path of nvpro-samples/vk_tessellated_clusters.

Product = CUDA kernels + C ABI in ``csrc/`` (libtess_clusters.so), driven through ``api.TessClusters``.
"""
from .api import Config, TessClusters, TessError  # noqa: F401
from .table import load_tess_table  # noqa: F401
