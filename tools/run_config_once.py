import sys; sys.path.insert(0,'.')
import tools.bench_configs as B
from vk_tessellated_clusters_b200 import api, table as T
import sys
k=int(sys.argv[1])
name, scene, fcs, cfg, hiz = B.configs()[k]()
tbl=T.load_tess_table()
gpu=api.TessClusters(cfg); gpu.set_tess_table(tbl); gpu.set_scene(scene)
if hiz: gpu.set_hiz(*hiz)
for _ in range(3): gpu.frame(fcs)
gpu.sync()
