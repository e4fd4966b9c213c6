import sys, os
sys.path.insert(0, '/root/repo')
import bench
from restir_embree_b200 import scenes
from restir_embree_b200.renderer import Renderer
sc = scenes.scene_config("1m")
r = Renderer(1920, 1080, device=0, seed=123)
r.upload_scene(sc); r.set_params(bench.bench_params())
for f in range(12):
    r.render_frame_device(bench.camera_at(sc, f), f)
r.synchronize(); r.close()
