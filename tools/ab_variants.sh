#!/bin/bash
# usage: tools/ab_variants.sh <tag> ...   for every psychopath_b200/libpsy_cuda_<tag>.so ("base" = the product library):
# image identical to the product library's on a mid-size scene, then bench C5 at $AB_SPP (default 16) spp
cat > /tmp/ab_img.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from psychopath_b200 import api
s = api.Scene.synth_instanced(asset_quads=200, n_assets=4, n_clusters=40, inst_per_cluster=8, lights_per_cluster=6,
                              cluster_xform_keys=2, res=(480, 270), spp=8, seed=11, motion_pct=1250)
a, st = api.Renderer(s).render()
g = api.Scene.synth_displaced_grid(96, n_vert_keys=4, n_xform_keys=4, res=(240, 135), spp=8, seed=3)
b, st2 = api.Renderer(g).render()
np.save(sys.argv[1], np.concatenate([a.reshape(-1), b.reshape(-1), [st['ray_count'], st2['ray_count']]]))
PY
for tag in "$@"; do
  lib=psychopath_b200/libpsy_cuda_$tag.so; [ "$tag" = base ] && lib=psychopath_b200/libpsy_cuda.so
  PSY_CUDA_LIB=$PWD/$lib python /tmp/ab_img.py /tmp/ab_$tag.npy 2>/dev/null
  same=$(python -c "
import numpy as np; a=np.load('/tmp/ab_base.npy'); b=np.load('/tmp/ab_$tag.npy'); print(np.array_equal(a,b,equal_nan=True))")
  out=$(PSY_CUDA_LIB=$PWD/$lib timeout 600 python bench.py --spp ${AB_SPP:-16} --steps 3 --warmup 2 --no-cpu-baseline --no-c3 2>/dev/null | tail -1)
  echo "$out" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step_rank0']
print('$tag', 'same_image=$same', 'Mrays/s=%.1f'%d['value'], 'trace=%.1f'%s['trace_ms'], 'shade=%.1f'%s['shade_ms'], 'ms/step=%.1f'%d['ms_per_step'])
"
done
