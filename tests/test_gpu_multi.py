"""Several GPUs behind one context (psy_init with n_devices > 1; `psychopath --gpus N`): the buckets are shared
between the GPUs through a common counter and the tiles gathered onto the first device with NCCL.  Needs two visible
GPUs (skipped on a one-GPU box; the driver's scaling run and `gpurun --gpus 2` exercise it)."""
import os
import subprocess

import numpy as np
import pytest

from psychopath_b200 import api
from test_gpu_parity import make_scene

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.skipif(n_gpus() < 2, reason="needs two GPUs")
def test_two_gpus_in_one_context_render_the_one_gpu_image():
    scene = api.Scene.synth_instanced(asset_quads=60, n_assets=4, n_clusters=30, inst_per_cluster=8, lights_per_cluster=6,
                                      cluster_xform_keys=2, res=(480, 270), spp=16, seed=5, motion_pct=1250)
    one = api.Renderer(scene, device=0)
    a, sa = one.render()
    two = api.Renderer(scene, device=(0, 1))
    b, sb = two.render(max_samples_per_bucket=1 << 17)  # 16 buckets shared by two GPUs
    assert np.array_equal(a, b, equal_nan=True)
    assert sa["ray_count"] == sb["ray_count"] and sa["samples"] == sb["samples"]
    assert sb["buckets"] >= 16
    c, sc = two.render()  # default bucket size for two GPUs
    assert np.array_equal(a, c, equal_nan=True)
    # probe rays and single buckets go to the first device of the context
    bx, _ = two.render_bucket(8, 8, 16, 16)
    assert np.array_equal(bx, a[8:24, 8:24])


@pytest.mark.skipif(n_gpus() < 2, reason="needs two GPUs")
def test_cli_gpus_flag(tmp_path):
    import cv2
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(repo, "bin", "psychopath")
    scene = os.path.join(repo, "tests", "golden", "cornell_box.psy")
    (tmp_path / "test_renders").mkdir()
    imgs = []
    for flags in (["--gpus", "1"], ["--gpus", "2"], ["--device-list", "1,0"]):
        out = subprocess.run([exe, "-i", scene, "--spp", "4", "--spb", "16384", "--counters"] + flags, cwd=tmp_path,
                             capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stderr
        assert "Per ray:" in out.stdout
        imgs.append(cv2.imread(str(tmp_path / "test_renders" / "cornell_box.png")))
    assert np.array_equal(imgs[0], imgs[1]) and np.array_equal(imgs[0], imgs[2])


def test_cli_probe_rays(tmp_path):
    """`psychopath --probe-rays FILE`: the closest-hit parity entry from the command line."""
    import oracle_lib as O
    from psychopath_b200 import abi
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(repo, "bin", "psychopath")
    path = os.path.join(repo, "tests", "golden", "cornell_box.psy")
    scene = api.Scene.from_psy_file(path)
    rays = O.camera_rays(scene.desc, 100, 100, 32, 32, 2, 1)
    rays.tofile(tmp_path / "rays.bin")
    out = subprocess.run([exe, "-i", path, "--probe-rays", str(tmp_path / "rays.bin")], cwd=tmp_path, capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    hits = np.fromfile(tmp_path / "rays.bin.hits", dtype=abi.HIT_DTYPE)
    oh, _, _ = O.trace_rays(scene.desc, rays)
    assert len(hits) == len(rays)
    for k in ("kind", "prim_id", "inst_id", "t"):
        assert np.array_equal(hits[k], oh[k]), k


def test_cli_renders_a_synthetic_baseline_scene(tmp_path):
    """`psychopath --synth NAME [--gpus N]`: BASELINE.json's synthetic configurations from the command line (C5 is the
    4K motion-blurred instanced scene of the north star; `tiny` has its structure at test size).  The picture must be
    the one the API renders for the same scene, on one GPU and — when there are two — on two."""
    import cv2
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(repo, "bin", "psychopath")
    runs = [["--gpus", "1"]] + ([["--gpus", "2"]] if n_gpus() >= 2 else [])
    imgs = []
    for flags in runs:
        out = subprocess.run([exe, "--synth", "tiny", "--counters"] + flags, cwd=tmp_path, capture_output=True, text=True,
                             timeout=300)
        assert out.returncode == 0, out.stderr
        assert "two time keys" in out.stdout
        imgs.append(cv2.imread(str(tmp_path / "synth_instanced.png")))
    assert imgs[0] is not None and imgs[0].shape[:2] == (180, 320) and imgs[0].mean() > 2
    for im in imgs[1:]:
        assert np.array_equal(imgs[0], im)
    scene = api.Scene.synth_instanced(asset_quads=40, n_assets=4, n_clusters=16, inst_per_cluster=6, lights_per_cluster=4,
                                      cluster_xform_keys=2, res=(320, 180), spp=8, seed=5, motion_pct=1250)
    xyz, _ = api.Renderer(scene).render()
    m = np.array([[2.6896551724, -1.2758620690, -0.4137931034], [-1.0221081721, 1.9782866167, 0.0438215555],
                  [0.0612244898, -0.2244897959, 1.1632653061]], dtype=np.float32)
    rgb = np.clip(xyz @ m.T, 0, None)
    srgb = np.where(rgb < 0.0031308, rgb * 12.92, 1.055 * np.power(np.maximum(rgb, 1e-12), 1 / 2.4) - 0.055)
    want = (np.clip(srgb, 0, 1) * 255).astype(np.int32)
    assert np.abs(imgs[0][..., ::-1].astype(np.int32) - want).mean() < 0.5
