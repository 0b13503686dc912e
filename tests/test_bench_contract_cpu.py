"""bench.py's contract where it can be checked without a GPU: the reference arm (CPU port on host
cores) prints exactly one JSON line with the agreed keys; the default arm refuses to run without
a CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

from helpers import ROOT


def test_sequence_state_follows_the_reference_schedule():
    sys.path.insert(0, ROOT)
    import bench
    mf = [0, 100, 200]
    assert bench.sequence_state(0, mf) == (0, 1, 1.0, 0.0)
    cur, nxt, w0, w1 = bench.sequence_state(150, mf)
    assert (cur, nxt) == (1, 2) and abs(w0 - 0.5) < 1e-12 and abs(w1 - 0.5) < 1e-12     # :84-87
    assert bench.sequence_state(100, mf)[:2] == (1, 2)                                   # switch at n_frame (:117)
    assert bench.sequence_state(250, mf) == (2, -1, 1.0, 0.0)                            # last model: no blend
    w, h = bench.sample_dims(33000)
    assert w % 16 == 0 and h % 16 == 0 and 16384 <= w * h <= 40000


def test_reference_arm_prints_one_json_line():
    env = dict(os.environ, PCM_REF_CORES="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--no-sweep"],
                         capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout[-1000:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "PC masker frames/sec at 1080p" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 2 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_default_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                         timeout=300, cwd=ROOT)
    assert res.returncode != 0 and "no CUDA device" in (res.stderr + res.stdout)
    assert not res.stdout.strip(), "nothing may be reported without a GPU"


def test_cpu_sequence_port_runs_like_main_py(tmp_path):
    """oracle/ref_sequence.py (the CPU arm of the sweep that bench.py samples): `script CONFIG RESULT [frames]` writes the
    reference's "{mean IoU};{seconds}" result line (main.py:366-368) plus the timing detail bench.py extrapolates from."""
    import yaml
    from helpers import PKG, polygons
    from pcm import sweep
    base = yaml.full_load(open(os.path.join(PKG, "config_benchmark.yaml")))
    prm = dict(n_estimators=20, max_depth=7, n_components=1, novelty_detection=False, over_segmentation="felzenszwalb",
               features="6 lab", dilation_kernel=7, prior_weight=0.1)
    cfg = sweep.sequence_config(base, polygons(), "soldier", prm, os.path.join(PKG, "Input/SegTrack2/Video"),
                                os.path.join(PKG, "Input/SegTrack2/Truth"))
    cp, op = tmp_path / "config-0-soldier.yaml", tmp_path / "results-0-soldier.csv"
    yaml.dump(cfg, open(cp, "w"))
    res = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_sequence.py"), str(cp), str(op), "2"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    l1, l2 = open(op).read().split("\n")
    iou, secs = map(float, l1.split(";"))
    assert 0.3 < iou <= 1.0 and secs > 0
    t_imp, t_train, n = l2.split(";")
    assert float(t_imp) > 0 and float(t_train) > 0 and int(n) == 2


def test_committed_traffic_capture_is_well_formed():
    """bench.py copies profiles/r02_traffic.json (tools/ncu_traffic.py over an ncu capture of the same command) into
    roofline.traffic / traffic_by_kernel: all five kernels of the chain, per-launch DRAM and L2 bytes, the cold-cache
    figure of the --set full capture for K1."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tr = json.load(open(os.path.join(root, "profiles", "r02_traffic.json")))
    assert set(tr["kernels"]) == {"planes", "score", "segment_decide", "mask_dilate", "iou"}
    for k, v in tr["kernels"].items():
        assert v["launches"] >= 10 and v["dram_bytes"] >= 0 and v["lts_bytes"] > 0 and v["duration_us"] > 0, k
    npx = 1920 * 1080
    assert tr["kernels"]["score"]["algorithmic_bytes"] == 11 * npx
    assert tr["kernels"]["score"]["dram_bytes_cold_cache"] > 7 * npx          # planes + labels from a flushed L2
    assert tr["frame"]["compulsory_bytes"] == 10 * npx
