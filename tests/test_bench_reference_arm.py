"""bench.py --impl reference (the CPU arm the driver runs next to the GPU arm) prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--n-fixed", "400", "--trials", "64"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "registrations/sec" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["unit"] == "registrations/s" and line["n_gpus"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    e2e = line["e2e"]
    assert e2e["value"] == line["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
    # both arms describe the workload with the same keys and values (the driver compares the two `config` dicts)
    sys.path.insert(0, ROOT)
    import bench
    assert line["config"] == bench.workload_config(360, 400, 64)
    assert cb["full_run_s"] > 0 and set(cb["full_run_stages_s"]) == {"mean_distance", "descriptors", "chi2", "lap", "ransac", "icp"}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--n-fixed", "400"], capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_specimens_are_posed_not_distorted():
    """The all-pairs specimens (BASELINE config 5) differ by pose and overall size: the first PCA axis of a specimen is
    the image of the atlas' axis to a few degrees for every pair (the fully anisotropic random affine swaps it in about
    half of the pairs, and the method - the reference's as well - cannot register those), and `vary=0` gives exactly
    equal counts."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from platymatch_b200.synthetic import make_specimens
    specs = make_specimens(6, 2000, seed=0)
    ax = [O.pca_first_axis(s["points"].T) for s in specs]
    worst = 0.0
    for i in range(6):
        for j in range(i + 1, 6):
            t = specs[j]["A"] @ np.linalg.inv(specs[i]["A"])
            a = t[:3, :3] @ ax[i]
            c = abs(a @ ax[j]) / np.linalg.norm(a) / np.linalg.norm(ax[j])
            worst = max(worst, np.degrees(np.arccos(min(1.0, c))))
    assert worst < 12.0, worst
    assert len({s["points"].shape[1] for s in specs}) > 1
    assert {s["points"].shape[1] for s in make_specimens(4, 500, seed=1, vary=0.0)} == {500}
