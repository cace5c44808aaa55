"""CPU checks of the bench.py contract that do not need a GPU: the reference arm (the oracle port timed on
the host cores) prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--points", "6000",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["unit"] == "correspondences/s" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and line["steps"] == 1 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    # value = the loop (kd-tree build + iterations, normals given), e2e = the whole RegisterSymm incl. the k=10 normals of
    # both clouds: each is like for like with the same key of the GPU arm (round-1 verdict, "ratio compares unlike regions")
    e2e = line["e2e"]
    assert 0 < e2e["value"] < line["value"] and e2e["unit"] == line["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_both_arms_name_the_same_workload():
    """the driver compares the arms' config.workload strings (same_config)"""
    import bench
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"workload": workload_name(n)') == 2
    assert "config 3" in bench.workload_name(1_000_000)


def test_default_arm_fails_loudly_without_a_gpu():
    """on a box without CUDA the product arm must not fall back to the CPU: it exits non-zero"""
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--points", "6000", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode != 0
    assert "reference" not in out.stdout


def test_pipelined_e2e_is_an_extra_not_the_headline():
    """`e2e.value` stays the strictly serial call (upload, index, normals, loop, read-back, one registration at a time);
    the stream of registrations over two contexts is reported beside it under `e2e.pipelined`"""
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert '"e2e": {"value": tot_e2e_corr / (max_e2e_ms * 1e-3)' in src
    assert '"pipelined": None if pipe_ms is None else' in src
    assert src.index("my_e2e_ms = float(np.sum(e2e_ms))") > src.index("pipe_ms = pe0.elapsed_time(pe1) / k_pipe")   # separate clocks
