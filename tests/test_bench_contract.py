"""bench.py contract pieces that can be checked without a GPU: the reference arm's JSON line, rank handling under torchrun,
and that the product arm fails loudly (no CPU fallback) when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "1", "--n-iq", "1020000"])   # short streams: CPU-suite budget
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "MS/s" and d["higher_is_better"] is True and d["scaling"] == "strong"
    assert d["metric"] == "calibrated IQ MSamples/s" and d["dtype"] == "f64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm states what it ran: whole streams of the configured length, `cores` of them per step, and the runtime probe
    assert d["config"]["iq_per_stream"] == 1020000 and "whole streams" in cb["sample"] and str(1020000) in cb["sample"]
    assert set(d["reference_runtime_probe"]["probed"]) >= {"octave", "matlab"}


def test_reference_arm_default_config_is_config5():
    sys.path.insert(0, ROOT)
    import bench
    c = bench.workload_config(8)
    assert c["streams_total"] == 1024 and c["streams_per_gpu"] == 128 and c["iq_per_stream"] == 21666667 and c["workload"].startswith("BASELINE config 5:")
    c1 = bench.workload_config(1)
    assert c1["streams_total"] == 1024 and c1["streams_per_gpu"] == 1024
    assert bench.outcome_of(-1, -1, -1, 0) == "no_fcch_found" and bench.outcome_of(213, 212, 594, 0) == "calibrated"
    assert bench.outcome_of(213, -1, -1, 2) == "fine_snr_gate" and bench.outcome_of(213, 212, -1, 4) == "sch_edge_abort"


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a CUDA device")
def test_product_arm_fails_loudly_without_a_gpu():
    p = _run(["--steps", "1", "--warmup", "3", "--streams", "2", "--n-iq", "1020000", "--no-cpu-baseline", "--no-e2e"], timeout=300)
    assert p.returncode != 0
    assert not any(ln.startswith("{") and '"value"' in ln for ln in p.stdout.splitlines())
