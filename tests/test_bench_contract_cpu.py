"""CPU: the driver-facing contract of bench.py / run_test.py that can be checked without a GPU -- the reference arm prints
exactly ONE JSON line on stdout with the keys the driver reads, and the product arms refuse to run (loudly, non-zero exit)
when there is no CUDA device instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*argv, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, *argv], cwd=ROOT, capture_output=True, text=True, timeout=600, env=e)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = _run("bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0")     # one pair of the full path through the oracle port
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [x for x in p.stdout.splitlines() if x.strip()]
    assert len(lines) == 1, p.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "image-pairs/sec" and line["unit"] == "pairs/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1 and line["warmup"] == 0
    assert abs(line["ms_per_step"] * line["steps"] * 1e-3 * line["value"] - line["steps"]) < 1e-6      # the printed step time is the time spent
    assert line["status"]["ok"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    assert line["cpu_baseline"]["value"] == line["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]
    # both arms print the SAME config (the driver compares them)
    import importlib.util
    spec = importlib.util.spec_from_file_location("oryon_bench_cfg", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert line["config"] == bench.workload_config()


def test_reference_arm_under_a_multi_rank_launch_only_rank0_works():
    p = _run("bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", env={"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.parametrize("script", ["bench.py", "run_test.py"])
def test_product_arms_fail_loudly_without_a_gpu(script):
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    p = _run(script)
    assert p.returncode != 0 and "no CUDA device" in (p.stderr + p.stdout) and "no CPU fallback" in (p.stderr + p.stdout)
    assert p.stdout.strip() == "", "nothing that could be mistaken for a result line"


def test_eager_baseline_helper_matches_the_oracle_formulation():
    """`bench.py --eager-baseline` times the reference's formulation as PyTorch eager ops; on a tiny CPU case the helper runs and its
    arithmetic is the oracle's (same broadcast cosine / amin / argmin)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("oryon_bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.eager_torch_sample("cpu", torch.float32, 48, D=16, n=200, row_chunk=20, repeats=1) > 0
    import oryon_oracle as oracle
    g = torch.Generator().manual_seed(0)
    f1, f2 = torch.randn(48, 16, generator=g), torch.randn(200, 16, generator=g)
    d = 0.5 * (-1 * torch.nn.functional.cosine_similarity(f1.unsqueeze(1), f2.unsqueeze(0), dim=2) + 1)
    md, mi = oracle.match_rows(f1, f2)
    assert torch.equal(torch.amin(d, 1), md) and torch.equal(torch.argmin(d, 1), mi)
