"""CPU: the reference arm of bench.py (the reference's own Training._run_batch from oracle/_ref when that copy exists,
else the oracle port, timed on the host cores) prints ONE JSON line with the contract's keys.  The GPU arm's line is checked by the driver on a B200; here only that the shared schema cannot rot unnoticed."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, timeout=600, cwd=ROOT).stdout.decode()
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out[-2000:]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "elbo_grad_trajectories_per_sec" and d["unit"] == "traj/s"
    assert d["config"]["workload"] == "dr_constant_icml" and d["value"] > 0 and d["higher_is_better"] is True
    have_ref = os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "vi-hds", "vihds")) or os.path.isdir("/root/reference/vihds")
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    for k in ("workload", "spec", "batch_per_gpu", "global_batch", "iw", "trajectories_per_step", "T", "state_width", "n_theta",
              "solver", "parallelism"):  # the same keys as the CUDA arm's config (bench.workload_config)
        assert k in d["config"], k
    assert d["config"]["batch_per_gpu"] == 36 and d["config"]["iw"] == 200 and d["config"]["T"] == 86
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
