"""bench.py contract checks that need no GPU: CLI defaults, the JSON keys of the reference arm on a tiny sample, and
the clock-sampler window logic."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_cli_defaults():
    import bench
    old = sys.argv
    sys.argv = ["bench.py"]
    try:
        a = bench.parse_args()
    finally:
        sys.argv = old
    assert a.gpus == 1 and a.steps >= 3 and a.warmup >= 3 and a.impl == "b200" and a.image_size == 64


def test_reference_arm_line_has_contract_keys():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                          "--batch", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "train image-pairs/sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_idle_on_nonzero_rank():
    import os
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3",
                          "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_clock_sampler_window():
    import bench
    s = bench.ClockSampler(0)
    s.proc = subprocess.Popen([sys.executable, "-c", "pass"])       # stands in for nvidia-smi
    s.lines = [(1.0, "1000, 1965, 300, Not Active, Not Active, Not Active, Not Active"),
               (2.0, "1900, 1965, 900, Not Active, Not Active, Not Active, Active"),
               (2.5, "1950, 1965, 900, Not Active, Not Active, Not Active, Not Active"),
               (9.0, "500, 1965, 100, Active, Not Active, Not Active, Not Active")]
    s.t0, s.t1 = 1.5, 3.0
    c = s.stop()
    assert c["samples"] == 2 and c["sm_mhz"] in (1900.0, 1950.0) and c["sm_max_mhz"] == 1965.0
    assert c["reasons"] == ["sw_power_cap"]


def test_product_and_b200_arm_never_import_the_oracle():
    """oracle/ is test infrastructure: only tests/, smoke() and the CPU legs of bench.py may touch it."""
    import ast
    tree = ast.parse((ROOT / "bench.py").read_text())
    allowed = {"cpu_reference_pairs_per_s"}                      # cpu_baseline leg and --impl reference
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        mods = [n.module for n in ast.walk(fn) if isinstance(n, ast.ImportFrom) and n.module] + \
               [a.name for n in ast.walk(fn) if isinstance(n, ast.Import) for a in n.names]
        if any(m.split(".")[0] == "oracle" for m in mods):
            assert fn.name in allowed, f"bench.py:{fn.name} imports oracle"
    for py in (ROOT / "discogan_modernized_b200").glob("*.py"):
        t = ast.parse(py.read_text())
        mods = [n.module for n in ast.walk(t) if isinstance(n, ast.ImportFrom) and n.module] + \
               [a.name for n in ast.walk(t) if isinstance(n, ast.Import) for a in n.names]
        assert not any(m.split(".")[0] == "oracle" for m in mods), f"{py.name} imports oracle"
