"""bench.py's output contract on the arm that runs without a GPU (--impl reference: the CPU port of the path).
One JSON line with the keys the driver reads; the GPU arm's line carries the same keys plus roofline/phases."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "2^16_vs_2^8",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "items/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["gpu_launches"] == 0 and d["value"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "items/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the cold figure (plaintext transforms inside run(), the reference's one-query-per-process case) rides along
    assert d["cpu_baseline"]["cold"]["value"] > 0   # no ordering check against the warm figure: one step each on shared cores is too noisy


def test_reference_arm_under_torchrun_env_uses_all_cores_and_never_maps_the_product_library():
    """torch.distributed.run exports OMP_NUM_THREADS=1 and WORLD_SIZE; the arm must still use every host core, keep
    the GPU arm's config (strong scaling: one query whatever N is) and must not load libpsi_b200.so."""
    env = dict(os.environ, OMP_NUM_THREADS="1", WORLD_SIZE="4", RANK="0", LOCAL_RANK="0")
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--workload', '2^16_vs_2^8', '--gpus', '4', "
            "'--steps', '1', '--warmup', '1']; runpy.run_path(%r, run_name='__main__'); "
            "maps = open('/proc/self/maps').read(); print('PRODUCT_MAPPED' if 'libpsi_b200' in maps else 'PRODUCT_NOT_MAPPED'); "
            "print('ORACLE_MAPPED' if 'libpsi_oracle' in maps else 'ORACLE_NOT_MAPPED')" % os.path.join(ROOT, "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "PRODUCT_NOT_MAPPED" in r.stdout and "ORACLE_MAPPED" in r.stdout
    d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["n_gpus"] == 4 and d["scaling"] == "strong"
    assert d["config"]["server_items_total"] == 1 << 16 and d["config"]["bins_per_gpu"] == "8/4"
    # ranks other than 0 print nothing and exit 0
    env["RANK"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "2^16_vs_2^8",
                        "--gpus", "4", "--steps", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)
