"""bench.py --impl reference on CPU: the contract of the reference arm's JSON line.  It times the CPU path on OUR arm's
job, so `config`, `metric`, `unit` and `higher_is_better` must be the objects our arm prints; under torchrun only rank 0
runs.  (The arm executes oracle/: it is the one place besides the cpu_baseline leg where bench.py may.)"""
import json
import os
import subprocess
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None, gpus=1):
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(gpus), "--rows", "400000",
                        "--steps", "2", "--warmup", "1", "--no-ref-engine"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    return [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_line_names_our_arms_job():
    sys.path.insert(0, ROOT)
    import bench as B
    lines = _run(gpus=2)
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "ms" and d["higher_is_better"] is False
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 2
    assert d["value"] > 0 and d["value"] == d["e2e"]["value"] == d["cpu_baseline"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    # the same `config` our arm builds for this job at N = 2 (rank 0 holds the first half of the chunks)
    args = SimpleNamespace(rows=400000, sf=100.0, workers=4)
    from quickstep_b200 import synth as S
    shape = S.db_shape(400000)
    l_hi = S.chunk_rows(shape, S.rank_chunks(shape, 2, 0)[-1])[1][1]
    assert d["config"] == B.bench_config(args, 400000, 2, l_hi, B.n_host_workers(args, 2))
    assert d["result_check"]["q1_groups"] == 4


def test_reference_arm_other_ranks_print_nothing():
    assert _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, gpus=2) == []
