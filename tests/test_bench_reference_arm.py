"""``bench.py --impl reference`` (the CPU arm the driver runs beside the GPU arm): the JSON contract of its line, and
that under a multi-rank launch only rank 0 works and prints.  Small workload; no GPU involved."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env, *args):
    env = dict(os.environ, **extra_env)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        if k not in extra_env:
            env.pop(k, None)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--workload", "config2_1k_x_30"] + list(args), env=env, capture_output=True, text=True, timeout=300)


def test_reference_arm_line():
    res = _run({})
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                       # ONE JSON line on stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["metric"] == "em_iteration_cell_site_per_s" and line["unit"] == "cell*site/s" and line["higher_is_better"] is True
    assert line["config"]["workload"] == "config2_1k_x_30" and line["dtype"] == "f64" and line["vs_baseline"] is None
    cpu, e2e = line["cpu_baseline"], line["e2e"]
    assert cpu["kind"] in ("port", "reference") and cpu["cores"] >= 1 and cpu["sample"] and cpu["value"] == line["value"]
    assert e2e == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and line["ms_per_step"] > 0


def test_reference_arm_under_torchrun_env_uses_every_core_and_no_gpu_library():
    """torchrun exports OMP_NUM_THREADS=1 to every rank: the arm must still use the cores the process may run on
    (round 1's N >= 2 lines ran on one thread); and it must not touch the GPU library at all (PMM_LIB points nowhere)"""
    res = _run({"RANK": "0", "LOCAL_RANK": "0", "WORLD_SIZE": "2", "OMP_NUM_THREADS": "1", "PMM_LIB": "/nonexistent/libpmm_b200.so"},
               "--gpus", "2")
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip())
    assert line["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and line["n_gpus"] == 2
    assert "libpmm_b200" not in res.stderr


def test_both_arms_print_the_same_config():
    sys.path.insert(0, ROOT)
    import bench
    from laml_b200.synth import random_priors
    spec = bench.workload_spec("config2_1k_x_30")
    qptr, _ = random_priors(spec["n_sites"], spec["a_lo"], spec["a_hi"], seed=5, uniform=True)
    ref_cfg = json.loads(_run({}).stdout.strip())["config"]
    assert ref_cfg == bench.config_dict("config2_1k_x_30", spec, qptr, 1)       # run_ours prints exactly config_dict(...)
    import inspect
    src = inspect.getsource(bench.run_ours)
    assert '"config": config_dict(args.workload, spec, qptr, world),' in src


def test_reference_arm_other_ranks_exit_quietly():
    res = _run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2")
    assert res.returncode == 0, res.stderr[-2000:]
    assert res.stdout.strip() == ""


def test_time_em_step_helper():
    """bench.time_em_step (the ``e2e_em_step`` leg of the GPU arm) on the oracle-backed stand-in engine: a time per
    step when the calls work, the exception's repr -- not a raise -- when they do not"""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from laml_b200.dist import ShardedEngine
    from laml_b200.synth import make_workload
    from test_sharded_gloo import OracleBackedEngine
    w = make_workload(40, 10, 1, 8, 0.2, 0.15, seed=4)
    eng = OracleBackedEngine(w.forest, w.codes, w.qptr, w.qval, 40)
    sh = ShardedEngine(eng, 10, stats_tensor=eng.stats)
    calls = []
    t = bench.time_em_step(sh, np.maximum(w.d, 0.01), 0.2, 0.15, float((w.codes != -1).sum()), 40, 3, 3, lambda: calls.append(1))
    assert isinstance(t, float) and t > 0 and len(calls) == 2

    class Broken:
        def em_step(self, *a):
            raise RuntimeError("boom")
    assert "boom" in bench.time_em_step(Broken(), None, 0.1, 0.1, 1.0, 1, 2, 2, lambda: None)
    # a failure on ANOTHER rank: this rank's calls work, the collective AND says no -- no barrier is entered
    calls.clear()
    t = bench.time_em_step(sh, np.maximum(w.d, 0.01), 0.2, 0.15, float((w.codes != -1).sum()), 40, 3, 3, lambda: calls.append(1),
                           all_ok=lambda ok: False)
    assert t == "failed on another rank" and calls == []
