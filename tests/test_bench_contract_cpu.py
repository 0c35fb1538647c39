"""bench.py's reference arm runs on host cores only, so its JSON line — the contract the driver parses — can be checked here: one
line on stdout, the keys of the base contract plus this tier's `cpu_baseline` and `e2e`, the metric and workload of BASELINE.json."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=tmp_path)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                       # stdout carries exactly one JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decode_tokens_per_sec_batch1_greedy" and d["unit"] == "tokens/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d, key
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic" and d["n_gpus"] == 1
    assert "llama2-7b" in d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["unit"] == "tokens/s" and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == "tokens/s" and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert 0.01 < d["value"] < 10 and abs(d["ms_per_step"] * d["value"] - 1000.0) < 1e-6 * 1000     # a 7B-shaped token takes seconds on one core


def test_ranks_other_than_zero_do_no_work(tmp_path):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=120, cwd=tmp_path, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def _bench_module():
    sys.path.insert(0, ROOT)
    import bench
    return bench


def test_token_check_counts_from_the_first_generated_token():
    """golden_token_check: `generated` starts with the token that follows the last prompt position; the golden stream is indexed by position."""
    import argparse
    import numpy as np
    bench = _bench_module()
    gold = np.load(os.path.join(ROOT, "tests", "golden", "bench_cfg4_stream.npz"))["tokens"]
    P = bench.PROMPT_LEN
    args = argparse.Namespace(config="llama2-7b", wdtype="bf16", kvdtype="bf16", prompt_len=P)
    W, K = 5, 20
    same = gold[P - 1:P - 1 + W + K + 1].copy()
    r = bench.golden_token_check(args, same, P + W)
    assert r["generated_compared"] == W + K + 1 and r["generated_identical_prefix"] == W + K + 1 and r["timed_steps_identical"] == K == r["timed_steps"]
    assert r["oracle_first_generated_token"] == int(gold[P - 1]) and r["first_timed_position"] == P + W
    off = same.copy()
    off[3] += 1                                   # a warm-up token differs: prefix 3, the timed steps (from index W + 1) still all match
    r = bench.golden_token_check(args, off, P + W)
    assert r["generated_identical_prefix"] == 3 and r["timed_steps_identical"] == K
    assert bench.golden_token_check(argparse.Namespace(config="tinyllama-1.1b", wdtype="bf16", kvdtype="bf16", prompt_len=P), same, P + W) is None


def test_clock_sampler_reads_nvidia_smi_lines(tmp_path):
    """ClockSampler.stop: median / min clock, throttle reasons, and how many samples fell inside the marked (timed) region."""
    import datetime
    bench = _bench_module()

    class Done:
        def terminate(self): pass
        def wait(self, timeout=None): return 0
        def kill(self): pass

    s = bench.ClockSampler(0)
    s.path = str(tmp_path / "clocks.csv")
    t0 = datetime.datetime(2026, 1, 2, 3, 4, 5)
    rows = [(t0 + datetime.timedelta(milliseconds=50 * i), clk, cap) for i, (clk, cap) in
            enumerate([(1965, "Not Active"), (1965, "Not Active"), (1800, "Active"), (1965, "Not Active")])]
    with open(s.path, "w") as f:
        for ts, clk, cap in rows:
            f.write(f"{ts.strftime('%Y/%m/%d %H:%M:%S.%f')[:-3]}, {clk}, 1965, 950.5, Not Active, Not Active, Not Active, {cap}\n")
        f.write("garbage line\n")
    s.proc, s.f = Done(), open(os.devnull, "w")
    s.t0, s.t1 = rows[1][0], rows[2][0]
    out = s.stop()
    assert out["samples"] == 4 and out["samples_in_timed_region"] == 2 and out["sm_mhz"] == 1965.0 and out["sm_mhz_min"] == 1800.0
    assert out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"] and abs(out["power_w_max"] - 950.5) < 1e-6
    assert not os.path.exists(s.path)
