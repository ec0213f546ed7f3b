"""The reference arm of bench.py (`--impl reference`) needs no GPU: it times the CPU port of the path and must print one
JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3",
                          "--records", "20000"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "records/s" and d["higher_is_better"] is True and d["steps"] == 2
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["vs_baseline"] is None and d["dtype"] == "u8"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and 1 <= cb["cores"] <= (os.cpu_count() or 1) and cb["value"] == d["value"] and "sample" in cb
    # both of the reference's chunk settings are timed (its own NUM_CHUNKS = 8, and 4 chunks per usable core)
    assert [s["num_chunks"] for s in cb["settings"]] == [8, 4 * cb["cpus"]["usable"]]
    assert cb["cpus"]["usable"] <= cb["cpus"]["affinity"]
    # the same configuration keys as the GPU arm (the driver compares the two lines' `config`)
    assert d["config"]["records_per_gpu"] == 20000 and d["config"]["num_chunks"] == 8 and "workload" in d["config"]


def test_usable_cores_respects_affinity_and_quota():
    sys.path.insert(0, ROOT)
    import bench
    n, aff, quota = bench.usable_cores()
    assert 1 <= n <= aff
    if quota is not None:
        assert n <= max(1, int(quota + 0.999))


def test_roofline_traffic_comes_from_a_committed_capture():
    """`roofline.traffic` of the bench line is the DRAM byte count of a committed ncu capture of the fused kernel for the
    bench's own workload and size (the first, current, entry of profiles/traffic.json) and names that capture."""
    sys.path.insert(0, ROOT)
    import bench
    traffic, capture = bench.traffic_of("kafka", 10_000_000)
    assert isinstance(traffic, int) and traffic > 0
    algorithmic = 1_170_821_397 + 80_000_008 + 1_670_247_592          # input + offsets + Arrow buffers
    assert 0.9 * algorithmic < traffic < 1.2 * algorithmic             # the fused pass reads the input once
    assert capture and os.path.exists(os.path.join(ROOT, capture.split(" ")[0]))
    assert bench.traffic_of("kafka", 123) == (None, None)              # no capture for other sizes: reported as null
