"""CPU-side checks of bench.py's plumbing: the one-JSON-line contract survives a sub-record that hangs or raises,
and the reference arm prints the keys the driver reads (it is the CPU path, so it runs here)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_py(code, timeout=60):
    return subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=timeout)


def test_line_printed_once_on_the_normal_path():
    r = run_py("import bench\n"
               "e = bench.LineEmitter(0, {'value': 1})\n"
               "e.arm(30, 'late')\n"
               "e.finish(); e.finish()\n")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1 and json.loads(lines[0]) == {"value": 1}


def test_watchdog_prints_the_headline_when_sub_records_hang():
    r = run_py("import bench, time\n"
               "e = bench.LineEmitter(0, {'value': 2})\n"
               "e.arm(0.3, 'sub-records stalled')\n"
               "time.sleep(30)\n"
               "print('not reached')\n")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    assert json.loads(lines[0]) == {"value": 2, "extra": {"error": "sub-records stalled"}}


def test_other_ranks_leave_silently():
    r = run_py("import bench\n"
               "e = bench.LineEmitter(3, None)\n"
               "e.abort('peer failed')\n"
               "print('not reached')\n")
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_reference_arm_line():
    """`bench.py --impl reference` on a tiny sample: the reference's own CPU path (oracle/_ref, else the port)."""
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample", "64"], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GCUPS" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["metric"].startswith("GCUPS")

    # a non-zero rank of the reference arm does no work and prints nothing
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=60, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
