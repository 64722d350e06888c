"""bench.py's bookkeeping, checked without a GPU: the algorithmic work figures it reports against the roofline equal
SURVEY.md §8(d)'s, and the reference arm prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_linear_work_per_forward_matches_the_survey():
    import bench
    ops = bench.linear_ops_per_forward()
    # 391.7 GMAC = 783 GOP per block, x28 = 21.93 TOP per forward (109-token prompt here: kv_linear is ~0.1 %)
    assert abs(ops / 1e12 - 21.93) < 0.05, ops
    assert bench.METRIC == "stdit_16x512x512_w8a8_denoise_steps_per_sec" and bench.UNIT == "steps/s"
    # PixArt-alpha 512 solver step (B = 1, CFG batch 2) = 2.17 TOP
    assert abs(bench.pixart_linear_ops() / 1e12 - 2.17) < 0.03, bench.pixart_linear_ops()


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-400:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "stdit_16x512x512_w8a8_denoise_steps_per_sec"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
