"""Job files for viyadb_b200/host/_build/vgpu_cli (the reference process with GpuQueryRunner plugged in)."""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CLI = os.path.join(ROOT, "viyadb_b200", "host", "_build", "vgpu_cli")
REF = os.path.join(ROOT, "oracle", "_ref")
sys.path.insert(0, os.path.join(HERE, "golden"))


def run_scenario(sc, stock_only=False, timeout=600):
    job = {"ref_root": os.path.join(REF, "root"), "state_dir": os.path.join(REF, "state"),
           "table": sc["table"], "queries": sc["queries"], "stock_only": stock_only}
    for k in ("rows", "generate", "rollup_ts", "reload_rows", "reload_mode"):
        if k in sc:
            job[k] = sc[k]
    fd, path = tempfile.mkstemp(suffix=".json", prefix="vgpu_job_")
    with os.fdopen(fd, "w") as f:
        json.dump(job, f)
    p = subprocess.run([CLI, path], capture_output=True, text=True, timeout=timeout)
    os.unlink(path)
    lines = p.stdout.strip().splitlines()
    if not lines:
        raise RuntimeError(f"vgpu_cli produced no output (rc={p.returncode}): {p.stderr[-800:]}")
    return json.loads(lines[-1])


if __name__ == "__main__":
    # pre-warm the reference's JIT cache (oracle/_ref/state) for every scenario: run in the authoring
    # container, where /root/reference and g++ can compile; the .so files then travel with gpurun
    import scenarios
    for sc in scenarios.SCENARIOS + scenarios.SELECT_SCENARIOS:
        out = run_scenario(sc, stock_only=True)
        bad = [r for r in out.get("results", []) if "error" in r.get("stock", {})]
        print(sc["name"], "fatal: " + out["fatal"] if "fatal" in out else f"{len(out['results'])} queries, {len(bad)} errors")
