"""Pipeline probe: time a kernel with parts of its pipeline disabled (MOCO_DEBUG_MODE).  Results are wrong by
construction; only time matters.  Today only the statistics kernel still honours bit 1 (no epilogue math); the
one-pass/dq kernel's modes (1 = no exps, 2 = no MMA issue, 4 = no TMA) were removed after they had served
(profiles/r1_onepass_pipeline_probe.jsonl) -- a per-element `if (debug)` was itself a 2x slowdown there.
tools/trace_probe.py (clock64 timeline) is the tool for that kernel now."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cases = sys.argv[1:] or ["tc1_c5", "dq2_c5"]
for mode in [0, 1, 2, 4, 3, 5, 6, 7]:
    for case in cases:
        env = dict(os.environ, MOCO_DEBUG_MODE=str(mode))
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_lab.py"), case], capture_output=True, text=True, env=env, timeout=200)
        line = [l for l in p.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(json.dumps({"mode": mode, "case": case, "error": p.stderr[-300:]})); continue
        d = json.loads(line[-1])
        print(json.dumps({"mode": mode, "case": case, "stats_us": round(d.get("stats_kernel_us", -1), 1), "dq_us": round(d.get("dq_kernel_us", -1), 1)}))
        sys.stdout.flush()
