"""Run tools/tier_check.py over a matrix of (library build, volume shape) and collect the lines.

    python tools/perf_matrix.py gpurun_out/perf_matrix.json [points]

Builds come from tools/_variants/ (tools/build_variant.py) next to the product library.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out_path = sys.argv[1]
points = sys.argv[2] if len(sys.argv) > 2 else "400000000"
var = os.path.join(ROOT, "tools", "_variants")
libs = {"product": None}
for f in sorted(os.listdir(var)) if os.path.isdir(var) else []:
    if f.startswith("liblrm_") and f.endswith(".so") and "skel" not in f:
        libs[f[len("liblrm_"):-3]] = os.path.join(var, f)
shapes = [("3", "512", "0")]
rows = []
for name, lib in libs.items():
    for cell, dim, kern in (shapes if name == "product" else shapes[:1]):
        env = dict(os.environ)
        if lib:
            env["LRM_B200_LIB"] = lib
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "tier_check.py"), points, "lattice", cell, dim, kern],
                           capture_output=True, text=True, env=env, cwd=ROOT)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
        try:
            row = json.loads(line)
        except Exception:
            row = {"error": (r.stderr or r.stdout)[-800:]}
        row["build"], row["cell"], row["dim"], row["kernel"] = name, cell, dim, kern
        rows.append(row)
        print(json.dumps({k: row.get(k) for k in ("build", "cell", "dim", "kernel", "flags_equal", "max_vec_diff_mm", "auto_equal",
                                                  "reach_equal", "error")} |
                         {m: row.get(m, {}).get("fused_gpoints_s") for m in ("two_tier", "three_tier", "auto")} |
                         {"dist": row.get("three_tier", {}).get("dist_gpoints_s"),
                          "reach": row.get("three_tier", {}).get("reach_gpoints_s"),
                          "first_s": row.get("three_tier", {}).get("first_call_s")}), flush=True)
        with open(out_path, "w") as f:
            json.dump(rows, f, indent=1)
