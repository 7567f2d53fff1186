O=gpurun_out/final1
mkdir -p $O
timeout 300 python tools/bench_posit.py --config c5 --check 500 > $O/posit_c5.json 2> $O/posit_c5.err; echo "c5 rc=$?"
timeout 400 python tools/bench_posit.py --config c4 --check 100 > $O/posit_c4.json 2> $O/posit_c4.err; echo "c4 rc=$?"
timeout 400 python bench.py --steps 20 --warmup 3 > $O/bench_weak.json 2> $O/bench_weak.err; echo "bench rc=$?"
python - <<PY
import json
for f in ("posit_c5", "posit_c4"):
    for l in open("$O/%s.json" % f).read().strip().splitlines():
        d = json.loads(l); print(f, {k: d[k] for k in list(d)[:2]}, d.get("kernel_ms_max"), d.get("wall_ms"), d.get("check", {}).get("parity"))
d = json.loads(open("$O/bench_weak.json").read().strip().splitlines()[-1])
print("bench", d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"]); p = d["positionability"]; print({k: p[k] for k in p if k not in ("config",)})
PY
