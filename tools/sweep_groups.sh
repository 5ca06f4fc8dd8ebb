for g in 2 3 4 5 6 8; do
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-metric2 --groups $g 2>/dev/null > gpurun_out/g$g.json
python - <<PY
import json
d=json.load(open("gpurun_out/g$g.json")); print("groups", $g, "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1))
PY
done
