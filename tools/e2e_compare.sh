# Developer tool: host-entry wall clock (tools/time_e2e_breakdown.py) next to the e2e leg of the default bench, same box.
python tools/time_e2e_breakdown.py | tail -2
python bench.py --steps 100 --warmup 5 --no-sub > gpurun_out/e2e_cmp.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/e2e_cmp.json").read().strip().splitlines()[-1])
print("bench e2e solves/s %.0f = %.4f ms; device %.4f ms" % (d["e2e"]["value"], 4096 / d["e2e"]["value"] * 1e3, d["ms_per_step"]))
PY
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv,noheader
