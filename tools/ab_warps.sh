for cfg in "8 8" "4 4" "2 2" "4 8" "8 4"; do
  set -- $cfg
  FB200_FWD_WARPS=$1 FB200_BWD_WARPS=$2 timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/ab_$1_$2.json 2> gpurun_out/ab.err
  python - <<PY
import json
o=json.loads([l for l in open("gpurun_out/ab_$1_$2.json") if l.startswith("{")][-1])
print("fwd_warps=$1 bwd_warps=$2", round(o["value"],1), {k: round(v,4) for k,v in o["stage_ms"].items()})
PY
done
