# 8-GPU visit: headline line with the other configs, then the NCCL CTA-count sweep
TAG=${1:-r2_n8}
N=${2:-8}
run() {  # name, env..., -- bench args
  name=$1; shift
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus $N "$@" > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_${name}.json").read().strip().splitlines()[-1])
    print("${name}", round(d["value"]), "samples/s", round(d["ms_per_step"], 3), "ms  e2e", round(d["e2e"]["value"]), "replica", d.get("replica_check"), d.get("replica"), "ddp", d.get("ddp_leg"))
    for k, v in (d.get("other_configs") or {}).items():
        print("   ", k, round(v["samples_s"]), "samples/s", round(v["ms_per_step"], 3), "ms hbm_frac", round(v["step_hbm_frac"], 3), "e2e", round(v["e2e"]["value"]))
except Exception as e:
    print("${name} failed", e)
PY
}
run default NCCL_DEBUG=WARN -- --steps 20 --warmup 5 --others cfg3,cfg4
run maxctas4 NCCL_MAX_CTAS=4 -- --steps 20 --warmup 5 --others none
run maxctas2 NCCL_MAX_CTAS=2 -- --steps 20 --warmup 5 --others none
run maxctas8 NCCL_MAX_CTAS=8 -- --steps 20 --warmup 5 --others none
