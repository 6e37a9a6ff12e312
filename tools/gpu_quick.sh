# One short GPU visit: a subset of the GPU tests (full failure text kept), then a short bench without the CPU-baseline leg.
#   gpurun --timeout 200 -- 'bash tools/gpu_quick.sh TAG "pytest -k expression" [extra env assignments...]'
# Box time is what is charged, and it has been charged at up to 5x the command's run time (r1: 175 s for a 35 s run), so
# keep each visit to what the next decision needs: the -k subset first, the whole suite + tools/gpu_round.sh once per milestone.
TAG=${1:-quick}
EXPR=${2:-}
shift 2 2>/dev/null
for kv in "$@"; do export "$kv"; done
if [ -n "$EXPR" ]; then
  timeout 300 python -m pytest tests -m gpu -q -x -k "$EXPR" > gpurun_out/${TAG}_pytest.log 2>&1
else
  timeout 300 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
fi
grep -n "^E  \|^FAILED\|^ERROR" gpurun_out/${TAG}_pytest.log | head -30
tail -1 gpurun_out/${TAG}_pytest.log
timeout 90 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -14
