# round 2 visit: gpu tests (optionally a -k subset), then a bench (optionally without the other configs)
#   gpurun -- 'bash tools/gpu_r2.sh TAG "pytest -k expr or empty" "bench extra args"'
TAG=${1:-r2_x}
EXPR=${2:-}
BARGS=${3:---others none --no-cpu-baseline}
if [ -n "$EXPR" ]; then
  timeout 1500 python -m pytest tests -m gpu -q -rP -k "$EXPR" > gpurun_out/${TAG}_pytest.log 2>&1
else
  timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/${TAG}_pytest.log 2>&1
fi
grep -n "^E  \|^FAILED\|^ERROR\|worst cases" gpurun_out/${TAG}_pytest.log | head -60
tail -2 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 $BARGS > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -${4:-45}
