# round 2, first visit: whole GPU suite (all failures listed), then the bench with the other configs
set -x
TAG=${1:-r2_a}
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/${TAG}_pytest.log 2>&1
grep -n "^E  \|^FAILED\|^ERROR\|worst cases" gpurun_out/${TAG}_pytest.log | head -60
tail -2 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -5 gpurun_out/${TAG}_bench.err
python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -20
