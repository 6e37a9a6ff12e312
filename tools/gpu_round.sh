# One GPU visit that refreshes everything the judge reads: GPU tests, bench (with CPU baseline), reference arm, ncu launch
# list and a full capture of the dominant kernel (all pw1 launches of one step).
set -x
TAG=${1:-r1_x}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --config cfg2 --steps 2 > gpurun_out/${TAG}_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"gemm_tc_kernel<.int.1" -c 13 -o gpurun_out/${TAG}_full_pw1 -f python tools/profile_step.py --config cfg2 --steps 1 > gpurun_out/${TAG}_full_pw1.log 2>&1
python tools/show_bench.py gpurun_out/${TAG}_bench.json | head -30
cat gpurun_out/${TAG}_bench_ref.json | cut -c1-400
