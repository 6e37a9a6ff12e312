# One GPU visit that refreshes everything the judge reads (round 2): GPU tests, bench (with the other configs and the CPU
# baseline), reference arm, ncu launch list, --set full captures of the two dominant kernels (pw1, sparse pw2) and a
# per-kernel capture of one whole step.  The .ncu-rep files are summarised ON THE BOX and deleted (gpurun returns <= 64 MiB).
set -x
TAG=${1:-r2_x}
TITLE=${2:-"Round 2"}
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/${TAG}_pytest.log 2>&1; tail -2 gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_step.py --config cfg2 --steps 2 > gpurun_out/${TAG}_prof.log 2>&1
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"gemm_tc_kernel<.int.1" -c 13 -o gpurun_out/${TAG}_full_pw1 -f python tools/profile_step.py --config cfg2 --steps 1 > gpurun_out/${TAG}_full_pw1.log 2>&1
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:"gemm_tc_kernel<.int.0, .bool.1, .bool.1, .bool.1, .bool.1>" -c 12 -o gpurun_out/${TAG}_full_pw2 -f python tools/profile_step.py --config cfg2 --steps 1 > gpurun_out/${TAG}_full_pw2.log 2>&1
python tools/summarize_ncu.py gpurun_out/${TAG}_launches.csv --steps 2 --rep gpurun_out/${TAG}_full_pw1.ncu-rep --title "$TITLE" > gpurun_out/${TAG}_launches.md
python tools/summarize_ncu.py gpurun_out/${TAG}_launches.csv --steps 2 --rep gpurun_out/${TAG}_full_pw2.ncu-rep --title "$TITLE (sparse pw2 capture)" | sed -n '/ncu --set full/,$p' >> gpurun_out/${TAG}_launches.md
python tools/traffic_from_rep.py gpurun_out/${TAG}_full_pw1.ncu-rep pw1 "profiles/${TAG}_launches.md" > gpurun_out/${TAG}_traffic_pw1.json
python tools/traffic_from_rep.py gpurun_out/${TAG}_full_pw2.ncu-rep pw2 "profiles/${TAG}_launches.md" > gpurun_out/${TAG}_traffic_pw2.json
python tools/launch_table.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launch_table.txt
rm -f gpurun_out/${TAG}_full_pw1.ncu-rep gpurun_out/${TAG}_full_pw2.ncu-rep
timeout 900 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section ComputeWorkloadAnalysis --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -o gpurun_out/${TAG}_step -f python tools/profile_step.py --config cfg2 --steps 1 > gpurun_out/${TAG}_step.log 2>&1
python tools/ncu_kernel_table.py gpurun_out/${TAG}_step.ncu-rep --title "$TITLE: every kernel of one step" > gpurun_out/${TAG}_kernels.md
rm -f gpurun_out/${TAG}_step.ncu-rep
python tools/show_bench.py gpurun_out/${TAG}_bench.json 2>/dev/null | head -12
du -sh gpurun_out
