# usage: bash tools/ncu_one.sh <regex> <skip> <count> <outname> [extra profile_step args]
N="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
P="python tools/profile_step.py --config cfg2 --steps 1 $5"
timeout 600 $N -k regex:"$1" -s $2 -c $3 -o gpurun_out/$4 -f $P > gpurun_out/$4.log 2>&1
tail -3 gpurun_out/$4.log
