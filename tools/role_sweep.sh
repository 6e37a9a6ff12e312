# Builds variants of the library with other warp-role splits (-DTC_W_<set>_E=.. -DTC_W_<set>_S=..) and times every GEMM shape of a cfg2 step
# with each (tools/dbg_sweep2.py, CUDA-graph timing).  Build here, run on the GPU box:
#   bash tools/role_sweep.sh build "A:-DTC_W_BW_E=16 -DTC_W_BW_S=2" "B:-DTC_W_BW_E=16 -DTC_W_BW_S=4" ...
#   gpurun -- 'bash tools/role_sweep.sh run A B ...'
mode=$1; shift
cd "$(dirname "$0")/.."
if [ "$mode" = build ]; then
  for v in "$@"; do
    tag=${v%%:*}; flags=${v#*:}
    (cd mmearth_train_b200 && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -shared -Xcompiler -fPIC $flags \
        -o lib/libmpmae_var_$tag.so csrc/mpmae.cu > /tmp/var_$tag.log 2>&1; echo "$tag rc=$?") &
  done
  wait
else
  for tag in "$@"; do
    echo "== variant $tag"
    if [ "$tag" = base ]; then lib=mmearth_train_b200/lib/libmpmae.so; else lib=mmearth_train_b200/lib/libmpmae_var_$tag.so; fi
    SWEEP_GRAPH=1 MPMAE_LIB=$PWD/$lib timeout 120 python tools/dbg_sweep2.py 2>&1 | grep -v "^dbg\|^knob\|dW"
  done
fi
