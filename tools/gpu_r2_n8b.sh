# 8-GPU visit (final state of round 2): the bench line the driver's scaling run takes, with the other configs
TAG=${1:-r2_n8b}
N=${2:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 --others cfg3,cfg4 > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err
tail -3 gpurun_out/${TAG}.err
python tools/show_bench.py gpurun_out/${TAG}.json | grep -E "^value|cfg|replica|ddp|step_api|eager|e2e_raw" | cut -c1-300
