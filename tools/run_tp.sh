set -x
N=${1:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/tp_check.py 2>&1 | grep -E "tp_check|fused|Error|error" | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N > gpurun_out/bench_r1_tp$N.json 2> gpurun_out/bench_r1_tp$N.err; tail -c 900 gpurun_out/bench_r1_tp$N.json; tail -3 gpurun_out/bench_r1_tp$N.err
