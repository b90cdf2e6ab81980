# One 8-GPU session (gpurun --gpus 8): multi-GPU parity tests, the headline config at 8/4/2/1 GPUs, config 5 on the
# full 100M rows, the single-process group.  Logs in gpurun_out/${T}_*.
set -x
cd ${GRAFT_REPO_ROOT:-.}
T=${1:-s8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/${T}_multi.log 2>&1; tail -3 gpurun_out/${T}_multi.log
for n in 8 4 2; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29600 + n)) bench.py --gpus $n --steps 200 --warmup 10 > gpurun_out/${T}_c4_g$n.log 2>&1
  tail -1 gpurun_out/${T}_c4_g$n.log | cut -c1-200
done
timeout 300 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_c4_g1.log 2>&1; tail -1 gpurun_out/${T}_c4_g1.log | cut -c1-200
timeout 400 $TR --nproc-per-node 8 --master-port 29620 bench.py --gpus 8 --config 5 --steps 30 --warmup 5 > gpurun_out/${T}_c5_g8.log 2>&1; tail -1 gpurun_out/${T}_c5_g8.log | cut -c1-300
timeout 300 python tools/group_probe.py 10000000 8 > gpurun_out/${T}_group.log 2>&1; cat gpurun_out/${T}_group.log | tail -4
