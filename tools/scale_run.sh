#!/bin/bash
# usage: tools/scale_run.sh N  -> bench.py on N GPUs of one box (cfg4 weak, cfg5 strong), logs under gpurun_out/
N=$1
export HGP_BENCH_WATCHDOG=240
for cfg in cfg4 cfg5; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 5 --warmup 3 --config $cfg --no-cpu > gpurun_out/scale_${cfg}_n${N}.log 2>&1
  echo "rc=$?" >> gpurun_out/scale_${cfg}_n${N}.log
  grep '"metric"' gpurun_out/scale_${cfg}_n${N}.log | cut -c1-420
done
