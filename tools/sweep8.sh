#!/bin/bash
# Global-batch sweep of the SRGAN step on N GPUs (BASELINE configs 4 and 5): tools/sweep8.sh N OUT_PREFIX
N=${1:-8}; P=${2:-gpurun_out/sweep}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
[ "$N" = "1" ] && TR="python"
run() {  # name, extra args...
  name=$1; shift
  $TR bench.py --gpus $N --steps 8 --warmup 3 --engine bf16 --skip-checks --no-cpu "$@" > ${P}_${name}.json 2> ${P}_${name}.err
  echo "$name rc=$? $(python -c "
import json,sys
try:
    d=json.loads(open('${P}_${name}.json').read().strip().splitlines()[-1]); print('img/s %.1f ms %.2f global_batch %d' % (d['value'], d['ms_per_step'], d['config']['global_batch']))
except Exception as e: print('no line', e)")"
}
