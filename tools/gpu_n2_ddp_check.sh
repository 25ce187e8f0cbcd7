#!/bin/bash
# 2 GPUs: the data-parallel graphed training test and the driver-style bench line with the weak / train keys
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_training_gpu.py -m gpu -q -k "ddp" 2>&1 | tail -3 | tee gpurun_out/r3v_ddp_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 --steps 20 --warmup 5 \
   --extras weak,train --no-cpu-baseline > gpurun_out/r3v_bench_n2.json 2> gpurun_out/r3v_bench_n2.err; echo "bench N=2 exit $?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r3v_bench_n2.json') if l.startswith(chr(123))][0])
print(round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'weak', round(d.get('weak',{}).get('value',0)), 'train', round(d['train']['value']), round(d['train']['ms_per_step'],2))"
