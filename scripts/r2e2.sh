#!/bin/bash
# round-2 session E2 (2 GPUs): the full default bench line under torchrun exactly as the driver launches it, multi-GPU tests
OUT=gpurun_out/r2e2; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_threads.py tests/test_sharding.py -m gpu -x -q > $OUT/pytest_threads.txt 2>&1; echo "pytest rc=$?" >> $OUT/pytest_threads.txt; tail -3 $OUT/pytest_threads.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; echo "bench rc=$?" >> $OUT/bench_n2.err; tail -3 $OUT/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 > $OUT/bench_reference_n2.json 2>> $OUT/bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2e2/bench_n2.json') if l.startswith('{')][-1])
print('N=2 value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['config']['rank0_device'])
print('recon',d['reconstruct']['roofline_frac'],d['reconstruct']['single_loss'],d['reconstruct']['kernel_source'])
print('sustained',d['sustained']['roofline_frac'],'variant',d['variant_30000MiB']['roofline_frac'],d['variant_30000MiB']['check'][:60])
print('batch',{k:d['batch'][k] for k in ('value','roofline_frac','digest','volumes_per_gpu','check')})
r=[l for l in open('gpurun_out/r2e2/bench_reference_n2.json') if l.startswith('{')]
print('reference lines:',len(r), json.loads(r[-1])['value'] if r else None)
PY
