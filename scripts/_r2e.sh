mkdir -p gpurun_out/r2e
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e/pytest_gpu.txt 2>&1; echo "rc=$?" >> gpurun_out/r2e/pytest_gpu.txt
timeout 300 python scripts/bench_needles.py > gpurun_out/r2e/needles.jsonl 2> gpurun_out/r2e/needles.err; echo "rc=$?" >> gpurun_out/r2e/needles.err
grep -E "passed|failed|rc=" gpurun_out/r2e/pytest_gpu.txt | tail -3; cat gpurun_out/r2e/needles.jsonl; tail -3 gpurun_out/r2e/needles.err
