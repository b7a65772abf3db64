set -x
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py 2> gpurun_out/bench_err.log | tee gpurun_out/bench_n1.json | cut -c1-600
tail -3 gpurun_out/bench_err.log
