mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bvh_cache" 2>&1 | tail -5 | tee gpurun_out/pytest_r5f.log
timeout 600 python tools/bvh_bench.py 8 2>&1 | tee gpurun_out/bvh_bench_r5f.txt
timeout 600 python tools/bvh_bench.py 6 2>&1 | tee -a gpurun_out/bvh_bench_r5f.txt
