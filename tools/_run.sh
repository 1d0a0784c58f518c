mkdir -p gpurun_out
./tools/micro/tma_stage_test 2>&1 | tee gpurun_out/tma_stage_test.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "deterministic or staged or merged or radiance or closest" 2>&1 | tail -15 | tee gpurun_out/pytest_r5b.log
timeout 600 python tools/bin_exp.py flat_block 256,384,768 diamond_scene.json evaluation/cbox-d6.json many_point_lights.json 2>&1 | tee gpurun_out/exp_r5b.txt
