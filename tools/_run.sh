mkdir -p gpurun_out
(timeout 300 python tools/bin_exp.py turn_trace_blocks 2,3 synthetic_room.json
timeout 300 python tools/bin_exp.py refill 12,16,20,28,32 synthetic_room.json
timeout 300 python tools/bin_exp.py split_turns 2,3,6,8 synthetic_room.json
timeout 300 python tools/bin_exp.py wide_rays_per_group 0,1,4 synthetic_room.json) 2>&1 | tee gpurun_out/exp_r5j.txt
