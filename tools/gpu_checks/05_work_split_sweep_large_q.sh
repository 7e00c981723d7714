mkdir -p gpurun_out
timeout 120 python tools/sweep_work_split.py flat 16,32,64,128,256 0,1,2,4,8 > gpurun_out/r1j_split_flat.log 2>&1; echo "flat rc=$?" >> gpurun_out/r1j_rc.txt
timeout 120 python tools/sweep_work_split.py pq 16,32,64,128,256 0,1,2,4 > gpurun_out/r1j_split_pq.log 2>&1; echo "pq rc=$?" >> gpurun_out/r1j_rc.txt
