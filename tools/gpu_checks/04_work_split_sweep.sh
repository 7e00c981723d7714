mkdir -p gpurun_out
timeout 120 python tools/sweep_work_split.py flat > gpurun_out/r1i_split_flat.log 2>&1; echo "flat rc=$?" >> gpurun_out/r1i_rc.txt
timeout 120 python tools/sweep_work_split.py pq > gpurun_out/r1i_split_pq.log 2>&1; echo "pq rc=$?" >> gpurun_out/r1i_rc.txt
