mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -x -m gpu > gpurun_out/r1k_pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/r1k_rc.txt
timeout 100 python tools/sweep_work_split.py flat 1,2,4,8,16,32,64,128,256,1024 0 > gpurun_out/r1k_lat_flat.log 2>&1; echo "flat rc=$?" >> gpurun_out/r1k_rc.txt
timeout 100 python tools/sweep_work_split.py pq 1,2,4,8,16,32,64,128,256,1024 0 > gpurun_out/r1k_lat_pq.log 2>&1; echo "pq rc=$?" >> gpurun_out/r1k_rc.txt
