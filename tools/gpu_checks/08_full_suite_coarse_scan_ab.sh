mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -x -m gpu > gpurun_out/r1m_pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/r1m_rc.txt
timeout 60 python tools/sweep_work_split.py flat 1,2,4,8,16 0 > gpurun_out/r1m_lat_flat.log 2>&1; echo "flat rc=$?" >> gpurun_out/r1m_rc.txt
B2VS_COARSE_SCAN=0 timeout 60 python tools/sweep_work_split.py flat 1,4,8 0 > gpurun_out/r1m_lat_flat_tcprobe.log 2>&1; echo "flat_tc rc=$?" >> gpurun_out/r1m_rc.txt
timeout 60 python tools/sweep_work_split.py pq 1,2,4,8 0 > gpurun_out/r1m_lat_pq.log 2>&1; echo "pq rc=$?" >> gpurun_out/r1m_rc.txt
B2VS_COARSE_SCAN=0 timeout 60 python tools/sweep_work_split.py pq 1,4 0 > gpurun_out/r1m_lat_pq_tcprobe.log 2>&1; echo "pq_tc rc=$?" >> gpurun_out/r1m_rc.txt
