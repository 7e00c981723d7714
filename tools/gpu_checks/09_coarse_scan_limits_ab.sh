mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_ivf.py -q -x -m gpu -k "coarse or graph or grouped_scan_ragged or more_than_128" > gpurun_out/r1n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1n_rc.txt
B2VS_COARSE_SCAN_MAXQ=32 B2VS_COARSE_SCAN_CTAS=8 timeout 60 python tools/sweep_work_split.py flat 8,16,32 0 > gpurun_out/r1n_flat_wide.log 2>&1; echo "flat_wide rc=$?" >> gpurun_out/r1n_rc.txt
timeout 60 python tools/sweep_work_split.py flat 8,16,32 0 > gpurun_out/r1n_flat_default.log 2>&1; echo "flat_default rc=$?" >> gpurun_out/r1n_rc.txt
B2VS_COARSE_SCAN_MAXQ=32 B2VS_COARSE_SCAN_CTAS=8 timeout 60 python tools/sweep_work_split.py pq 4,8,16 0 > gpurun_out/r1n_pq_wide.log 2>&1; echo "pq_wide rc=$?" >> gpurun_out/r1n_rc.txt
timeout 60 python tools/sweep_work_split.py pq 4,8,16 0 > gpurun_out/r1n_pq_default.log 2>&1; echo "pq_default rc=$?" >> gpurun_out/r1n_rc.txt
