mkdir -p gpurun_out
timeout 330 python -m pytest tests -q -x -m gpu --durations=8 > gpurun_out/r1h_pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/r1h_rc.txt
timeout 100 python tools/bench_ivf_latency.py flat graph 1,8,32,64 > gpurun_out/r1h_latency_flat.log 2>&1; echo "lat_flat rc=$?" >> gpurun_out/r1h_rc.txt
timeout 100 python tools/bench_ivf_latency.py pq graph 1,8,32,64 > gpurun_out/r1h_latency_pq.log 2>&1; echo "lat_pq rc=$?" >> gpurun_out/r1h_rc.txt
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1h_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r1h_rc.txt
echo done >> gpurun_out/r1h_rc.txt
