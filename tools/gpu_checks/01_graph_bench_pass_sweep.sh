mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r1f_gpu.txt 2>&1
timeout 170 python -m pytest tests/test_gpu_ivf.py -q -x -k cuda_graph > gpurun_out/r1f_graph_test.log 2>&1; echo "graph_test rc=$?" >> gpurun_out/r1f_rc.txt
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err; echo "bench rc=$?" >> gpurun_out/r1f_rc.txt
timeout 150 python tools/bench_ivf_latency.py flat graph 1,8,32,64,128 > gpurun_out/r1f_latency_flat.log 2>&1; echo "lat_flat rc=$?" >> gpurun_out/r1f_rc.txt
timeout 120 python tools/bench_ivf_latency.py pq graph 1,8,32,64,128 > gpurun_out/r1f_latency_pq.log 2>&1; echo "lat_pq rc=$?" >> gpurun_out/r1f_rc.txt
B2VS_GRAPH=1 timeout 240 python -m pytest tests/test_gpu_ivf.py tests/test_gpu_managers.py -q -x > gpurun_out/r1f_graph_on_suite.log 2>&1; echo "graph_on_suite rc=$?" >> gpurun_out/r1f_rc.txt
for p in 16,1 32,1 64,1 64,4,1 8,1; do B2VS_PASSES=$p timeout 60 python bench.py --n-db 1250000 --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; l=json.loads(sys.stdin.readline()); print('$p', l['value'], l['ms_per_step'], l['roofline']['frac'])" >> gpurun_out/r1f_pass_sweep.txt 2>&1; done
echo done >> gpurun_out/r1f_rc.txt
