mkdir -p gpurun_out
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r1l_bench.json 2> gpurun_out/r1l_bench.err; echo "bench rc=$?" >> gpurun_out/r1l_rc.txt
timeout 100 python tools/bench_ivf_latency.py flat graph 1,8,32,64 > gpurun_out/r1l_latency_flat.log 2>&1; echo "lat_flat rc=$?" >> gpurun_out/r1l_rc.txt
timeout 100 python tools/bench_ivf_latency.py pq graph 1,8,32,64 > gpurun_out/r1l_latency_pq.log 2>&1; echo "lat_pq rc=$?" >> gpurun_out/r1l_rc.txt
