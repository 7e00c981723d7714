mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_encoder_handoff.py -q -x -m gpu > gpurun_out/r1g_encode_test.log 2>&1; echo "encode_test rc=$?" >> gpurun_out/r1g_rc.txt
timeout 150 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1g_q1_flat_launches.csv python tools/q1_probe.py flat > gpurun_out/r1g_q1_flat.log 2>&1; echo "q1_flat rc=$?" >> gpurun_out/r1g_rc.txt
timeout 150 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1g_q1_pq_launches.csv python tools/q1_probe.py pq > gpurun_out/r1g_q1_pq.log 2>&1; echo "q1_pq rc=$?" >> gpurun_out/r1g_rc.txt
echo done >> gpurun_out/r1g_rc.txt
