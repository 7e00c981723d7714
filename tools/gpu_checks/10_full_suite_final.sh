mkdir -p gpurun_out
timeout 100 python -m pytest tests -q -x -m gpu > gpurun_out/r1o_pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" >> gpurun_out/r1o_rc.txt
