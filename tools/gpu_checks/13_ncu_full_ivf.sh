#!/bin/bash
# ncu --set full captures of the two grouped-scan kernels on the bench corpora (one search each):
#   C3: bf_tc_kernel<1,true> main pass = 3rd bf_tc launch of the profiled search (coarse, seed, main)
#   C4: pq_tc_kernel main pass         = 2nd pq_tc launch (seed, main)
# run under gpurun; summaries -> profiles/ via tools/ncu_summary.py
set -x
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bf_tc_kernel -s 2 -c 1 \
    -o gpurun_out/r2_c3_scan -f python tools/ivf_phase_probe.py C3 16 > gpurun_out/r2_c3_scan.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:pq_tc_kernel -s 1 -c 1 \
    -o gpurun_out/r2_c4_scan -f python tools/ivf_phase_probe.py C4 16 > gpurun_out/r2_c4_scan.log 2>&1
ls -la gpurun_out/*.ncu-rep
