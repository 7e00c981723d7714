#!/bin/bash
# ncu --set full captures of the grouped-scan kernels on the bench corpora (one search each, graphs
# and the planning overlap off so the launch order is the program order):
#   C3: bf_tc_kernel launches of one search = two-pass coarse probe (pass 1, pass 2), seed pass, MAIN
#       pass -> the 4th launch
#   C4: pq_tc_kernel launches = seed pass, MAIN pass -> the 2nd launch; and the two coarse-probe
#       passes (bf_tc_kernel launches 1-2)
# run under gpurun; summaries -> profiles/ via tools/ncu_summary.py
set -x
export B2VS_GRAPH=0 B2VS_PLAN_OVERLAP=0
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bf_tc_kernel -s 3 -c 1 \
    -o gpurun_out/r2_c3_scan -f python tools/ivf_phase_probe.py C3 16 > gpurun_out/r2_c3_scan.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:pq_tc_kernel -s 1 -c 1 \
    -o gpurun_out/r2_c4_scan -f python tools/ivf_phase_probe.py C4 16 > gpurun_out/r2_c4_scan.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bf_tc_kernel -c 2 \
    -o gpurun_out/r2_c4_coarse -f python tools/ivf_phase_probe.py C4 16 > gpurun_out/r2_c4_coarse.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2_c3_launches.csv python tools/ivf_phase_probe.py C3 16 > gpurun_out/r2_c3_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2_c4_launches.csv python tools/ivf_phase_probe.py C4 16 > gpurun_out/r2_c4_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep
