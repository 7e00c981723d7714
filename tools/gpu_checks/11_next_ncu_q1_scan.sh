# NOT run in round 1 (GPU budget spent): --set full captures of the two kernels that dominate a
# single-query IVF-Flat search (grouped list scan in work-table mode, seed pass), for the next round.
mkdir -p gpurun_out
python tools/q1_probe.py flat > gpurun_out/q1_plain.log 2>&1 || exit 1    # must exit 0 without ncu first
ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'bf_tc_kernel|ivf_seed_tau_kernel' -o gpurun_out/q1_scan_seed -f \
    python tools/q1_probe.py flat > gpurun_out/q1_ncu.log 2>&1
python tools/ncu_summary.py gpurun_out/q1_scan_seed.ncu-rep gpurun_out/r2_q1_scan_seed
