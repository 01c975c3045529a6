#!/bin/bash
# Round-2 ncu captures (one GPU, under gpurun): launch list of a short 50 M-DOF solve and --set full captures of the
# kernels that carry an iteration.  Results land in gpurun_out/; the summaries copied to profiles/ are what is judged.
mkdir -p gpurun_out
export BFM_QUIET=1
CELLS=${1:-10000x2500}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_${CELLS}.csv python tools/profile_target.py $CELLS 6 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:"k_spmv_mg|k_spmv<|k_update_xr|k_update_p" -s 4 -c 10 -o gpurun_out/r2_prof_level0 python tools/profile_target.py $CELLS 4 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:"k_mg_restrict|k_mg_prolong" -s 2 -c 8 -o gpurun_out/r2_prof_transfer python tools/profile_target.py $CELLS 3 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:"k_blk_spmv" -s 2 -c 6 -o gpurun_out/r2_prof_coarse python tools/profile_target.py $CELLS 3 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:"k_assemble|k_mg_rap|k_residual_dd" -c 5 -o gpurun_out/r2_prof_setup python tools/profile_target.py $CELLS 3 2>&1 | tail -2
ls -la gpurun_out/*.ncu-rep
