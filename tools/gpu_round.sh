mkdir -p gpurun_out
export BFM_QUIET=1
echo "== new tests"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "galerkin_product or multilevel_preconditioner_changes or deterministic" 2>&1 | tail -4
echo "== ncu launch list (defaults: smoothed aggregation)"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_sa_default_10000x2500.csv python tools/profile_target.py 10000x2500 8 2>&1 | tail -1
echo "== ncu --set full: set-up kernels of smoothed aggregation"
ncu --set full --clock-control none --import-source on -k regex:"k_mg_ap|k_mg_ptq|k_mg_smooth" -c 6 -o gpurun_out/r2_prof_sa_setup python tools/profile_target.py 10000x2500 2 2>&1 | tail -2
ncu -i gpurun_out/r2_prof_sa_setup.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY'
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size"]
idx = [hdr.index(w) for w in want if w in hdr]
out = open("gpurun_out/r2_ncu_sa_setup_extract.csv", "w")
for r in rows:
    line = ",".join('"' + r[i] + '"' for i in idx)
    print(line[:300]); out.write(line + "\n")
PY
ls -la gpurun_out/r2_prof_sa_setup.ncu-rep
