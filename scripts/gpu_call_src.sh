cd "$GRAFT_REPO_ROOT"
export DIAG_EVERY=1 DIAG_PLANS="4x2@0.3,0x0"
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --import-source on --clock-control none -k regex:enet_cd_cluster -s 2 -c 1 -o gpurun_out/r2_cdsrc python scripts/cd_timers.py > gpurun_out/r2_cdsrc.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2_cdsrc.ncu-rep --page source --csv > gpurun_out/r2_cdsrc_source.csv 2>/dev/null; echo "export rc=$?"
ncu -i gpurun_out/r2_cdsrc.ncu-rep --page raw --csv > gpurun_out/r2_cdsrc_raw.csv 2>/dev/null
ls -la gpurun_out/; rm -f gpurun_out/r2_cdsrc.ncu-rep
