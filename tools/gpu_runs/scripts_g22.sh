cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 400 compute-sanitizer --tool memcheck python tools/racecheck_small_m8.py > gpurun_out/memcheck_m8.log 2>&1; tail -4 gpurun_out/memcheck_m8.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/racecheck_small_m8.py > gpurun_out/racecheck_m8.log 2>&1; grep -c "Race reported" gpurun_out/racecheck_m8.log; tail -3 gpurun_out/racecheck_m8.log
grep "Race reported\|and Read\|and Write" gpurun_out/racecheck_m8.log | sed 's/void pnb:://; s/(CUtensorMap_st.*)+0x[0-9a-f]* / /' | cut -c1-200 | sort | uniq -c | sort -rn | head -20
