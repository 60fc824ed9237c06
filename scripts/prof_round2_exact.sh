# exact-order kernel with the feeder warp (profiles/r02c_*): launch list, full capture at config 3, one band alone, DRAM bytes at config 5
set -x
cd $GRAFT_REPO_ROOT
python scripts/prof_run.py c3 exact 2 > gpurun_out/r02c_plain_c3_exact.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02c_launches_c3_exact.csv python scripts/prof_run.py c3 exact 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_project_systolic4 -s 1 -c 1 -f -o gpurun_out/r02c_systolic4_c3 python scripts/prof_run.py c3 exact 2 > /dev/null 2>&1
python scripts/prof_run.py c5 exact 2 > gpurun_out/r02c_plain_c5_exact.log 2>&1 || exit 1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"k_project_systolic4|k_advect32" -s 2 -c 2 --csv --log-file gpurun_out/r02c_c5_exact_dram.csv python scripts/prof_run.py c5 exact 2 > /dev/null 2>&1
ls -la gpurun_out | grep r02c
