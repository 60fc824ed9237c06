# red-black kernel at KB = 3 (profiles/r02b_*): launch list, full captures at configs 3 and 4, DRAM bytes at config 5
set -x
cd $GRAFT_REPO_ROOT
python scripts/prof_run.py c3 rb 2 > gpurun_out/r02b_plain_c3_rb.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02b_launches_c3_rb.csv python scripts/prof_run.py c3 rb 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_project_rb -s 40 -c 1 -f -o gpurun_out/r02b_rb_c3 python scripts/prof_run.py c3 rb 2 > /dev/null 2>&1
python scripts/prof_run.py c4 rb 1 > gpurun_out/r02b_plain_c4_rb.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_project_rb -s 8 -c 1 -f -o gpurun_out/r02b_rb_c4 python scripts/prof_run.py c4 rb 1 > /dev/null 2>&1
python scripts/prof_run.py c5 rb 1 > gpurun_out/r02b_plain_c5_rb.log 2>&1 || exit 1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"k_project_rb|k_advect32" -s 20 -c 32 --csv --log-file gpurun_out/r02b_c5_rb_dram.csv python scripts/prof_run.py c5 rb 1 > /dev/null 2>&1
ls -la gpurun_out | tail -12
