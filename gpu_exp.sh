mkdir -p gpurun_out
ncu --set full --cache-control none --clock-control none --import-source on -k regex:step_kernel -s 200 -c 3 -o gpurun_out/prof_cp64k_hot python bench.py --profile --steps 300 --warmup 50 --no-graph > gpurun_out/ncu_hot.log 2>&1
ncu -i gpurun_out/prof_cp64k_hot.ncu-rep --page raw --csv | python profiles/ncu_pick.py 'gpu__time_duration.sum|dram__bytes_read.sum$|dram__bytes_write.sum$|sm__warps_active.avg.pct_of_peak_sustained_active|launch__registers|smsp__inst_executed.sum$|smsp__issue_active.avg.pct|thread_inst_executed_per_inst|lts__t_sector_hit_rate.pct|stalled_.*per_issue_active|sm__cycles_elapsed.max|launch__waves|launch__grid_size'
ncu -i gpurun_out/prof_cp64k_hot.ncu-rep --page source --csv > gpurun_out/prof_cp64k_hot_source.csv
