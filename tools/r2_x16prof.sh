CMD="python bench.py --size 512 --steps 1 --warmup 3 --no-cpu --no-e2e"
ncu --set full --clock-control none --import-source on -k regex:xpass16 -s 3 -c 1 -o /tmp/prof_x16 $CMD > gpurun_out/ncu_x16.log 2>&1
ncu -i /tmp/prof_x16.ncu-rep --page raw --csv > gpurun_out/x16_raw.csv 2>/dev/null
ncu -i /tmp/prof_x16.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/x16_source.csv.gz
python tools/ncu_summary.py gpurun_out/x16_raw.csv
python tools/sass_mix.py gpurun_out/x16_source.csv.gz 25
python tools/ncu_regions.py gpurun_out/x16_source.csv.gz 131584 6
