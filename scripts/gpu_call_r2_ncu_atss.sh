# ncu evidence for the ATSS kernels after the end-of-round changes (5x5 candidate window + rank by counting in the
# threshold kernel, per-CTA GT cull in the assign kernel, all-levels entry point); the zoo section has exited 0 without
# ncu in this same call before each ncu pass.
mkdir -p gpurun_out
python scripts/kernel_zoo.py atss > gpurun_out/zoo_atss_late.jsonl 2>gpurun_out/zoo_atss_late.err; echo "zoo rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_atss_late.csv python scripts/kernel_zoo.py --once atss > /dev/null 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:mydet -c 3 -o gpurun_out/late_atss -f python scripts/kernel_zoo.py --once atss > gpurun_out/ncu_late_atss.log 2>&1
ncu -i gpurun_out/late_atss.ncu-rep --page raw --csv > gpurun_out/late_atss.csv 2>/dev/null
python scripts/hot_lines.py gpurun_out/late_atss.ncu-rep > gpurun_out/late_hot_atss.txt 2>/dev/null
rm -f gpurun_out/late_atss.ncu-rep
