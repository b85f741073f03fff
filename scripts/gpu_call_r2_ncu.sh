# ncu evidence for every kernel family outside the two of the bench step + the final bench kernels.
# Each ncu pass wraps a command that has already exited 0 without ncu in this same call.  The .ncu-rep files are
# exported to CSV pages on the box and deleted (gpurun_out/ is capped at 64 MiB).
mkdir -p gpurun_out
python scripts/kernel_zoo.py > gpurun_out/zoo.jsonl 2>gpurun_out/zoo.err; echo "zoo rc=$?"
for sec in rot rotbench rot1 rotclu atss fcos rowmax iou iourot dense pre decode_yolo decode_rapid; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$sec.csv python scripts/kernel_zoo.py --once $sec > gpurun_out/ncu_l_$sec.log 2>&1
done
full() {   # name, launch cap, command...
  local name=$1 cap=$2; shift 2
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:mydet -c $cap -o gpurun_out/full_$name -f "$@" > gpurun_out/ncu_f_$name.log 2>&1
  ncu -i gpurun_out/full_$name.ncu-rep --page raw --csv > gpurun_out/full_$name.csv 2>/dev/null
  python scripts/hot_lines.py gpurun_out/full_$name.ncu-rep > gpurun_out/hot_$name.txt 2>/dev/null
  rm -f gpurun_out/full_$name.ncu-rep
}
full rot 24 python scripts/kernel_zoo.py --once rot
full rotbench 24 python scripts/kernel_zoo.py --once rotbench
full rotclu 24 python scripts/kernel_zoo.py --once rotclu
full atss 16 python scripts/kernel_zoo.py --once atss
full fcos 4 python scripts/kernel_zoo.py --once fcos
full rowmax 2 python scripts/kernel_zoo.py --once rowmax
full iou 3 python scripts/kernel_zoo.py --once iou
full iourot 2 python scripts/kernel_zoo.py --once iourot
full dense 14 python scripts/kernel_zoo.py --once dense
full pre 6 python scripts/kernel_zoo.py --once pre
# the two kernels of the bench step, final code (48-register post-process), eager launches
python bench.py --steps 6 --warmup 3 --no-cpu --no-rot --no-flow --no-ge10k --launch eager > gpurun_out/bench_eager.log 2>&1; echo "bench eager rc=$?"
for k in decode_kernel postprocess_small_kernel; do
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:$k -s 8 -c 1 -o gpurun_out/full_bench_$k -f python bench.py --steps 6 --warmup 3 --no-cpu --no-rot --no-flow --no-ge10k --launch eager > gpurun_out/ncu_f_bench_$k.log 2>&1
  ncu -i gpurun_out/full_bench_$k.ncu-rep --page raw --csv > gpurun_out/full_bench_$k.csv 2>/dev/null
  rm -f gpurun_out/full_bench_$k.ncu-rep
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 96 --warmup 3 --no-cpu --no-rot --no-flow --no-ge10k > gpurun_out/ncu_l_bench.log 2>&1
du -sh gpurun_out
