# ncu evidence for the kernels changed at the end of round 2 (zero-numerator divide skip, vector stores, programmatic
# dependent launch): each ncu pass wraps a command that has already exited 0 without ncu in this same call.
mkdir -p gpurun_out
python scripts/kernel_zoo.py iou rowmax eager > gpurun_out/zoo_late.jsonl 2>gpurun_out/zoo_late.err; echo "zoo rc=$?"
full() {   # name, launch cap, command...
  local name=$1 cap=$2; shift 2
  ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:mydet -c $cap -o gpurun_out/late_$name -f "$@" > gpurun_out/ncu_late_$name.log 2>&1
  ncu -i gpurun_out/late_$name.ncu-rep --page raw --csv > gpurun_out/late_$name.csv 2>/dev/null
  python scripts/hot_lines.py gpurun_out/late_$name.ncu-rep > gpurun_out/late_hot_$name.txt 2>/dev/null
  rm -f gpurun_out/late_$name.ncu-rep
}
full iou 4 python scripts/kernel_zoo.py --once iou
full rowmax 1 python scripts/kernel_zoo.py --once rowmax
python bench.py --steps 6 --warmup 3 --no-cpu --no-rot --no-flow --no-ge10k --launch eager > gpurun_out/bench_eager_late.log 2>&1; echo "bench eager rc=$?"
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:postprocess_small_kernel -s 8 -c 1 -o gpurun_out/late_pp -f python bench.py --steps 6 --warmup 3 --no-cpu --no-rot --no-flow --no-ge10k --launch eager > gpurun_out/ncu_late_pp.log 2>&1
ncu -i gpurun_out/late_pp.ncu-rep --page raw --csv > gpurun_out/late_pp.csv 2>/dev/null
rm -f gpurun_out/late_pp.ncu-rep
du -sh gpurun_out
