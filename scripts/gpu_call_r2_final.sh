mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-200
for sec in rot rotbench rot1 rotclu; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$sec.csv python scripts/kernel_zoo.py --once $sec > /dev/null 2>&1
done
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:mydet -c 24 -o gpurun_out/full_rotbench -f python scripts/kernel_zoo.py --once rotbench > gpurun_out/ncu_f_rotbench.log 2>&1
ncu -i gpurun_out/full_rotbench.ncu-rep --page raw --csv > gpurun_out/full_rotbench.csv 2>/dev/null
python scripts/hot_lines.py gpurun_out/full_rotbench.ncu-rep > gpurun_out/hot_rotbench.txt 2>/dev/null
rm -f gpurun_out/full_rotbench.ncu-rep
python scripts/kernel_zoo.py rot rotbench rot1 rotclu > gpurun_out/zoo_rot.jsonl 2>/dev/null
du -sh gpurun_out
