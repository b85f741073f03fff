mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for f in "" "--local-exchange --no-protocol" "--local-exchange --consumer fused" "--local-exchange --consumer two-kernel"; do
  python bench.py --steps 960 --warmup 10 --no-rot --no-cpu --no-flow --no-ge10k $f 2>>gpurun_out/bench3.err | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','exchange','gpu_launches_per_step')})"
done
python scripts/kernel_zoo.py > gpurun_out/zoo.jsonl 2>gpurun_out/zoo.err; echo "zoo rc=$?"; cat gpurun_out/zoo.jsonl
for sec in rot rotclu atss fcos rowmax iou iourot dense pre decode_yolo decode_rapid; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$sec.csv python scripts/kernel_zoo.py --once $sec > gpurun_out/ncu_l_$sec.log 2>&1
done
for sec in rot atss fcos rowmax iou iourot pre; do
  ncu --set full --import-source on --clock-control none -k regex:mydet -c 40 -o gpurun_out/full_$sec -f python scripts/kernel_zoo.py --once $sec > gpurun_out/ncu_f_$sec.log 2>&1
done
ls -la gpurun_out | head -50
