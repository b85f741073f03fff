#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out, in one gpurun.
#   gpurun --timeout 900 -- 'bash scripts/first_gpu_call.sh'
# Results land in gpurun_out/ (copy the summaries worth keeping into profiles/).
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
# 1. the verified path must still be green, then the pre-processing tests on their own with the xfail marker ignored
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc=$?"
python -m pytest tests/test_zz_preprocess_gpu.py tests/test_zz_fullmodel_gpu.py tests/test_zz_sweep_fixpoint_gpu.py -q -m gpu --runxfail > gpurun_out/pytest_preprocess.log 2>&1; echo "preprocess rc=$?"
tail -3 gpurun_out/pytest_gpu.log gpurun_out/pytest_preprocess.log
# 2. numbers: pre-processing (1080p -> 608 and CEPDOF-like 2048 -> 1024), the >= 10k-candidate bench point
python scripts/preprocess_bench.py 64 1080 1920 608 > gpurun_out/preprocess_bench.log 2>&1
python scripts/preprocess_bench.py 32 2048 2048 1024 >> gpurun_out/preprocess_bench.log 2>&1
python scripts/preprocess_bench.py 1 1080 1920 608 >> gpurun_out/preprocess_bench.log 2>&1
cat gpurun_out/preprocess_bench.log
python bench.py --img-size 768 --steps 960 --warmup 20 --no-rot > gpurun_out/bench_768.log 2>&1; tail -1 gpurun_out/bench_768.log | cut -c1-900
# 2b. the experimental fixed-point sweep against the block sweep (rotated NMS us / image, dense scenes)
for fx in 0 1; do MYDET_SWEEP_FIXPOINT=$fx python scripts/rot_bench.py 1; MYDET_SWEEP_FIXPOINT=$fx python scripts/dense_bench.py 704 64; MYDET_SWEEP_FIXPOINT=$fx python scripts/dense_bench.py 1536 8; done > gpurun_out/sweep_fixpoint_bench.log 2>&1
cat gpurun_out/sweep_fixpoint_bench.log
# 3. launch list of the pre-processing kernels (only after the plain run above exited 0)
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/preprocess_launches.csv \
    python scripts/preprocess_bench.py 64 1080 1920 608 > gpurun_out/ncu_preprocess.log 2>&1
grep -c "mydet::pre" gpurun_out/preprocess_launches.csv
