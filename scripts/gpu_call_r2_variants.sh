# A/B of variant builds (mydetection_b200/_tune/*.so, MYDET_LIB) on the bench step; prints value / ms_per_step per variant
mkdir -p gpurun_out
B="python bench.py --steps 960 --warmup 10 --no-rot --no-cpu --no-flow --no-ge10k"
show() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', {k:d[k] for k in ('value','ms_per_step','exchange')}, d['roofline']['kernel_ms_decode_only'], d['matches_oracle'])"; }
for v in ${VARIANTS:-nreg32 nreg36 nreg44 nreg48}; do MYDET_LIB=$PWD/mydetection_b200/_tune/libmydet_$v.so $B | show $v; done
MYDET_LIB=$PWD/mydetection_b200/_tune/libmydet_nreg48.so $B --pp-priority 0 | show nreg48_prio0
MYDET_LIB=$PWD/mydetection_b200/_tune/libmydet_nreg32.so $B --pp-priority 0 | show nreg32_prio0
MYDET_LIB=$PWD/mydetection_b200/_tune/libmydet_nreg40.so $B --decode-streams 3 | show nreg40_dec3
