# 8-GPU box: N=1 on the same box, then N=8 / 4 / 2 with the exchange protocol (default) and N=8 without; the dense-scene
# strong-scaling script (configs[4]) at 1 and 8 GPUs
mkdir -p gpurun_out
show() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', {k:d.get(k) for k in ('n_gpus','value','ms_per_step','exchange','exchange_verified','gpu_launches_per_step')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])"; }
A="--steps 960 --warmup 10 --no-rot --no-cpu --no-flow --no-ge10k"
python bench.py $A 2>gpurun_out/s1.err | show n1
run() { n=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n $A "$@" 2>gpurun_out/s$n.err | show "n$n $*"; }
run 8
run 8 --no-protocol
run 4
run 2
python scripts/dense_sharded.py 704 2>/dev/null | tail -1 | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29549 scripts/dense_sharded.py 704 2>/dev/null | tail -1 | cut -c1-300
