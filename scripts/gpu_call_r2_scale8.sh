# 8-GPU box: N=1 on the same box, then N=8 with and without the exchange protocol, with and without multicast
mkdir -p gpurun_out
show() { tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', {k:d.get(k) for k in ('n_gpus','value','ms_per_step','exchange','exchange_verified','gpu_launches_per_step')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])"; }
A="--steps 960 --warmup 10 --no-rot --no-cpu --no-flow --no-ge10k"
python bench.py $A 2>gpurun_out/s1.err | show n1
for f in "" "--no-protocol" "--no-protocol --no-multicast"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 $A $f 2>gpurun_out/s8.err | show "n8 $f"
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 $A 2>gpurun_out/s4.err | show n4
tail -3 gpurun_out/s8.err
