#!/bin/bash
# smoke + reference arm + N-GPU bench under torchrun (N = $1, default 2)
N=${1:-2}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/smoke.log | cut -c1-400
python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-500 gpurun_out/bench_ref.json
python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 rc=$?"; cut -c1-250 gpurun_out/bench_n1.json
for n in 2 4 8; do
  if [ $n -le $N ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
    echo "n$n rc=$?"; cut -c1-250 gpurun_out/bench_n$n.json
  fi
done
nproc; lscpu | grep "Model name"
