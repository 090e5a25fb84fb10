#!/bin/bash
# One multi-GPU measurement session on an N-GPU box (gpurun --gpus N): pipeline tests across devices, bench.py at N ranks,
# the host-link ceiling at 1/2/4/N ranks, the CLI with --gpus N.  Everything lands in gpurun_out/.
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -14 > gpurun_out/r02_topo.txt
lscpu | grep -i "model name\|^CPU(s)\|numa" > gpurun_out/r02_cpu.txt
python -m pytest tests/test_gpu_parity.py -x -q -k "pipeline" 2>&1 | tail -2
rm -f gpurun_out/r02_h2d_peak.jsonl
for n in 1 2 4 $N; do
  [ $n -le $N ] && timeout 120 $TR --nproc-per-node $n --master-port $((29600+n)) tools/h2d_peak.py gpurun_out/r02_h2d_peak.jsonl 2>/dev/null | tail -1
done
timeout 400 $TR --nproc-per-node $N --master-port 29555 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench N=$N rc=$?"
grep -v "Time spent" gpurun_out/r02_bench_n$N.err | tail -3
timeout 300 python tools/cli_throughput.py 1800 gpurun_out/r02_cli_n$N.json 2>&1 | tail -8
