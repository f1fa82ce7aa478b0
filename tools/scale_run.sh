#!/bin/bash
# Runs on an 8-GPU box: bench.py at N = 1, 2, 4, 8 back to back (the multi-GPU table of DESIGN.md comes from ONE box),
# then the multi-GPU tests.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-r02}
nvidia-smi topo -m > gpurun_out/${TAG}_topo.txt 2>&1
python bench.py --gpus 1 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "n1 rc=$?"
for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
  echo "n$N rc=$?"
done
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q > gpurun_out/${TAG}_mgpu_pytest.log 2>&1; echo "mgpu tests rc=$?"
tail -5 gpurun_out/${TAG}_mgpu_pytest.log
