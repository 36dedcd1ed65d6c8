#!/bin/bash
# where does the first e2e interval at N=2 go?  (a) one process, 64 seeds; (b) torchrun 2 ranks x 64 seeds, twice
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 python bench.py --gpus 1 --seeds 64 --steps 5 --warmup 3 --no-cpu --no-env-roofline > gpurun_out/r2w3_1proc_64seeds.json 2> gpurun_out/r2w3_a.err
for i in 1 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$i bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --no-env-roofline > gpurun_out/r2w3_2gpu_$i.json 2> gpurun_out/r2w3_b$i.err
done
python - <<'PY'
import json
for f in ("r2w3_1proc_64seeds","r2w3_2gpu_1","r2w3_2gpu_2"):
    d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
    print(f, round(d["value"]/1e6,2), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"]/1e6,2), d["e2e"]["wall_split_rank0"])
PY
