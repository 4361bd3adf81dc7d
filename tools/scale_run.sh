#!/bin/bash
# usage: tools/scale_run.sh N   -- both benches on N GPUs of one box (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python tools/train_bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/train_n$N.json
  python bench.py --gpus 1 --steps 20000 --warmup 2000 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/pyramid_n$N.json
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/train_bench.py --steps 20 --warmup 5 2>/dev/null | tail -1 > gpurun_out/train_n$N.json
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20000 --warmup 2000 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/pyramid_n$N.json
fi
python - <<PY
import json
for f in ("gpurun_out/train_n$N.json", "gpurun_out/pyramid_n$N.json"):
    try:
        d = json.load(open(f)); print(f, d["n_gpus"], round(d["value"], 1), d["unit"], round(d["ms_per_step"], 4), "ms/step")
    except Exception as e: print(f, "FAILED", e)
PY
