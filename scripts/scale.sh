#!/bin/bash
# usage: scripts/scale.sh N  -> bench.py on N GPUs exactly as the driver launches it (default workload = BASELINE
# configs[4], 30k x 30k gene pairs x 20k cells); the JSON line goes to gpurun_out/scale_N.json
N=$1
OUT=gpurun_out/scale_$N.json
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $OUT 2> gpurun_out/scale_$N.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 10 --warmup 3 > $OUT 2> gpurun_out/scale_$N.err
fi
echo "rc=$?"
python - <<PY
import json
d = json.loads(open("$OUT").read().strip().splitlines()[-1])
print("N=", d["n_gpus"], "value=%.3e" % d["value"], "ms/step=%.2f" % d["ms_per_step"], "e2e ms=%.2f" % d["e2e"]["ms_per_step"],
      d["stage_ms"], "frac=%.3f" % d["roofline"]["frac"], "parity", d["parity"]["ok"])
PY
