#!/bin/bash
# usage: scripts/scale.sh N  -> runs the C2 and C5 benches on N GPUs, JSON lines into gpurun_out/scale_N.jsonl
N=$1
OUT=gpurun_out/scale_$N.jsonl
: > $OUT
for wl in c2_bulk_20kx200 c5_allref_30kx20k; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --workload $wl >> $OUT 2> gpurun_out/scale_$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --workload $wl >> $OUT 2> gpurun_out/scale_$N.err
  fi
  echo "rc=$? $wl"
done
python - <<PY
import json
for l in open("$OUT"):
    l=l.strip()
    if not l.startswith("{"): continue
    d=json.loads(l)
    print(d["config"]["workload"], "N=",d["n_gpus"], "value=%.3e"%d["value"], "ms/step=%.2f"%d["ms_per_step"], "e2e ms=%.2f"%d["e2e"]["ms_per_step"], d["stage_ms"], "frac=%.3f"%d["roofline"]["frac"])
PY
