#!/bin/bash
# usage: tools/scale.sh <mode> <transport> <workload> <N...>   (run under gpurun --gpus 8)
mode=$1; tr=$2; wl=$3; shift 3
for n in "$@"; do
  out=gpurun_out/scale_${mode}_${tr}_${wl}_$n
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --mode $mode --transport $tr --workload $wl > $out.json 2> $out.err || tail -5 $out.err
  python - <<PY
import json
d=json.loads(open("$out.json").read())
print("$mode/$tr $wl N=$n ms/step=%.4f value=%.0f Mtri/s fps=%.0f kernels=%s" % (d["ms_per_step"], d["value"], d["frames_per_s"], {k: round(v*1e3) for k,v in d["kernel_ms"].items()}))
PY
done
