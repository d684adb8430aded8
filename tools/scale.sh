#!/bin/bash
# usage: tools/scale.sh <mode> <workload> <N...>   (run under gpurun --gpus 8)
mode=$1; wl=$2; shift 2
for n in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --mode $mode --workload $wl > gpurun_out/scale_${mode}_${wl}_$n.json 2> gpurun_out/scale_${mode}_${wl}_$n.err || tail -5 gpurun_out/scale_${mode}_${wl}_$n.err
  python - <<PY
import json
d=json.load(open("gpurun_out/scale_${mode}_${wl}_$n.json"))
print("$mode $wl N=$n ms/step=%.4f value=%.0f Mtri/s fps=%.0f kernels=%s" % (d["ms_per_step"], d["value"], d["frames_per_s"], {k: round(v*1e3) for k,v in d["kernel_ms"].items()}))
PY
done
