#!/bin/bash
# usage (under gpurun, one GPU): tools/profile.sh <tag> [workload]
# Writes into gpurun_out/: <tag>_launches.csv (every launch with its device time), <tag>_{setup,tile,vertex}.ncu-rep
# (ncu --set full, one launch each, after warm-up). The driver is tools/ab.py (no torch import), default library only.
tag=${1:-prof}; wl=${2:-c3}
lib=${AXR_PROFILE_VARIANT:-default}
mkdir -p gpurun_out
cmd="python tools/ab.py --workload $wl --steps 2 --warmup 2 --no-parity $lib"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_launches.log 2>&1
for k in setup:k_setup_raster tile:k_tile_shade vertex:k_vertex_xform; do
  name=${k%%:*}; pat=${k##*:}
  ncu --set full --clock-control none --import-source on -k regex:$pat -s 3 -c 1 -f -o gpurun_out/${tag}_${name} $cmd > gpurun_out/${tag}_${name}.log 2>&1
done
ls -la gpurun_out/${tag}_*
